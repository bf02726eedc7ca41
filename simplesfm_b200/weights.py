"""State-dict -> packed float32 buffers in the order `include/sfm_b200.h` documents.

Key layouts follow the reference constructors (`simple_sfm/models/superpoint.py:84-99`,
`simple_sfm/models/superglue.py:205-231`); loading is strict like `load_state_dict`
(superpoint.py:101, superglue.py:233): missing or unexpected keys raise RuntimeError.
"""
from __future__ import annotations

from typing import Dict, List, Union

import torch

SP_ORDER = ("conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b",
            "convPa", "convPb", "convDa", "convDb")


def load_state_dict_file(path_or_dict: Union[str, Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    if isinstance(path_or_dict, dict):
        return path_or_dict
    return torch.load(str(path_or_dict), map_location="cpu")


def _strict(sd: Dict[str, torch.Tensor], expected: List[str], who: str) -> None:
    missing = [k for k in expected if k not in sd]
    unexpected = [k for k in sd if k not in set(expected)]
    if missing or unexpected:
        raise RuntimeError(f"Error(s) in loading state_dict for {who}: "
                           f"Missing key(s): {missing[:8]}; Unexpected key(s): {unexpected[:8]}")


def superpoint_keys() -> List[str]:
    return [f"{n}.{s}" for n in SP_ORDER for s in ("weight", "bias")]


def superglue_keys() -> List[str]:
    keys = ["bin_score"]
    for i in range(5):
        keys += [f"kenc.encoder.{3 * i}.weight", f"kenc.encoder.{3 * i}.bias"]
        if i < 4:
            keys += [f"kenc.encoder.{3 * i + 1}.{s}" for s in
                     ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")]
    for L in range(18):
        p = f"gnn.layers.{L}"
        for j in range(3):
            keys += [f"{p}.attn.proj.{j}.weight", f"{p}.attn.proj.{j}.bias"]
        keys += [f"{p}.attn.merge.weight", f"{p}.attn.merge.bias", f"{p}.mlp.0.weight", f"{p}.mlp.0.bias"]
        keys += [f"{p}.mlp.1.{s}" for s in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")]
        keys += [f"{p}.mlp.3.weight", f"{p}.mlp.3.bias"]
    keys += ["final_proj.weight", "final_proj.bias"]
    return keys


def pack_superpoint(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    _strict(sd, superpoint_keys(), "SuperPoint")
    parts = []
    for n in SP_ORDER:
        parts += [sd[n + ".weight"].reshape(-1), sd[n + ".bias"].reshape(-1)]
    return torch.cat([p.detach().to("cpu", torch.float32) for p in parts]).contiguous()


def pack_superglue(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    _strict(sd, superglue_keys(), "SuperGlue")
    names = [k for k in superglue_keys() if not k.endswith("num_batches_tracked")]
    # header order: proj.0, proj.1, proj.2, merge  (superglue_keys() lists them in that order too)
    return torch.cat([sd[k].detach().to("cpu", torch.float32).reshape(-1) for k in names]).contiguous()
