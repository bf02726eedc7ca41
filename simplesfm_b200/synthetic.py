"""Seeded synthetic state-dicts, images and feature scenes.

The pretrained `superpoint_v1.pth` / `superglue_{indoor,outdoor}.pth` blobs are stripped
from the reference tree (`.MISSING_LARGE_BLOBS:1-4`) and there is no network, so parity
tests and benchmarks run on random weights with the *identical key layout*
(shapes follow `simple_sfm/models/superpoint.py:84-99` and
`simple_sfm/models/superglue.py:205-231`).  The weights are scaled so that detector logits
spread (default init gives a near-uniform softmax and massive score ties) and so that
SuperGlue produces confident matches on related feature sets.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

SP_CONVS = (("conv1a", 1, 64, 3), ("conv1b", 64, 64, 3), ("conv2a", 64, 64, 3), ("conv2b", 64, 64, 3),
            ("conv3a", 64, 128, 3), ("conv3b", 128, 128, 3), ("conv4a", 128, 128, 3), ("conv4b", 128, 128, 3),
            ("convPa", 128, 256, 3), ("convPb", 256, 65, 1), ("convDa", 128, 256, 3), ("convDb", 256, 256, 1))

KENC_CH = (3, 32, 64, 128, 256, 256)
GNN_LAYERS = 18


def superpoint_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, cin, cout, k in SP_CONVS:
        std = math.sqrt(2.0 / (cin * k * k))
        if name == "conv1a":
            std *= 3.0          # input is a single [0,1] channel: lift the signal
        if name == "convPb":
            std *= 1.2          # spread the 65 detector logits
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * std
        sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.05
    return sd


def superglue_state_dict(seed: int = 0, variant: str = "indoor") -> Dict[str, torch.Tensor]:
    """`variant` only perturbs the seed: indoor and outdoor checkpoints share one layout."""
    g = torch.Generator().manual_seed(seed * 2 + (1 if variant == "outdoor" else 0) + 1000)
    sd: Dict[str, torch.Tensor] = {"bin_score": torch.tensor(1.0 + 0.5 * (variant == "outdoor"))}

    def conv(prefix, cin, cout, gain=1.0, zero_bias=False):
        sd[prefix + ".weight"] = torch.randn(cout, cin, 1, generator=g) * (gain / math.sqrt(cin))
        sd[prefix + ".bias"] = torch.zeros(cout) if zero_bias else torch.randn(cout, generator=g) * 0.05

    def bn(prefix, c):
        sd[prefix + ".weight"] = 0.5 + torch.rand(c, generator=g)
        sd[prefix + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[prefix + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)

    for i in range(5):
        last = i == 4
        conv(f"kenc.encoder.{3 * i}", KENC_CH[i], KENC_CH[i + 1], gain=0.05 if last else 1.4, zero_bias=last)
        if not last:
            bn(f"kenc.encoder.{3 * i + 1}", KENC_CH[i + 1])
    for L in range(GNN_LAYERS):
        p = f"gnn.layers.{L}"
        conv(p + ".attn.merge", 256, 256, gain=1.0)
        for j in range(3):
            conv(p + f".attn.proj.{j}", 256, 256, gain=4.0 if j < 2 else 1.0)
        conv(p + ".mlp.0", 512, 512, gain=1.4)
        bn(p + ".mlp.1", 512)
        conv(p + ".mlp.3", 512, 256, gain=0.04, zero_bias=True)
    # final projection: a scaled random orthogonal map keeps descriptor similarity so that
    # related keypoints score high (scores = <m0, m1> / 16)
    # LAPACK's QR rounds differently for different thread counts (observed: 1 thread vs 8 differ by 1e-5), and
    # torchrun starts every rank with OMP_NUM_THREADS=1: pin the count the golden vectors were generated with, so
    # that every process -- tests, bench ranks, golden script -- builds bit-identical weights
    n_threads = torch.get_num_threads()
    torch.set_num_threads(8)
    try:
        q, _ = torch.linalg.qr(torch.randn(256, 256, generator=g))
    finally:
        torch.set_num_threads(n_threads)
    sd["final_proj.weight"] = (q * 10.0)[:, :, None].contiguous()
    sd["final_proj.bias"] = torch.randn(256, generator=g) * 0.05
    return sd


def textured_image(seed: int, H: int = 480, W: int = 640, shift: Tuple[int, int] = (0, 0)) -> torch.Tensor:
    """Smooth random texture in [0,1], (1,1,H,W) float32; `shift` crops a translated window of
    the same underlying texture (overlapping views of one scene)."""
    g = torch.Generator().manual_seed(seed)
    pad = 64
    base = torch.rand(1, 1, H + 2 * pad, W + 2 * pad, generator=g)
    k = torch.tensor([1., 4., 6., 4., 1.])
    k2 = (k[:, None] * k[None, :] / 256.0)[None, None]
    x = base
    for _ in range(2):
        x = torch.nn.functional.conv2d(x, k2, padding=2)
    x = (x - x.amin()) / (x.amax() - x.amin())
    x = 0.6 * x + 0.4 * base
    dy, dx = shift
    return x[:, :, pad + dy:pad + dy + H, pad + dx:pad + dx + W].contiguous()


def feature_scene(n_images: int, n_kp: int, seed: int = 0, H: int = 480, W: int = 640,
                  noise: float = 0.25, shared: float = 0.7, device="cpu"
                  ) -> List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
    """Direct synthetic features (BASELINE config 2/4): per image
    (descriptors (256,K) unit-norm f32, keypoints (K,2) integer-valued f32 in the border-safe
    window, scores (K,) sorted descending).  A fraction `shared` of each image's keypoints are
    noisy views of a common landmark pool so that true correspondences exist."""
    g = torch.Generator().manual_seed(seed + 77)
    pool = max(n_kp * 5 // 4, 16)
    land_desc = torch.nn.functional.normalize(torch.randn(256, pool, generator=g), dim=0)
    land_xy = torch.stack([torch.randint(4, W - 4, (pool,), generator=g),
                           torch.randint(4, H - 4, (pool,), generator=g)], 1).float()
    out = []
    n_sh = int(n_kp * shared)
    for _ in range(n_images):
        sel = torch.randperm(pool, generator=g)[:n_sh]
        d = land_desc[:, sel] + noise * torch.randn(256, n_sh, generator=g) / 16.0
        xy = land_xy[sel] + torch.randint(-3, 4, (n_sh, 2), generator=g).float()
        d2 = torch.randn(256, n_kp - n_sh, generator=g)
        xy2 = torch.stack([torch.randint(4, W - 4, (n_kp - n_sh,), generator=g),
                           torch.randint(4, H - 4, (n_kp - n_sh,), generator=g)], 1).float()
        d = torch.nn.functional.normalize(torch.cat([d, d2], 1), dim=0)
        xy = torch.cat([xy, xy2], 0)
        xy[:, 0].clamp_(4, W - 5)
        xy[:, 1].clamp_(4, H - 5)
        perm = torch.randperm(n_kp, generator=g)
        sc = torch.sort(torch.rand(n_kp, generator=g) * 0.5 + 0.005, descending=True).values
        out.append((d[:, perm].contiguous().to(device), xy[perm].contiguous().to(device), sc.to(device)))
    return out


def save_state_dict(sd: Dict[str, torch.Tensor], path: str) -> str:
    torch.save(sd, path)
    return path
