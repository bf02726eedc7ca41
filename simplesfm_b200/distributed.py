"""Multi-GPU plumbing for the matching path: one process per GPU, torch.distributed.

The path shards naturally, so there is no collective inside the per-pair kernels:
  phase 1  images are split in contiguous blocks, each rank runs SuperPoint on its block;
  phase 2  ONE all-gather of fixed-capacity feature records (NCCL over NVLink on the GPU box);
  phase 3  the lexicographic pair list (matcher.py:143-144) is cut into contiguous ranges aligned
           to `super_glue_batch` chunks, so chunk composition -- and therefore the reference's
           chunk-minimum truncation and train-mode BatchNorm statistics -- is identical to a
           single-GPU run at any world size;
  phase 4  match tables are gathered (variable length, two all-gathers: lengths then rows).
The reference has no distributed code at all; this module is new, not a mirror.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_pairs: int, chunk: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the pair list owned by `rank`: contiguous, aligned to `chunk`, balanced to
    within one chunk."""
    chunk = max(1, int(chunk))
    n_chunks = (n_pairs + chunk - 1) // chunk
    base, rem = divmod(n_chunks, world)
    c_lo = rank * base + min(rank, rem)
    c_hi = c_lo + base + (1 if rank < rem else 0)
    return min(n_pairs, c_lo * chunk), min(n_pairs, c_hi * chunk)


def image_block(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) block of images whose features `rank` extracts."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _comm_device(like: torch.device) -> torch.device:
    return like if dist.get_backend() == "nccl" else torch.device("cpu")


def allgather_feature_records(desc: torch.Tensor, kpts: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor,
                              n_images: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-gather per-rank record blocks (desc (n_r,256,kcap), kpts (n_r,kcap,2), scores (n_r,kcap),
    counts (n_r,)) produced with `image_block` into full (n_images, ...) records on every rank.
    Blocks are padded to the largest block so a single all_gather_into_tensor per array suffices."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [image_block(n_images, r, world) for r in range(world)]
    blk = max(hi - lo for lo, hi in sizes)
    dev = _comm_device(desc.device)
    outs = []
    for t in (desc, kpts, scores, counts):
        pad = torch.zeros((blk,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[:t.shape[0]].copy_(t)
        full = torch.empty((world * blk,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(full, pad)
        parts = [full[r * blk:r * blk + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
        outs.append(torch.cat(parts, 0).to(desc.device))
    return tuple(outs)


def gather_match_tables(local: Dict[Tuple[int, int], np.ndarray], match_list: np.ndarray
                        ) -> Dict[Tuple[int, int], np.ndarray]:
    """All ranks contribute the tables of their pair range; every rank returns the full table in
    pair-list order (identical at any world size)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    keys = list(local.keys())
    lens = torch.tensor([len(local[k]) for k in keys], dtype=torch.int64)
    rows = (np.concatenate([local[k].reshape(-1, 2) for k in keys], 0) if keys else np.zeros((0, 2), np.uint32))
    meta = torch.tensor([len(keys), int(lens.sum())], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas = [m.cpu().tolist() for m in metas]
    max_k = max(m[0] for m in metas)
    max_r = max(m[1] for m in metas)
    keyt = torch.zeros(max(max_k, 1), 3, dtype=torch.int64)
    if keys:
        keyt[:len(keys), 0] = torch.tensor([int(k[0]) for k in keys])
        keyt[:len(keys), 1] = torch.tensor([int(k[1]) for k in keys])
        keyt[:len(keys), 2] = lens
    rowt = torch.zeros(max(max_r, 1), 2, dtype=torch.int64)
    if len(rows):
        rowt[:len(rows)] = torch.from_numpy(rows.astype(np.int64))
    keyt, rowt = keyt.to(dev), rowt.to(dev)
    all_keys = [torch.zeros_like(keyt) for _ in range(world)]
    all_rows = [torch.zeros_like(rowt) for _ in range(world)]
    dist.all_gather(all_keys, keyt)
    dist.all_gather(all_rows, rowt)
    merged: Dict[Tuple[int, int], np.ndarray] = {}
    for r in range(world):
        kt = all_keys[r].cpu().numpy()
        rt = all_rows[r].cpu().numpy()
        off = 0
        for i in range(metas[r][0]):
            a, b, n = kt[i]
            merged[(np.int64(a), np.int64(b))] = rt[off:off + n].astype(np.uint32)
            off += n
    # pair-list order, like a single-process run
    out = {}
    for a, b in match_list:
        k = (np.int64(a), np.int64(b))
        if k in merged:
            out[k] = merged[k]
    return out
