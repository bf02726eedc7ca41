"""Multi-GPU plumbing for the matching path: one process per GPU, torch.distributed.

The path shards naturally, so there is no collective inside the per-pair kernels:
  phase 1  images are dealt round-robin (image i -> rank i mod G); each rank decodes and runs SuperPoint on
           its own images only;
  phase 2  ONE all-gather of fixed-capacity feature records (NCCL over NVLink on the GPU box): afterwards
           every rank holds the features of every image, in image order;
  phase 3  the lexicographic pair list (matcher.py:143-144) is cut into contiguous ranges aligned to
           `super_glue_batch` chunks, so chunk composition -- and therefore the reference's chunk-minimum
           truncation and train-mode BatchNorm statistics -- is identical to a single-GPU run at any world size;
  phase 4  match tables travel as compacted uint32 rows + per-pair lengths to RANK 0 ONLY (the single sqlite
           writer): one small gather of sizes, one gather of lengths, one gather of rows; the other ranks
           do no per-pair work for pairs that are not theirs.
The reference has no distributed code at all; this module is new, not a mirror.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------- partitioning
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
    """[lo, hi) contiguous block of images (kept for callers that know the image count up front)."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of_image(i: int, world: int) -> int:
    """Round-robin ownership: works for sources of unknown length (video streams) and balances ragged tails."""
    return i % world


def active_group(enabled: bool, group=None):
    """(group, rank, world) when pair sharding is switched on and a process group exists, else (None, 0, 1)."""
    if not enabled or not (dist.is_available() and dist.is_initialized()):
        return None, 0, 1
    world = dist.get_world_size(group)
    if world <= 1:
        return None, 0, 1
    return (group if group is not None else dist.group.WORLD), dist.get_rank(group), world


def _comm_device(like: torch.device, group) -> torch.device:
    return like if dist.get_backend(group) == "nccl" else torch.device("cpu")


# ---------------------------------------------------------------------------------------------- phase 2
class Records(NamedTuple):
    """Feature records of every image, in image order, resident on this rank's device."""
    desc: torch.Tensor      # (n, 256, kcap) f32
    kpts: torch.Tensor      # (n, kcap, 2) f32
    scores: torch.Tensor    # (n, kcap) f32
    counts: List[int]       # keypoints per image
    hw: List[Tuple[int, int]]


def allgather_records(local: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Tuple[int, int]]],
                      n_local_slots: Optional[int], device: torch.device, group=None) -> Records:
    """`local[s]` = (descriptors (256, K), keypoints (K, 2), scores (K,), (H, W)) of the s-th image this rank owns
    (global image index s * world + rank).  One all-gather of packed records
    [desc 256*kcap | kpts 2*kcap | scores kcap | K, H, W, valid] per slot; returns the records of all images
    in image order on every rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cdev = _comm_device(device, group)
    # sizes: the keypoint capacity and the slot count must agree on every rank (two scalars, one tiny all-reduce)
    kmax = max([int(k.shape[0]) for _, k, _, _ in local] + [1])
    sizes = torch.tensor([kmax, len(local) if n_local_slots is None else n_local_slots], dtype=torch.int64, device=cdev)
    dist.all_reduce(sizes, op=dist.ReduceOp.MAX, group=group)
    kcap = (int(sizes[0]) + 3) // 4 * 4
    slots = int(sizes[1])
    rec_len = 259 * kcap + 4
    send = torch.zeros(max(slots, 1), rec_len, dtype=torch.float32, device=device)
    meta_host = torch.zeros(send.shape[0], 4, dtype=torch.int32)
    for s, (d, k, sc, hw) in enumerate(local):
        n = int(k.shape[0])
        if n:
            send[s, :256 * kcap].view(256, kcap)[:, :n].copy_(d, non_blocking=True)
            send[s, 256 * kcap:258 * kcap].view(kcap, 2)[:n].copy_(k, non_blocking=True)
            send[s, 258 * kcap:259 * kcap][:n].copy_(sc, non_blocking=True)
        meta_host[s] = torch.tensor([n, int(hw[0]), int(hw[1]), 1], dtype=torch.int32)
    send[:, 259 * kcap:].copy_(meta_host.view(torch.float32))      # integers travel as raw bits
    send_c = send.to(cdev)
    full = torch.empty(world * send.shape[0], rec_len, dtype=torch.float32, device=cdev)
    dist.all_gather_into_tensor(full, send_c, group=group)
    full = full.to(device)
    meta_all = full[:, 259 * kcap:].contiguous().view(torch.int32).cpu().numpy().reshape(world, send.shape[0], 4)
    # image i lives in slot i // world of rank i % world
    order = [(r, s) for s in range(send.shape[0]) for r in range(world) if meta_all[r, s, 3] == 1]
    rows = torch.tensor([r * send.shape[0] + s for r, s in order], dtype=torch.int64, device=device)
    picked = full.index_select(0, rows)
    n = len(order)
    desc = picked[:, :256 * kcap].reshape(n, 256, kcap).contiguous()
    kpts = picked[:, 256 * kcap:258 * kcap].reshape(n, kcap, 2).contiguous()
    scores = picked[:, 258 * kcap:259 * kcap].contiguous()
    counts = [int(meta_all[r, s, 0]) for r, s in order]
    hw = [(int(meta_all[r, s, 1]), int(meta_all[r, s, 2])) for r, s in order]
    return Records(desc, kpts, scores, counts, hw)


# ---------------------------------------------------------------------------------------------- phase 4
class CompactTables(NamedTuple):
    """Match tables of a run of pairs without per-pair objects: `rows[offsets[p]:offsets[p+1]]` is the (L, 2)
    uint32 table of `pairs[p]`."""
    pairs: np.ndarray     # (P, 2) int64, 1-based image ids
    lens: np.ndarray      # (P,) int64
    rows: np.ndarray      # (sum(lens), 2) uint32

    def to_dict(self):
        """{(np.int64 id1, np.int64 id2): (L, 2) uint32} in pair order -- the reference's match table
        (matcher.py:103-109)."""
        # plain slices of the flat row array (views): `np.split` and per-element array indexing cost 2.5x as much,
        # and on the database writer's rank this loop runs over the pairs of ALL ranks
        offs = np.concatenate([[0], np.cumsum(self.lens)]).tolist()
        rows = self.rows
        a, b = list(self.pairs[:, 0]), list(self.pairs[:, 1])
        return {(a[i], b[i]): rows[offs[i]:offs[i + 1]] for i in range(len(a))}


def concat_tables(parts: Sequence[CompactTables]) -> CompactTables:
    parts = [p for p in parts if len(p.pairs)]
    if not parts:
        return CompactTables(np.zeros((0, 2), np.int64), np.zeros((0,), np.int64), np.zeros((0, 2), np.uint32))
    return CompactTables(np.concatenate([p.pairs for p in parts]), np.concatenate([p.lens for p in parts]),
                         np.concatenate([p.rows for p in parts]))


def gather_tables_to_root(local: CompactTables, wanted: np.ndarray, chunk: int, device: torch.device, group=None,
                          dst: int = 0) -> Optional[CompactTables]:
    """Every rank contributes the compact tables of its `shard_range` of `wanted`; rank `dst` returns the tables of
    all pairs in pair-list order, the other ranks return None.  Payload: uint32 rows + int32 lengths, padded to the
    largest rank (shards are balanced to within one chunk, so the padding is a few per cent)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cdev = _comm_device(device, group)
    sizes = torch.tensor([len(local.lens), len(local.rows)], dtype=torch.int64, device=cdev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [s.cpu().tolist() for s in all_sizes]
    max_p = max(max(s[0] for s in all_sizes), 1)
    max_r = max(max(s[1] for s in all_sizes), 1)
    lens_t = torch.zeros(max_p, dtype=torch.int32)
    lens_t[:len(local.lens)] = torch.from_numpy(local.lens.astype(np.int32))
    rows_t = torch.zeros(max_r, 2, dtype=torch.int32)
    if len(local.rows):
        rows_t[:len(local.rows)] = torch.from_numpy(local.rows.view(np.int32))
    lens_t, rows_t = lens_t.to(cdev), rows_t.to(cdev)
    is_dst = rank == dst
    lens_all = [torch.empty_like(lens_t) for _ in range(world)] if is_dst else None
    rows_all = [torch.empty_like(rows_t) for _ in range(world)] if is_dst else None
    dist.gather(lens_t, lens_all, dst=dst, group=group)
    dist.gather(rows_t, rows_all, dst=dst, group=group)
    if not is_dst:
        return None
    parts = []
    for r in range(world):
        lo, hi = shard_range(len(wanted), chunk, r, world)
        n_p, n_r = all_sizes[r]
        assert n_p == hi - lo, f"rank {r} sent {n_p} pairs for the shard [{lo}, {hi})"
        parts.append(CompactTables(np.asarray(wanted[lo:hi], dtype=np.int64).reshape(-1, 2),
                                   lens_all[r][:n_p].cpu().numpy().astype(np.int64),
                                   rows_all[r][:n_r].cpu().numpy().view(np.uint32).reshape(-1, 2)))
    return concat_tables(parts)


def assert_same_everywhere(values: Sequence[int], device: torch.device, group=None, what: str = "input") -> None:
    """Sharding assumes every rank was handed the same scene; compare a few integers across ranks."""
    cdev = _comm_device(device, group)
    t = torch.tensor(list(values), dtype=torch.int64, device=cdev)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if not torch.equal(lo, hi):
        raise RuntimeError(f"pair sharding: ranks disagree on the {what} ({lo.tolist()} vs {hi.tolist()}); every rank "
                           f"must be given the same frames / names")
