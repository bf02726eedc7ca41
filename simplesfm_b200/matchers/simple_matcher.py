"""Drop-in mirror of `simple_sfm/matchers/simple_matcher.py` (mutual nearest neighbour on the
L2 distance of unit descriptors) running on libsfm_b200."""
__all__ = ['SimpleMatcher']

import numpy as np
import torch

from .. import _lib
from .._lib import check, lib, ptr, require_cuda, stream_ptr


class SimpleMatcher(object):
    def __init__(self, nn_thresh: float):
        self.nn_thresh = nn_thresh

    def match_two_way_cuda(self, desc1: torch.Tensor, desc2: torch.Tensor, nn_thresh: float = None) -> torch.Tensor:
        """
        Two-way nearest neighbour matching (simple_matcher.py:66-114).
        desc1 (256,N), desc2 (256,M) unit-norm CUDA f32 -> (3,L) f32 rows [i, j, distance].
        """
        if nn_thresh is None:
            nn_thresh = self.nn_thresh
        assert desc1.shape[0] == desc2.shape[0]
        if desc1.shape[0] != 256:
            raise _lib.SfmError(-5, "SimpleMatcher: descriptor dimension must be 256")
        d1 = require_cuda(desc1, "desc1", torch.float32)
        d2 = require_cuda(desc2, "desc2", torch.float32)
        N, M = d1.shape[1], d2.shape[1]
        if N == 0 or M == 0:
            return torch.zeros((3, 0), device=d1.device)
        if nn_thresh < 0.0:
            raise ValueError('\'nn_thresh\' should be non-negative')
        L = lib()
        cap = min(N, M)
        out = torch.empty(3, cap, dtype=torch.float32, device=d1.device)
        count = torch.empty(1, dtype=torch.int32, device=d1.device)
        ws = torch.empty(max(L.sfm_simple_match_workspace_bytes(N, M), 256), dtype=torch.uint8, device=d1.device)
        with torch.cuda.device(d1.device):
            check(L.sfm_simple_match(ptr(d1), N, N, ptr(d2), M, M, float(nn_thresh), ptr(out), cap, ptr(count),
                                     ptr(ws), ws.numel(), stream_ptr(d1.device)))
        return out[:, :int(count.item())]

    def match_pairs(self, descriptors, pairs0: np.ndarray, super_chunk: int = 2048):
        """Mutual-NN tables of many pairs without a host synchronisation per pair (the reference loops
        `match_two_way_cuda(...).t()[:, :2]` with a blocking read-back each, matcher.py:115-131).
        descriptors: list of (256, K_i) CUDA f32; pairs0 (P, 2) 0-based image indices.
        Returns (rows (sum L, 2) uint32 [index0, index1], lens (P,) int64): every pair is queued into one output
        buffer per super-chunk, then counts and rows come back with one device->host copy each."""
        pairs0 = np.asarray(pairs0, dtype=np.int64).reshape(-1, 2)
        L = lib()
        rows_out, lens_out = [], []
        if len(pairs0) == 0:
            return np.zeros((0, 2), np.uint32), np.zeros((0,), np.int64)
        if self.nn_thresh < 0.0:
            raise ValueError('\'nn_thresh\' should be non-negative')
        descs = [require_cuda(d, "descriptors", torch.float32) for d in descriptors]
        dev = descs[0].device
        ks = np.array([d.shape[1] for d in descs], dtype=np.int64)
        for s0 in range(0, len(pairs0), super_chunk):
            pp = pairs0[s0:s0 + super_chunk]
            n0, n1 = ks[pp[:, 0]], ks[pp[:, 1]]
            cap = int(max(np.minimum(n0, n1).max(), 1))
            out = torch.empty(len(pp), 3, cap, dtype=torch.float32, device=dev)
            counts = torch.zeros(len(pp), dtype=torch.int32, device=dev)
            ws_bytes = max(int(L.sfm_simple_match_workspace_bytes(int(max(n0.max(), 1)), int(max(n1.max(), 1)))), 256)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                st = stream_ptr(dev)
                for i, (a, b) in enumerate(pp):
                    N, M = int(ks[a]), int(ks[b])
                    if N == 0 or M == 0:
                        continue
                    check(L.sfm_simple_match(ptr(descs[a]), N, N, ptr(descs[b]), M, M, float(self.nn_thresh),
                                             ptr(out[i]), cap, ptr(counts[i:i + 1]), ptr(ws), ws.numel(), st))
            lens = counts.cpu().numpy().astype(np.int64)                       # one sync per super-chunk
            mx = int(lens.max()) if len(lens) else 0
            ij = out[:, :2, :max(mx, 1)].cpu().numpy()                           # (p, 2, mx) float indices
            keep = np.arange(max(mx, 1))[None, :] < lens[:, None]
            rows = np.stack([ij[:, 0][keep], ij[:, 1][keep]], axis=1).astype(np.uint32)
            rows_out.append(rows.reshape(-1, 2))
            lens_out.append(lens)
        return np.concatenate(rows_out), np.concatenate(lens_out)

    def match_two_way(self, desc1: torch.Tensor, desc2: torch.Tensor, nn_thresh: float = None) -> np.array:
        """numpy flavour of the reference (simple_matcher.py:11-64): same matches, float64 (3,L)."""
        if desc1.shape[1] == 0 or desc2.shape[1] == 0:
            return np.zeros((3, 0))
        return self.match_two_way_cuda(desc1, desc2, nn_thresh).cpu().numpy().astype(np.float64)
