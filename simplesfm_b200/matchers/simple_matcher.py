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

    def match_two_way(self, desc1: torch.Tensor, desc2: torch.Tensor, nn_thresh: float = None) -> np.array:
        """numpy flavour of the reference (simple_matcher.py:11-64): same matches, float64 (3,L)."""
        if desc1.shape[1] == 0 or desc2.shape[1] == 0:
            return np.zeros((3, 0))
        return self.match_two_way_cuda(desc1, desc2, nn_thresh).cpu().numpy().astype(np.float64)
