"""Drop-in mirror of `simple_sfm/models/superglue.py` running on libsfm_b200 (sm_100a CUDA).

`SuperGlue(weights_path, ...)(data) -> {'matches0','matches1','matching_scores0','matching_scores1'}`
with the reference's constructor arguments (superglue.py:205-233) and forward contract
(:235-290), plus the free functions `normalize_keypoints` (:62-69), `log_optimal_transport`
(:151-171), `log_sinkhorn_iterations` (:142-148) and `arange_like` (:174-175).

Additive kwargs: `bn_mode` ('train' = batch statistics, what the reference actually executes
because it never calls `.eval()`; 'eval' = running statistics folded into the convolutions),
`precision` ('fp32' FFMA exact path / 'bf16' tcgen05 path), `device`.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch

from .. import _lib
from .._lib import Handle, check, lib, ptr, require_cuda, stream_ptr
from ..weights import load_state_dict_file, pack_superglue


def normalize_keypoints(kpts: torch.Tensor, image_shape) -> torch.Tensor:
    """ Normalize keypoints locations based on image image_shape (tiny elementwise op; inside
    `SuperGlue.forward` it is fused into the keypoint-encoder front end) """
    _, _, height, width = image_shape
    size = kpts.new_tensor([[float(width), float(height)]])
    center = size / 2
    scaling = size.max(1, keepdim=True).values * 0.7
    return (kpts - center[:, None, :]) / scaling[:, None, :]


def arange_like(x, dim: int):
    return x.new_ones(x.shape[dim]).cumsum(0) - 1


def attention(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor):
    """ softmax(q^T k / sqrt(d)) v per head (superglue.py:85-89), reference layout: query (b, 64, 4, n),
    key/value (b, 64, 4, m) CUDA f32 -> (out (b, 64, 4, n), prob (b, 4, n, m)).  This is the stand-alone fp32
    form with the probabilities materialised; `SuperGlue.forward` uses the fused kernels instead. """
    q = require_cuda(query, "query", torch.float32)
    k = require_cuda(key, "key", torch.float32)
    v = require_cuda(value, "value", torch.float32)
    b, d, h, n = q.shape
    m = k.shape[3]
    if d != 64 or h != 4 or k.shape[:3] != (b, 64, 4) or v.shape != k.shape:
        raise _lib.SfmError(-5, f"attention: expected (b,64,4,n)/(b,64,4,m) operands, got {tuple(q.shape)}, "
                                f"{tuple(k.shape)}, {tuple(v.shape)}")
    L = lib()
    out = torch.empty_like(q)
    prob = torch.empty(b, 4, n, m, dtype=torch.float32, device=q.device)
    ws = torch.empty(max(L.sfm_attention_workspace_bytes(b, n, m), 256), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        check(L.sfm_attention(ptr(q), ptr(k), ptr(v), b, n, m, ptr(out), ptr(prob), ptr(ws), ws.numel(),
                              stream_ptr(q.device)))
    return out, prob


def log_optimal_transport(scores: torch.Tensor, alpha, iters: int) -> torch.Tensor:
    """ Perform Differentiable Optimal Transport in Log-space for stability (forward only).
    scores (b, m, n) CUDA f32 -> (b, m+1, n+1). """
    s = require_cuda(scores, "scores", torch.float32)
    b, m, n = s.shape
    a = float(alpha)
    L = lib()
    Z = torch.empty(b, m + 1, n + 1, dtype=torch.float32, device=s.device)
    ws = torch.empty(max(L.sfm_sinkhorn_workspace_bytes(b, m, n), 256), dtype=torch.uint8, device=s.device)
    with torch.cuda.device(s.device):
        check(L.sfm_log_optimal_transport(ptr(s), b, m, n, a, int(iters), ptr(Z), ptr(ws), ws.numel(),
                                          stream_ptr(s.device)))
    return Z


def log_sinkhorn_iterations(Z: torch.Tensor, log_mu: torch.Tensor, log_nu: torch.Tensor, iters: int) -> torch.Tensor:
    """ Perform Sinkhorn Normalization in Log-space for stability (superglue.py:142-148): any couplings
    Z (b, m, n) and marginals log_mu (b, m), log_nu (b, n), CUDA f32 -> Z + u + v. """
    z = require_cuda(Z, "Z", torch.float32)
    b, m, n = z.shape
    mu = require_cuda(log_mu, "log_mu", torch.float32).expand(b, m).contiguous()
    nu = require_cuda(log_nu, "log_nu", torch.float32).expand(b, n).contiguous()
    out = torch.empty_like(z)
    ws = torch.empty(max(4 * b * (m + n), 256), dtype=torch.uint8, device=z.device)
    with torch.cuda.device(z.device):
        check(lib().sfm_log_sinkhorn_iterations(ptr(z), ptr(mu), ptr(nu), b, m, n, int(iters), ptr(out), ptr(ws),
                                                ws.numel(), stream_ptr(z.device)))
    return out


class SuperGlue:
    """SuperGlue feature matching middle-end (B200-native mirror of the reference class)."""

    default_config = {
        'descriptor_dim': 256,
        'weights': 'indoor',
        'keypoint_encoder': [32, 64, 128, 256],
        'GNN_layers': ['self', 'cross'] * 9,
        'sinkhorn_iterations': 100,
        'match_threshold': 0.2,
    }

    def __init__(self, weights_path, descriptor_dim=256, keypoint_encoder=[32, 64, 128, 256],
                 GNN_layers=['self', 'cross'] * 9, sinkhorn_iterations=100, match_threshold=0.2,
                 device="cuda", bn_mode="train", precision="fp32"):
        if descriptor_dim != 256 or list(keypoint_encoder) != [32, 64, 128, 256] \
                or list(GNN_layers) != ['self', 'cross'] * 9:
            raise _lib.SfmError(-5, "only the shipped SuperGlue architecture (256-d, 18 layers) is supported")
        self.descriptor_dim = descriptor_dim
        self.keypoint_encoder = keypoint_encoder
        self.GNN_layers = GNN_layers
        self.sinkhorn_iterations = int(sinkhorn_iterations)
        self.match_threshold = float(match_threshold)
        self.bn_mode = {"train": _lib.BN_TRAIN, "eval": _lib.BN_EVAL}[bn_mode]
        self.precision = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[precision]
        packed = pack_superglue(load_state_dict_file(weights_path))       # strict, like load_state_dict
        self.bin_score = float(packed[0])
        self.handle = Handle(device)
        self.device = self.handle.device
        check(lib().sfm_load_superglue(self.handle.h, packed.data_ptr(), packed.numel()))

    def cuda(self, *a, **k):
        return self

    def to(self, *a, **k):
        return self

    def eval(self):
        """nn.Module.eval(): switch BatchNorm to running statistics."""
        self.bn_mode = _lib.BN_EVAL
        return self

    def train(self, mode=True):
        self.bn_mode = _lib.BN_TRAIN if mode else _lib.BN_EVAL
        return self

    def __call__(self, data):
        return self.forward(data)

    # ------------------------------------------------------------------------------------------
    def match_records(self, desc0, kpts0, scores0, idx0, desc1, kpts1, scores1, idx1, B, K0, K1,
                      hw0, hw1, out=None, chunk_groups=1):
        """Run a batch of B pairs taken from per-image feature records (see include/sfm_b200.h).
        desc* (n,256,ldk), kpts* (n,kcap,2), scores* (n,kcap); idx* int32 [B] or None."""
        L = lib()
        dev = self.device
        nbytes = L.sfm_superglue_workspace_bytes(B, K0, K1, self.precision, chunk_groups)
        if nbytes == 0 and B > 0:
            raise _lib.SfmError(-1, L.sfm_last_error().decode())
        ws = self.handle.workspace("sg", nbytes)
        if out is None:
            out = (torch.empty(B, K0, dtype=torch.int64, device=dev), torch.empty(B, K1, dtype=torch.int64, device=dev),
                   torch.empty(B, K0, dtype=torch.float32, device=dev),
                   torch.empty(B, K1, dtype=torch.float32, device=dev))
        m0, m1, s0, s1 = out
        with torch.cuda.device(dev):
            check(L.sfm_superglue_forward(
                self.handle.h,
                ptr(desc0), ptr(kpts0), ptr(scores0), desc0.shape[-1], kpts0.shape[-2], ptr(idx0),
                ptr(desc1), ptr(kpts1), ptr(scores1), desc1.shape[-1], kpts1.shape[-2], ptr(idx1),
                B, K0, K1, hw0[0], hw0[1], hw1[0], hw1[1],
                self.sinkhorn_iterations, self.match_threshold, self.bn_mode, self.precision, chunk_groups,
                ptr(m0), ptr(m1), ptr(s0), ptr(s1), ptr(ws), ws.numel(), stream_ptr(dev)))
        return m0, m1, s0, s1, ws

    # ---- layer-0 hoisting (eval-mode BatchNorm, bf16 path) --------------------------------------
    def can_hoist(self) -> bool:
        return self.bn_mode == _lib.BN_EVAL and self.precision == _lib.PREC_BF16

    def hoist_layer0(self, desc, kpts, scores, n_img, K, hw):
        """Residual stream entering GNN layer 1 for every image of a record store: (n_img, K, 256) f32.  Layer 0 is
        a self layer, so with running BatchNorm statistics this part of the forward pass does not depend on the
        pair and is done once per image (desc (n,256,ldk), kpts (n,kcap,2), scores (n,kcap))."""
        L = lib()
        dev = self.device
        nbytes = L.sfm_superglue_workspace_bytes(n_img, K, K, self.precision, 1)
        ws = self.handle.workspace("sg", nbytes)
        store = torch.empty(n_img, K, 256, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(L.sfm_superglue_hoist_layer0(self.handle.h, ptr(desc), ptr(kpts), ptr(scores), desc.shape[-1],
                                               kpts.shape[-2], n_img, K, hw[0], hw[1], self.precision, ptr(store),
                                               ptr(ws), ws.numel(), stream_ptr(dev)))
        return store

    def match_hoisted(self, store0, idx0, store1, idx1, B, K0, K1):
        """Layers 1..17 + optimal transport + mutual matches for B pairs gathered from hoisted stores."""
        L = lib()
        dev = self.device
        nbytes = L.sfm_superglue_workspace_bytes(B, K0, K1, self.precision, 1)
        ws = self.handle.workspace("sg", nbytes)
        m0 = torch.empty(B, K0, dtype=torch.int64, device=dev)
        m1 = torch.empty(B, K1, dtype=torch.int64, device=dev)
        s0 = torch.empty(B, K0, dtype=torch.float32, device=dev)
        s1 = torch.empty(B, K1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(L.sfm_superglue_forward_hoisted(self.handle.h, ptr(store0), store0.shape[1], ptr(idx0), ptr(store1),
                                                  store1.shape[1], ptr(idx1), B, K0, K1, self.sinkhorn_iterations,
                                                  self.match_threshold, self.precision, ptr(m0), ptr(m1), ptr(s0),
                                                  ptr(s1), ptr(ws), ws.numel(), stream_ptr(dev)))
        return m0, m1, s0, s1, ws

    def tap(self, B, K0, K1, which, ws, chunk_groups=1):
        p, cnt = C.c_void_p(), C.c_size_t()
        check(lib().sfm_superglue_tap(B, K0, K1, self.precision, chunk_groups, which, ptr(ws), ws.numel(),
                                      C.byref(p), C.byref(cnt)))
        off = p.value - ws.data_ptr()
        return ws[off:off + 4 * cnt.value].view(torch.float32)

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Run SuperGlue on a pair of keypoints and descriptors"""
        desc0, desc1 = data['descriptors0'], data['descriptors1']
        kpts0, kpts1 = data['keypoints0'], data['keypoints1']

        if kpts0.shape[1] == 0 or kpts1.shape[1] == 0:  # no keypoints
            shape0, shape1 = kpts0.shape[:-1], kpts1.shape[:-1]
            return {
                'matches0': kpts0.new_full(shape0, -1, dtype=torch.int),
                'matches1': kpts1.new_full(shape1, -1, dtype=torch.int),
                'matching_scores0': kpts0.new_zeros(shape0),
                'matching_scores1': kpts1.new_zeros(shape1),
            }
        desc0 = require_cuda(desc0, "descriptors0", torch.float32)
        desc1 = require_cuda(desc1, "descriptors1", torch.float32)
        kpts0 = require_cuda(kpts0, "keypoints0", torch.float32)
        kpts1 = require_cuda(kpts1, "keypoints1", torch.float32)
        sc0 = require_cuda(data['scores0'], "scores0", torch.float32)
        sc1 = require_cuda(data['scores1'], "scores1", torch.float32)
        B, K0, K1 = kpts0.shape[0], kpts0.shape[1], kpts1.shape[1]
        hw0 = tuple(data['image0'].shape[-2:])          # only the shape is read (superglue.py:250-251)
        hw1 = tuple(data['image1'].shape[-2:])
        m0, m1, s0, s1, _ = self.match_records(desc0, kpts0, sc0, None, desc1, kpts1, sc1, None, B, K0, K1, hw0, hw1)
        return {
            'matches0': m0,  # use -1 for invalid match
            'matches1': m1,  # use -1 for invalid match
            'matching_scores0': s0,
            'matching_scores1': s1,
        }
