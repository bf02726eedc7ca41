"""ctypes binding of libsfm_b200.so (C ABI declared in include/sfm_b200.h).

There is no CPU fallback: if the shared library is missing or no B200 is visible, the
product path raises.  torch is used only as the owner of device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import threading
from typing import Dict, Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# SFM_LIB selects another build of the same library (A/B timing of kernel variants); the default is the in-tree build
LIB_PATH = os.environ.get("SFM_LIB") or os.path.join(HERE, "libsfm_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "sfm_b200.h")

BN_TRAIN, BN_EVAL = 0, 1
PREC_FP32, PREC_BF16 = 0, 1

c_float_p = C.POINTER(C.c_float)

_SIGNATURES = {
    "sfm_abi_version": (C.c_int, []),
    "sfm_last_error": (C.c_char_p, []),
    "sfm_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "sfm_launch_count": (C.c_longlong, []),
    "sfm_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "sfm_get_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]),
    "sfm_destroy": (C.c_int, [C.c_void_p]),
    "sfm_load_superpoint": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "sfm_superpoint_packed_floats": (C.c_size_t, []),
    "sfm_load_superglue": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "sfm_superglue_packed_floats": (C.c_size_t, []),
    "sfm_superpoint_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "sfm_superpoint_detect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float,
                                        C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfm_superpoint_describe": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                          C.c_void_p]),
    "sfm_superpoint_tap": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "sfm_simple_nms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sfm_sample_descriptors": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_void_p, C.c_void_p]),
    "sfm_superglue_workspace_bytes": (C.c_size_t, [C.c_int] * 5),
    "sfm_superglue_forward": (C.c_int, [C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfm_superglue_hoist_layer0": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                             C.c_void_p]),
    "sfm_superglue_forward_hoisted": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                C.c_void_p]),
    "sfm_superglue_tap": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "sfm_u8_to_unit_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfm_attention_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "sfm_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfm_sinkhorn_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "sfm_log_optimal_transport": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfm_log_sinkhorn_iterations": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfm_sinkhorn_match": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                     C.c_void_p]),
    "sfm_sinkhorn_half_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "sfm_sinkhorn_match_scaled_half": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                 C.c_void_p]),
    "sfm_sinkhorn_match_scaled": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                            C.c_void_p]),
    "sfm_tc_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "sfm_tc_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_longlong,
                                   C.c_longlong, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p,
                                   C.c_void_p]),
    "sfm_tc_conv3x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sfm_compact_matches": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfm_pack_tables": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfm_pair_list": (C.c_int, [C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfm_sparse_depth": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "sfm_simple_match_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "sfm_simple_match": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
}


class SfmError(RuntimeError):
    """A libsfm_b200 call failed (code, message from sfm_last_error)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libsfm_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None
_lock = threading.Lock()


def header_symbols() -> list:
    """Function names declared with SFM_API in include/sfm_b200.h."""
    with open(HEADER_PATH) as f:
        return re.findall(r"SFM_API\s+[\w\s\*]+?\b(sfm_\w+)\s*\(", f.read())


def lib() -> C.CDLL:
    """Load the shared library (no compute is performed); raises if it was not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f"{LIB_PATH} not found: run `python -m simplesfm_b200.build` "
                                   f"(there is no CPU fallback)")
            l = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(l, name)
                fn.restype = res
                fn.argtypes = args
            _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().sfm_last_error().decode("utf-8", "replace")
        raise SfmError(rc, msg)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not t.is_cuda:
        raise SfmError(-1, f"{name} must be a CUDA tensor (libsfm_b200 has no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class Handle:
    """One libsfm_b200 handle per device; owns the packed weights."""

    _cache: Dict[int, "Handle"] = {}

    def __init__(self, device):
        if not torch.cuda.is_available():
            raise SfmError(-2, "no CUDA device available; simplesfm_b200 has no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise SfmError(-1, f"device must be CUDA, got {device}")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().sfm_create(idx, C.byref(h)))
        self.h = h
        self._ws: Dict[str, torch.Tensor] = {}

    def workspace(self, key: str, nbytes: int) -> torch.Tensor:
        """Grow-only byte workspaces owned by torch (the C side never allocates them)."""
        t = self._ws.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = t
        return t

    def close(self):
        if self.h is not None and self.h.value:
            lib().sfm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
