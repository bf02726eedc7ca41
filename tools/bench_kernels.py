"""Steady-state CUDA-event timings of the tcgen05 kernels at bench shapes (hot L2, no profiler)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simplesfm_b200 import _lib
dev = torch.device("cuda:0")
h = _lib.Handle(dev)
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
M = int(sys.argv[1]) if len(sys.argv) > 1 else 20480

def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for name, N, K, K1, f32, acc in (("qkv", 768, 256, 256, False, 0), ("merge", 256, 256, 256, False, 0),
                                 ("mlp1", 512, 512, 256, False, 0), ("mlp2", 256, 512, 512, True, 1),
                                 ("mlp2-noacc-bf16only", 256, 512, 512, False, 0)):
    A = torch.randn(M, K1, device=dev).bfloat16()
    A2 = torch.randn(M, K - K1, device=dev).bfloat16() if K1 < K else None
    W = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    oh = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    of = torch.zeros(M, N, device=dev) if f32 else None
    fn = lambda: _lib.check(L.sfm_tc_gemm(h.h, A.data_ptr(), None if A2 is None else A2.data_ptr(), W.data_ptr(), M, N, K, K1,
                                          bias.data_ptr(), 1.0, 0, oh.data_ptr(), None if of is None else of.data_ptr(), acc, st))
    us = timeit(fn)
    print(f"gemm {name:22s} M={M} N={N} K={K}: {us:8.1f} us  {2.0*M*N*K/us/1e6:8.1f} TFLOP/s")

for B, Nk in ((5, 2048), (20, 2048)):
    T = 2 * B * Nk
    qkv = torch.randn(T, 768, device=dev).bfloat16()
    out = torch.empty(T, 256, device=dev, dtype=torch.bfloat16)
    fn = lambda: _lib.check(L.sfm_tc_attention(h.h, qkv.data_ptr(), T, B, Nk, Nk, 0, B * Nk, Nk, Nk, B * Nk, 0, out.data_ptr(), st))
    us = timeit(fn, 20)
    fl = 2 * B * 4.0 * Nk * Nk * 256
    print(f"attention B={B} N={Nk} both sides: {us:8.1f} us  {fl/us/1e6:8.1f} TFLOP/s")
