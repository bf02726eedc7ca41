"""CTA-pair (cta_group::2) GEMM check: bf16-output GEMMs against a torch reference, and timing against the
single-CTA kernel.  Run with SFM_TC_CG2=1 (pair kernel) and SFM_TC_CG2=0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simplesfm_b200 import _lib
dev = torch.device("cuda:0")
h = _lib.Handle(dev)
st = torch.cuda.current_stream().cuda_stream
print("SFM_TC_CG2 =", os.environ.get("SFM_TC_CG2"))
for (M, N, K, K1) in [(256, 256, 256, 256), (512, 512, 512, 256), (4096, 768, 256, 256), (163840, 512, 512, 256),
                      (163840, 768, 256, 256)]:
    g = torch.Generator(device="cpu").manual_seed(M + N)
    A = torch.randn(M, K, generator=g).to(dev).bfloat16()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    A1 = A[:, :K1].contiguous()
    A2 = A[:, K1:].contiguous() if K1 < K else None
    out = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    def run():
        _lib.check(_lib.lib().sfm_tc_gemm(h.h, A1.data_ptr(), None if A2 is None else A2.data_ptr(), W.data_ptr(),
                                          M, N, K, K1, bias.data_ptr(), 1.0, 0, out.data_ptr(), None, 0, st))
    run()
    torch.cuda.synchronize()
    rows = slice(0, min(M, 4096))
    ref = A[rows].float() @ W.float().t() + bias
    err = (out[rows].float() - ref).abs().max().item()
    tail = slice(max(0, M - 2048), M)
    err2 = (out[tail].float() - (A[tail].float() @ W.float().t() + bias)).abs().max().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): run()
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"M={M} N={N} K={K}: max err {err:.4f} / tail {err2:.4f}  {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")
