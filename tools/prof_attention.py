"""One attention launch at the bench shape (20 pairs x both sides x 4 heads, 2048 keypoints) for ncu:
    ncu --set full --import-source on --clock-control none -k regex:tc_attention --launch-skip 3 -c 1 -o out python tools/prof_attention.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simplesfm_b200 import _lib
dev = torch.device("cuda:0")
h = _lib.Handle(dev)
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
B, Nk = (int(sys.argv[1]) if len(sys.argv) > 1 else 20), 2048
T = 2 * B * Nk
qkv = torch.randn(T, 768, device=dev).bfloat16()
out = torch.empty(T, 256, device=dev, dtype=torch.bfloat16)
for _ in range(6):
    _lib.check(L.sfm_tc_attention(h.h, qkv.data_ptr(), T, B, Nk, Nk, 0, B * Nk, Nk, Nk, B * Nk, 0, out.data_ptr(), st))
torch.cuda.synchronize()
print("ok")
