"""Time line of one CTA of the attention kernel (debug build with -DSFM_ATT_TRACE):
    SFM_EXTRA_NVCC_FLAGS=-DSFM_ATT_TRACE python -m simplesfm_b200.build --force; cp simplesfm_b200/libsfm_b200.so /tmp/libtrace.so
    python -m simplesfm_b200.build --force;  SFM_LIB=/tmp/libtrace.so python tools/att_trace.py
Prints, per K/V step, clock offsets of: each softmax warp's loop top / S ready / s_free arrive / P V wait / p_full
arrive, and each issuer's S (g+1) and P V (g) issue."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from simplesfm_b200 import _lib
dev = torch.device("cuda:0")
h = _lib.Handle(dev)
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
B, Nk = 20, 2048
T = 2 * B * Nk
qkv = torch.randn(T, 768, device=dev).bfloat16()
out = torch.empty(T, 256, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    _lib.check(L.sfm_tc_attention(h.h, qkv.data_ptr(), T, B, Nk, Nk, 0, B * Nk, Nk, Nk, B * Nk, 0, out.data_ptr(), st))
torch.cuda.synchronize()
NG = 24
buf = np.zeros(NG * 64, np.int64)
L.sfm_debug_att_trace.argtypes = [C.c_void_p, C.c_int]
assert L.sfm_debug_att_trace(buf.ctypes.data, buf.size) == 0
tr = buf.reshape(NG, 64)
t0 = tr[tr > 0].min()
names = ["top", "Srdy", "sfree", "pvw0", "pvw1", "pfull"]
for g in range(NG):
    r = tr[g]
    base = r[r > 0].min() if (r > 0).any() else 0
    print(f"--- step {64 + g}: starts at +{base - t0}")
    for t in range(2):
        ws = [t * 4 + k for k in range(4)]
        line = f"  tile {t}: "
        for ev, nm in enumerate(names):
            vals = [int(r[w * 6 + ev] - base) if r[w * 6 + ev] else -1 for w in ws]
            line += f"{nm} {vals}  "
        print(line)
        i = 48 + t * 6
        print(f"          issuer: top {r[i] - base}, kv_full(g+1) {r[i + 1] - base}, S(g+1) issue {r[i + 2] - base}, PV(g) issue {r[i + 3] - base}, PV issued {r[i + 4] - base}, committed {r[i + 5] - base}")
