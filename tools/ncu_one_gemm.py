"""Launch one conv-GEMM configuration a few times (the command ncu wraps for a source-level profile).
    python tools/ncu_one_gemm.py rows K N taps bn splits occ kbs"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosd_b200._lib import _p, check, cur_stream, lib  # noqa: E402

c_int = ctypes.c_int
rows, c, n, taps, bn, sp, occ, kbs = (int(v) for v in sys.argv[1:9])
h = w = 1
if taps == 9:
    h = w = int(rows ** 0.5)
else:
    w = rows
x = torch.randn((1, h, w, c), device="cuda").bfloat16()
wt = (torch.randn((n, taps * c), device="cuda") * (taps * c) ** -0.5).bfloat16()
bias = torch.zeros((n,), device="cuda")
out = torch.empty((1, h, w, n), device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    check(lib().vsd_op_conv_gemm_timed(_p(x), c_int(1), c_int(h), c_int(w), c_int(c), c_int(c), c_int(taps), _p(wt), c_int(n), _p(out),
                                       c_int(n), _p(bias), c_int(bn), c_int(sp), c_int(occ), c_int(kbs), None, cur_stream()), "gemm")
torch.cuda.synchronize()
print("done")
