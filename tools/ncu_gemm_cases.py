"""Runs a few representative conv-GEMM shapes (for `ncu --set full -k regex:conv_gemm`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosd_b200 import ops  # noqa: E402

CASES = [  # nb, h, w, c, n, taps, block_n, splits
    (1, 16, 16, 1280, 1280, 9, 128, 6),
    (1, 64, 64, 320, 320, 9, 64, 1),
    (1, 1, 4096, 320, 320, 1, 64, 1),
    (1, 32, 32, 640, 640, 9, 128, 3),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (nb, h, w, c, n, taps, bn, sp) in CASES:
    x = torch.randn((nb, h, w, c), device="cuda").bfloat16()
    wt = (torch.randn((n, taps * c), device="cuda") * (taps * c) ** -0.5).bfloat16()
    out = torch.empty((nb, h, w, n), device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        flush.zero_()
        ops.conv_gemm(x, wt, taps, out=out, block_n=bn, splits=sp)
    torch.cuda.synchronize()
print("done")
