"""Phase timestamps of one CTA of the conv-GEMM kernel for a few shapes (bring-up tool; run under gpurun)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosd_b200._lib import _p, check, cur_stream, lib  # noqa: E402

c_int = ctypes.c_int
CASES = []  # nb,h,w,c,n,taps,bn,splits,occ,kbs
for kbs in (1, 2, 4):
    CASES += [
        (1, 64, 64, 320, 320, 9, 64, 1, 1, kbs),
        (1, 1, 128 * 148, 2048, 128, 1, 128, 1, 1, kbs),
    ]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = ["setup", "first_operands", "mainloop_issue", "accum_ready", "epilogue", "teardown"]
for (nb, h, w, c, n, taps, bn, sp, occ, kbs) in CASES:
    x = torch.randn((nb, h, w, c), device="cuda").bfloat16()
    wt = (torch.randn((n, taps * c), device="cuda") * (taps * c) ** -0.5).bfloat16()
    bias = torch.randn((n,), device="cuda")
    out = torch.empty((nb, h, w, n), device="cuda", dtype=torch.bfloat16)
    dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
    for cold in (True,):
        for _ in range(2):
            if cold:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(lib().vsd_op_conv_gemm_timed(_p(x), c_int(nb), c_int(h), c_int(w), c_int(c), c_int(x.stride(2)), c_int(taps),
                                               _p(wt), c_int(n), _p(out), c_int(out.stride(2)), _p(bias), c_int(bn), c_int(sp),
                                               c_int(occ), c_int(kbs), _p(dbg), cur_stream()), "timed")
            e1.record()
            torch.cuda.synchronize()
        d = dbg.cpu().tolist()
        ph = [d[i + 1] - d[i] for i in range(6)]
        kb = taps * c // 64 // sp
        print(f"shape {(nb,h,w,c,n,taps)} bn={bn} sp={sp} occ={occ} kbs={kbs} {'cold' if cold else 'warm'} kernel {e0.elapsed_time(e1)*1e3:7.1f} us | "
              f"CTA0 total {d[6]-d[0]:6d} cyc | " + " ".join(f"{nm}={v}" for nm, v in zip(names, ph)) + f" | kblocks/CTA {kb} -> {ph[2]/max(kb,1):.0f} cyc/kb",
              flush=True)
        t0 = d[1]
        print("   producer (empty-wait done, tma issued):", [(d[16 + 2 * i] - t0, d[17 + 2 * i] - t0) for i in range(12)])
        print("   mma      (full-wait done, commit issued):", [(d[64 + 2 * i] - t0, d[65 + 2 * i] - t0) for i in range(12)], flush=True)
