"""One attention case in its own process (bring-up: a device fault is sticky for the process).
    python tools/attn_one.py batch heads d nq nk [k_slot]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosd_b200 import ops  # noqa: E402

b, h, d, nq, nk = (int(v) for v in sys.argv[1:6])
slot = int(sys.argv[6]) if len(sys.argv) > 6 else nk
g = torch.Generator(device="cuda").manual_seed(5)
q = torch.randn((b, nq, h, d), device="cuda", generator=g).bfloat16()
k = torch.randn((b, nk, h, d), device="cuda", generator=g).bfloat16()
v = torch.randn((b, nk, h, d), device="cuda", generator=g).bfloat16()
qp = ops.pad_heads(q.reshape(b * nq, h * d), h, d)
kf = torch.zeros((b, slot, h * d), device="cuda", dtype=torch.bfloat16)
kf[:, :nk] = k.reshape(b, nk, h * d)
kp = ops.pad_heads(kf.reshape(b * slot, h * d), h, d)
vf = torch.zeros((b, slot, h * d), device="cuda", dtype=torch.bfloat16)
vf[:, :nk] = v.reshape(b, nk, h * d)
vt = vf.reshape(b * slot, h * d).t().contiguous()
torch.cuda.synchronize()
o = ops.attention(qp, kp, vt, b, h, d, nq, nk, k_rows_per_img=slot, vt_cols_per_img=slot)
torch.cuda.synchronize()
qf, kf2, vf2 = q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)
ref = torch.softmax((qf @ kf2.transpose(-1, -2)) * d ** -0.5, dim=-1) @ vf2
ref = ref.transpose(1, 2).reshape(b * nq, h * d)
err = float((o.float() - ref).norm() / ref.norm())
print("CASE", sys.argv[1:], "rel", err, "OK" if err < 1.5e-2 else "BAD", flush=True)
