"""Launch one attention configuration a few times (the command ncu wraps for a source-level profile).
    python tools/ncu_one_attn.py N d heads [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosd_b200 import ops  # noqa: E402

n, d, heads = (int(v) for v in sys.argv[1:4])
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn((batch * n, heads * d), device="cuda", generator=g).bfloat16()
k = torch.randn((batch * n, heads * d), device="cuda", generator=g).bfloat16()
v = torch.randn((batch * n, heads * d), device="cuda", generator=g).bfloat16()
qp, kp = ops.pad_heads(q, heads, d), ops.pad_heads(k, heads, d)
vt = v.t().contiguous()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    ops.attention(qp, kp, vt, batch, heads, d, n, n)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    ops.attention(qp, kp, vt, batch, heads, d, n, n)
e1.record()
torch.cuda.synchronize()
print(f"attention N={n} d={d} heads={heads} batch={batch}: {e0.elapsed_time(e1) * 100:.1f} us/launch")
