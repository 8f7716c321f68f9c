"""Per-kernel cost of dependent kernels inside one CUDA graph (PDL on/off via VSD_PDL): chains of tiny LayerNorms,
small GEMMs, and a GEMM->LN alternation. Bring-up tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosd_b200 import ops  # noqa: E402

dev = "cuda"
x = torch.randn((256, 1280), device=dev).bfloat16()
g = torch.ones(1280, device=dev); b = torch.zeros(1280, device=dev)
w = (torch.randn((1280, 1280), device=dev) * 0.03).bfloat16()
x4 = x.view(1, 1, 256, 1280)
out = torch.empty((1, 1, 256, 1280), device=dev, dtype=torch.bfloat16)
xs = torch.randn((4096, 320), device=dev).bfloat16()
gs = torch.ones(320, device=dev); bs = torch.zeros(320, device=dev)


def chain(fn, n=200):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr, stream=s):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * n)


print("PDL =", os.environ.get("VSD_PDL", "1"))
print(f"layernorm 256x1280 chain      : {chain(lambda: ops.layernorm(x, g, b)):.2f} us/kernel")
print(f"layernorm 4096x320 chain      : {chain(lambda: ops.layernorm(xs, gs, bs)):.2f} us/kernel")
print(f"gemm 256x1280x1280 bn128 chain: {chain(lambda: ops.conv_gemm(x4, w, 1, out=out, block_n=128, splits=1)):.2f} us/kernel")
print(f"gemm 256x1280x1280 bn64 chain : {chain(lambda: ops.conv_gemm(x4, w, 1, out=out, block_n=64, splits=1)):.2f} us/kernel")
print(f"gemm 256x1280x1280 bn32 chain : {chain(lambda: ops.conv_gemm(x4, w, 1, out=out, block_n=32, splits=1)):.2f} us/kernel")


def alt():
    ops.conv_gemm(x4, w, 1, out=out, block_n=64, splits=1)
    ops.layernorm(x, g, b)


print(f"gemm+ln alternating           : {chain(alt, 100)/2:.2f} us/kernel")
y = torch.empty_like(x)
print(f"torch add chain (reference)   : {chain(lambda: torch.add(x, 1, out=y)):.2f} us/kernel")
