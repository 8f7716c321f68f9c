"""Thin Python wrappers over the operator-level C-ABI entry points (include/videosd.h).

PyTorch is used only to own device memory; every computation is a hand-written CUDA kernel in libvideosd.so.
These wrappers exist for the parity tests and tools; the per-frame product path is engine.py.
"""
import ctypes

import torch

from ._lib import _p, check, cur_stream, lib

c_int = ctypes.c_int
c_float = ctypes.c_float


def conv_gemm(x, weight_ohwi, taps, bias=None, rowvec=None, residual=None, act=0, out_f32=False, block_n=0, splits=0,
              out=None):
    """x: bf16 NHWC tensor (nb,h,w,c) (last-dim contiguous; pixel stride taken from x.stride(2)).
    weight_ohwi: bf16 (n, taps*c). Returns (nb,h,w,n_out)."""
    assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.stride(3) == 1
    nb, h, w, c = x.shape
    ldx = x.stride(2)
    n = weight_ohwi.shape[0]
    assert weight_ohwi.dtype == torch.bfloat16 and weight_ohwi.is_contiguous() and weight_ohwi.shape[1] == taps * c
    n_out = n // 2 if act == 1 else n
    if out is None:
        out = torch.empty((nb, h, w, n_out), device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    ldo = out.stride(2)
    ldr = residual.stride(2) if residual is not None else 0
    check(lib().vsd_op_conv_gemm(_p(x), c_int(nb), c_int(h), c_int(w), c_int(c), c_int(ldx), c_int(taps),
                                 _p(weight_ohwi), c_int(n), _p(out), c_int(ldo), c_int(1 if out_f32 else 0), _p(bias),
                                 _p(rowvec), _p(residual), c_int(ldr), c_int(act), c_int(block_n), c_int(splits),
                                 cur_stream()), "vsd_op_conv_gemm")
    return out
