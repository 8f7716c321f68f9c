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
              out=None, halo=False, pair=False, persist=False, stride2=False, pad=1, cluster=None):
    """x: bf16 NHWC tensor (nb,h,w,c) (last-dim contiguous; pixel stride taken from x.stride(2)).
    weight_ohwi: bf16 (n, taps*c). Returns (nb,h,w,n_out). cluster: True / False forces / forbids the in-cluster split-K
    reduction (DSMEM exchange) when splits > 1; None = the library's default."""
    assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.stride(3) == 1
    nb, h, w, c = x.shape
    ldx = x.stride(2)
    n = weight_ohwi.shape[0]
    assert weight_ohwi.dtype == torch.bfloat16 and weight_ohwi.is_contiguous() and weight_ohwi.shape[1] == taps * c
    n_out = n // 2 if act == 1 else n
    if out is None:
        ho, wo = ((h + 2 * pad - 3 + (0 if pad else 1)) // 2 + 1, (w + 2 * pad - 3 + (0 if pad else 1)) // 2 + 1) if stride2 else (h, w)
        out = torch.empty((nb, ho, wo, n_out), device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    ldo = out.stride(2)
    ldr = residual.stride(2) if residual is not None else 0
    check(lib().vsd_op_conv_gemm(_p(x), c_int(nb), c_int(h), c_int(w), c_int(c), c_int(ldx), c_int(taps),
                                 _p(weight_ohwi), c_int(n), _p(out), c_int(ldo), c_int(1 if out_f32 else 0), _p(bias),
                                 _p(rowvec), _p(residual), c_int(ldr), c_int(act | (256 if halo else 0) | (512 if pair else 0) | (1024 if persist else 0) | ((2048 | (0 if pad else 4096)) if stride2 else 0) | (0 if cluster is None else (8192 if cluster else 16384))), c_int(block_n), c_int(splits),
                                 cur_stream()), "vsd_op_conv_gemm")
    return out


def gemm_candidates(x, weight_ohwi, taps, out, bias=None, rowvec=None, residual=None, act=0, stride2=False, pad=1):
    """The (block_n, splits, occ, kb_per_stage, mode) configurations the engine's autotuner may pick for this GEMM."""
    nb, h, w, c = x.shape
    n = weight_ohwi.shape[0]
    buf = (ctypes.c_int * (5 * 4096))()
    k = lib().vsd_op_gemm_candidates(_p(x), c_int(nb), c_int(h), c_int(w), c_int(c), c_int(x.stride(2)), c_int(taps),
                                     c_int(1 if stride2 else 0), c_int(pad), _p(weight_ohwi), c_int(n), _p(out), c_int(out.stride(2)),
                                     c_int(1 if out.dtype == torch.float32 else 0), _p(bias), _p(rowvec), _p(residual),
                                     c_int(residual.stride(2) if residual is not None else 0), c_int(act), buf, c_int(4096))
    if k < 0:
        check(k, "vsd_op_gemm_candidates")
    return [tuple(buf[i * 5 + j] for j in range(5)) for i in range(k)]


def conv_gemm_cfg(x, weight_ohwi, taps, out, cfg, bias=None, rowvec=None, residual=None, act=0, stride2=False, pad=1):
    """conv_gemm with one explicit configuration (a tuple from gemm_candidates) into `out`."""
    nb, h, w, c = x.shape
    n = weight_ohwi.shape[0]
    bn, sp, occ, kbs, mode = cfg
    check(lib().vsd_op_conv_gemm_cfg(_p(x), c_int(nb), c_int(h), c_int(w), c_int(c), c_int(x.stride(2)), c_int(taps),
                                     c_int(1 if stride2 else 0), c_int(pad), _p(weight_ohwi), c_int(n), _p(out), c_int(out.stride(2)),
                                     c_int(1 if out.dtype == torch.float32 else 0), _p(bias), _p(rowvec), _p(residual),
                                     c_int(residual.stride(2) if residual is not None else 0), c_int(act), c_int(bn), c_int(sp),
                                     c_int(occ), c_int(kbs), c_int(mode), cur_stream()), "vsd_op_conv_gemm_cfg")
    return out


def linear_ln(x, w_raw, gamma, beta, bias=None, eps=1e-5, swapped=False, act=0, block_n=0, stats=None, pair=False):
    """Linear(LayerNorm(x)) with the LayerNorm folded into the GEMM. x: bf16 (rows, c); w_raw: bf16 (n, c); stats: fp32
    (rows, nst, 2) from linear_stats (None: computed by a helper kernel).
    Returns (rows, n) -- (rows, n/2) for act=1 (GEGLU) -- or, swapped, (n, rows8) with rows padded to a multiple of 8."""
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1 and w_raw.dtype == torch.bfloat16 and w_raw.is_contiguous()
    rows, c = x.shape
    n = w_raw.shape[0]
    if swapped:
        out = torch.zeros((n, (rows + 7) // 8 * 8), device=x.device, dtype=torch.bfloat16)
    else:
        out = torch.empty((rows, n // 2 if act == 1 else n), device=x.device, dtype=torch.bfloat16)
    nst = 0 if stats is None else stats.shape[1]
    check(lib().vsd_op_linear_ln(_p(x), c_int(rows), c_int(c), c_int(x.stride(0)), _p(w_raw), c_int(n), _p(gamma), _p(beta), _p(bias),
                                 ctypes.c_float(eps), _p(out), c_int(out.stride(0)), c_int(1 if swapped else 0), c_int(act | (256 if pair else 0)),
                                 c_int(block_n), _p(stats), c_int(nst), cur_stream()), "vsd_op_linear_ln")
    return out


def linear_stats(x, w, bias=None, residual=None, block_n=0, splits=1, pair=False):
    """x W^T + bias (+ residual) -> (out bf16 (rows, n), stats fp32 (rows, n_tiles, 2)): the partial row sums a LayerNorm folded
    into the next GEMM consumes."""
    rows, c = x.shape
    n = w.shape[0]
    out = torch.empty((rows, n), device=x.device, dtype=torch.bfloat16)
    stats = torch.zeros((rows, (n + 31) // 32, 2), device=x.device, dtype=torch.float32)
    nt = lib().vsd_op_linear_stats(_p(x), c_int(rows), c_int(c), c_int(x.stride(0)), _p(w), c_int(n), _p(bias), _p(residual),
                                   c_int(residual.stride(0) if residual is not None else 0), _p(out), c_int(out.stride(0)), _p(stats),
                                   c_int(block_n), c_int(splits), c_int(1 if pair else 0), cur_stream())
    if nt <= 0:
        check(nt if nt else -1, "vsd_op_linear_stats")
    torch.cuda.synchronize()
    return out, stats.view(-1)[: rows * nt * 2].view(rows, nt, 2).contiguous()


def attn_dk_pad(d):
    return (d + 63) // 64 * 64


def pad_heads(x, heads, d):
    """(rows, heads*d) -> (rows, heads*dk_pad) with per-head zero padding (what the padded Q/K projections emit)."""
    rows = x.shape[0]
    dk = attn_dk_pad(d)
    out = torch.zeros((rows, heads, dk), device=x.device, dtype=x.dtype)
    out[:, :, :d] = x.view(rows, heads, d)
    return out.view(rows, heads * dk)


def attention(q_pad, k_pad, vt, batch, heads, d, nq, nk, q_rows_per_img=None, k_rows_per_img=None,
              vt_cols_per_img=None, out=None):
    """q_pad/k_pad: bf16 (batch*rows, heads*dk_pad); vt: bf16 (heads*d, batch*cols). Returns (batch*nq, heads*d)."""
    q_rows_per_img = nq if q_rows_per_img is None else q_rows_per_img
    k_rows_per_img = nk if k_rows_per_img is None else k_rows_per_img
    vt_cols_per_img = nk if vt_cols_per_img is None else vt_cols_per_img
    if out is None:
        out = torch.empty((batch * nq, heads * d), device=q_pad.device, dtype=torch.bfloat16)
    check(lib().vsd_op_attention(_p(q_pad), c_int(q_pad.stride(0)), _p(k_pad), c_int(k_pad.stride(0)), _p(vt),
                                 c_int(vt.stride(0)), _p(out), c_int(out.stride(0)), c_int(batch), c_int(heads),
                                 c_int(d), c_int(nq), c_int(nk), c_int(q_rows_per_img), c_int(k_rows_per_img),
                                 c_int(vt_cols_per_img), c_int(vt.shape[0]), cur_stream()), "vsd_op_attention")
    return out


def groupnorm(x, gamma, beta, groups=32, eps=1e-5, silu=False, out=None):
    """x: bf16 NHWC (nb,h,w,c)."""
    nb, h, w, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    check(lib().vsd_op_groupnorm(_p(x), c_int(x.stride(2)), _p(out), c_int(out.stride(2)), _p(gamma), _p(beta),
                                 c_int(nb), c_int(h * w), c_int(c), c_int(groups), c_float(eps), c_int(1 if silu else 0),
                                 cur_stream()), "vsd_op_groupnorm")
    return out


def layernorm(x, gamma, beta, eps=1e-5):
    """x: bf16 (rows, c)."""
    rows, c = x.shape
    out = torch.empty_like(x)
    check(lib().vsd_op_layernorm(_p(x), c_int(x.stride(0)), _p(out), c_int(out.stride(0)), _p(gamma), _p(beta),
                                 c_int(rows), c_int(c), c_float(eps), cur_stream()), "vsd_op_layernorm")
    return out


def upsample_nearest(x, ho, wo):
    nb, hi, wi, c = x.shape
    out = torch.empty((nb, ho, wo, c), device=x.device, dtype=x.dtype)
    check(lib().vsd_op_upsample_nearest(_p(x), c_int(x.stride(2)), _p(out), c_int(out.stride(2)), c_int(nb), c_int(hi),
                                        c_int(wi), c_int(ho), c_int(wo), c_int(c), cur_stream()), "vsd_op_upsample_nearest")
    return out


def im2col_s2(x):
    nb, hi, wi, c = x.shape
    ho, wo = (hi - 1) // 2 + 1, (wi - 1) // 2 + 1
    out = torch.empty((nb, ho, wo, 9 * c), device=x.device, dtype=x.dtype)
    check(lib().vsd_op_im2col_s2(_p(x), c_int(x.stride(2)), _p(out), c_int(nb), c_int(hi), c_int(wi), c_int(c), c_int(ho),
                                 c_int(wo), cur_stream()), "vsd_op_im2col_s2")
    return out


def conv3x3_small_cin(x, x_kind, weight_ohwi_f32, bias, relu=False):
    """x: (nb,h,w,cin) fp32 (kind 0/2) or u8 (kind 1). weight: fp32 (cout,3,3,cin)."""
    nb, h, w, cin = x.shape
    cout = weight_ohwi_f32.shape[0]
    out = torch.empty((nb, h, w, cout), device=x.device, dtype=torch.bfloat16)
    check(lib().vsd_op_conv3x3_small_cin(_p(x), c_int(x_kind), c_int(nb), c_int(h), c_int(w), c_int(cin),
                                         _p(weight_ohwi_f32), _p(bias), _p(out), c_int(cout), c_int(cout),
                                         c_int(1 if relu else 0), cur_stream()), "vsd_op_conv3x3_small_cin")
    return out


def add_noise(x0, noise, sqrt_alpha, sqrt_one_minus_alpha):
    out = torch.empty_like(x0)
    check(lib().vsd_op_add_noise(_p(x0), _p(noise), _p(out), c_float(sqrt_alpha), c_float(sqrt_one_minus_alpha),
                                 ctypes.c_long(x0.numel()), cur_stream()), "vsd_op_add_noise")
    return out


def lcm_step(eps, x, z, sc):
    """sc: dict with sqrt_alpha, sqrt_beta, c_skip, c_out, sqrt_alpha_prev, sqrt_beta_prev (python floats)."""
    x_prev = torch.empty_like(x)
    den = torch.empty_like(x)
    check(lib().vsd_op_lcm_step(_p(eps), _p(x), _p(z), _p(x_prev), _p(den), c_float(sc["sqrt_alpha"]),
                                c_float(sc["sqrt_beta"]), c_float(sc["c_skip"]), c_float(sc["c_out"]),
                                c_float(sc["sqrt_alpha_prev"]), c_float(sc["sqrt_beta_prev"]),
                                c_int(0 if z is None else 1), ctypes.c_long(x.numel()), cur_stream()), "vsd_op_lcm_step")
    return x_prev, den


def yuv420_to_rgb(y, u, v):
    """y: u8 (nb,h,w); u,v: u8 (nb,h/2,w/2) -> (nb,h,w,3) u8."""
    nb, h, w = y.shape
    rgb = torch.empty((nb, h, w, 3), device=y.device, dtype=torch.uint8)
    check(lib().vsd_op_yuv420_to_rgb(_p(y), _p(u), _p(v), _p(rgb), c_int(nb), c_int(h), c_int(w), cur_stream()),
          "vsd_op_yuv420_to_rgb")
    return rgb


def pack_rgb_yuv420(img, taesd_denorm=False, want_rgb=True):
    """img: fp32 (nb,h,w,ld>=3) -> rgb (nb,h,w,3) u8, y, u, v planes."""
    nb, h, w, ld = img.shape
    dev = img.device
    rgb = torch.empty((nb, h, w, 3), device=dev, dtype=torch.uint8) if want_rgb else None
    y = torch.empty((nb, h, w), device=dev, dtype=torch.uint8)
    u = torch.empty((nb, h // 2, w // 2), device=dev, dtype=torch.uint8)
    v = torch.empty((nb, h // 2, w // 2), device=dev, dtype=torch.uint8)
    check(lib().vsd_op_pack_rgb_yuv420(_p(img), c_int(ld), _p(rgb), _p(y), _p(u), _p(v), c_int(nb), c_int(h), c_int(w),
                                       c_int(1 if taesd_denorm else 0), cur_stream()), "vsd_op_pack_rgb_yuv420")
    return rgb, y, u, v


def sobel_control(rgb_u8, low=0.11, high=0.8):
    """rgb_u8: u8 (nb,h,w,3) -> control image fp32 (nb,h,w,3) in [0,1]."""
    nb, h, w, _ = rgb_u8.shape
    dev = rgb_u8.device
    mag = torch.empty((nb, h, w), device=dev, dtype=torch.float32)
    mx = torch.zeros((nb,), device=dev, dtype=torch.int32)
    ctl = torch.empty((nb, h, w, 3), device=dev, dtype=torch.float32)
    check(lib().vsd_op_sobel_control(_p(rgb_u8), _p(mag), _p(mx), _p(ctl), c_int(nb), c_int(h), c_int(w), c_float(low),
                                     c_float(high), cur_stream()), "vsd_op_sobel_control")
    return ctl


def crop_resize(rgb_u8, width, height):
    """rgb_u8: u8 (nb,in_h,in_w,3) on the device -> center-cropped, Lanczos-resized u8 (nb,height,width,3)."""
    from . import resample
    nb, in_h, in_w, _ = rgb_u8.shape
    dev = rgb_u8.device
    plan = resample.resize_plan(in_w, in_h, width, height)
    x0, y0, cw, ch = plan["crop"]
    (hb, hk, hks), (vb, vk, vks) = plan["h"], plan["v"]
    t = lambda a: torch.from_numpy(a.astype("int32")).contiguous().to(dev)  # noqa: E731
    hb, hk, vb, vk = t(hb), t(hk), t(vb), t(vk)
    tmp = torch.empty((nb, ch, width, 3), device=dev, dtype=torch.uint8)
    out = torch.empty((nb, height, width, 3), device=dev, dtype=torch.uint8)
    check(lib().vsd_op_crop_resize(_p(rgb_u8), c_int(in_w), c_int(in_h), c_int(x0), c_int(y0), c_int(cw), c_int(ch), _p(tmp), _p(out),
                                   c_int(width), c_int(height), _p(hb), _p(hk), c_int(hks), _p(vb), _p(vk), c_int(vks), c_int(nb),
                                   cur_stream()), "vsd_op_crop_resize")
    return out


def conv3x3_direct(x, weight_ohwi, bias, stride=1, silu=True):
    """x: bf16 (nb,h,w,cin); weight: bf16 (cout, 9*cin)."""
    nb, h, w, cin = x.shape
    cout = weight_ohwi.shape[0]
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    y = torch.empty((nb, ho, wo, cout), device=x.device, dtype=torch.bfloat16)
    check(lib().vsd_op_conv3x3_direct(_p(x), c_int(x.stride(2)), c_int(nb), c_int(h), c_int(w), c_int(cin), _p(weight_ohwi),
                                      _p(bias), _p(y), c_int(cout), c_int(cout), c_int(stride), c_int(1 if silu else 0),
                                      cur_stream()), "vsd_op_conv3x3_direct")
    return y
