"""Host side of the GPU center-crop + Lanczos resize.

The reference does this on the CPU for every frame: diffusert/videopipeline.py:92-107 (center crop to the target aspect
ratio, then `img.resize((width, height), resample=Image.Resampling.LANCZOS)`). Pillow's resize is a two-pass separable
convolution on uint8 with fixed-point coefficients; this module computes the same per-output-pixel windows and 22-bit
integer coefficients (Pillow's `precompute_coeffs` / `normalize_coeffs_8bpc`), the CUDA kernels `resample_h_kernel` /
`resample_v_kernel` (csrc/bw_kernels.cu) apply them: horizontal pass to a uint8 intermediate, then vertical, with the
same rounding (`(1 << 21) + sum` then `>> 22`, clamped). The result is bit-identical to Pillow (tests/test_host.py pins
the coefficient maths against PIL on the CPU, tests/test_gpu_ops.py the kernels).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
LANCZOS_SUPPORT = 3.0


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3.0)
    return 0.0


def lanczos_coeffs(in_size, out_size, in0=0.0, in1=None):
    """-> (bounds int32 [out][2] = (first input index, count), coeffs int32 [out][ksize], ksize)."""
    in1 = float(in_size) if in1 is None else float(in1)
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def center_crop_box(in_w, in_h, width, height):
    """The reference's crop rectangle (videopipeline.py:92-106) after PIL's int(round()) of the float box."""
    if in_w / in_h > width / height:
        new_w = in_h * (width / height)
        box = ((in_w - new_w) / 2, 0, (in_w + new_w) / 2, in_h)
    else:
        new_h = in_w * (height / width)
        box = (0, (in_h - new_h) / 2, in_w, (in_h + new_h) / 2)
    return tuple(int(round(v)) for v in box)


def resize_plan(in_w, in_h, width, height):
    """Everything the engine needs for one input geometry."""
    x0, y0, x1, y1 = center_crop_box(in_w, in_h, width, height)
    cw, ch = x1 - x0, y1 - y0
    hb, hk, hks = lanczos_coeffs(cw, width)
    vb, vk, vks = lanczos_coeffs(ch, height)
    return {"crop": (x0, y0, cw, ch), "h": (hb, hk, hks), "v": (vb, vk, vks), "identity": (cw == width and ch == height)}


def resize_reference_numpy(img_u8, width, height):
    """CPU restatement of the two passes with the tables above (used by the tests to pin them against Pillow)."""
    in_h, in_w, _ = img_u8.shape
    plan = resize_plan(in_w, in_h, width, height)
    x0, y0, cw, ch = plan["crop"]
    src = img_u8[y0:y0 + ch, x0:x0 + cw].astype(np.int64)
    (hb, hk, _), (vb, vk, _) = plan["h"], plan["v"]
    if cw != width:
        tmp = np.zeros((ch, width, 3), dtype=np.int64)
        for xx in range(width):
            xmin, n = hb[xx]
            acc = (src[:, xmin:xmin + n, :] * hk[xx, :n].astype(np.int64)[None, :, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
            tmp[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    else:
        tmp = src
    if ch != height:
        out = np.zeros((height, width, 3), dtype=np.int64)
        for yy in range(height):
            ymin, n = vb[yy]
            acc = (tmp[ymin:ymin + n] * vk[yy, :n].astype(np.int64)[:, None, None]).sum(axis=0) + (1 << (PRECISION_BITS - 1))
            out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
    else:
        out = tmp
    return out.astype(np.uint8)
