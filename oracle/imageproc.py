"""Oracle image I/O (TEST INFRASTRUCTURE ONLY).

* preprocess / postprocess restate diffusers' VaeImageProcessor as called at
  diffusert/lcm/lcm_controlnet.py:457 and :609-611 [diffusers-knowledge].
* yuv420_to_rgb / rgb_to_yuv420 DEFINE the colour conversion that the reference gets from PyAV/libswscale at
  diffusert/server.py:108 (frame.to_image()) and :117 (VideoFrame.from_image + encoder reformat). libswscale's
  bits depend on its build and SIMD path and it is not available here, so the spec is stated (Appendix D.3):
  ITU-R BT.601 limited range, 8-bit integer matrices with +128 rounding and >>8, nearest (2x2 replicate) chroma on
  the way in, rounded 2x2 box average of RGB on the way out, even width and height. The CUDA kernels must match
  these functions bit-exactly.
"""
import numpy as np
import torch


def yuv420_to_rgb(y, u, v):
    """y: (H,W) u8; u, v: (H/2, W/2) u8 -> (H,W,3) u8."""
    y = np.asarray(y)
    H, W = y.shape
    assert H % 2 == 0 and W % 2 == 0 and u.shape == (H // 2, W // 2) and v.shape == u.shape
    c = y.astype(np.int32) - 16
    d = np.repeat(np.repeat(np.asarray(u).astype(np.int32), 2, axis=0), 2, axis=1) - 128
    e = np.repeat(np.repeat(np.asarray(v).astype(np.int32), 2, axis=0), 2, axis=1) - 128
    r = (298 * c + 409 * e + 128) >> 8
    g = (298 * c - 100 * d - 208 * e + 128) >> 8
    b = (298 * c + 516 * d + 128) >> 8
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def rgb_to_yuv420(rgb):
    """rgb: (H,W,3) u8 -> y (H,W), u (H/2,W/2), v (H/2,W/2) u8."""
    rgb = np.asarray(rgb).astype(np.int32)
    H, W, _ = rgb.shape
    assert H % 2 == 0 and W % 2 == 0
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    y = ((66 * r + 129 * g + 25 * b + 128) >> 8) + 16
    # chroma from the rounded 2x2 mean of RGB
    m = (rgb.reshape(H // 2, 2, W // 2, 2, 3).sum(axis=(1, 3)) + 2) >> 2
    mr, mg, mb = m[..., 0], m[..., 1], m[..., 2]
    u = ((-38 * mr - 74 * mg + 112 * mb + 128) >> 8) + 128
    v = ((112 * mr - 94 * mg - 18 * mb + 128) >> 8) + 128
    return (np.clip(y, 0, 255).astype(np.uint8), np.clip(u, 0, 255).astype(np.uint8),
            np.clip(v, 0, 255).astype(np.uint8))


def preprocess(rgb_u8):
    """(B,H,W,3) or (H,W,3) u8 -> (B,3,H,W) fp32 in [-1,1]: astype(f32)/255 -> NCHW -> 2x-1."""
    a = np.asarray(rgb_u8)
    if a.ndim == 3:
        a = a[None]
    x = a.astype(np.float32) / 255.0
    x = torch.from_numpy(x.transpose(0, 3, 1, 2).copy())
    return 2.0 * x - 1.0


def postprocess(image):
    """(B,3,H,W) float in [-1,1] -> (B,H,W,3) u8: (x/2+0.5).clamp(0,1) -> NHWC -> (x*255).round() (half-even)."""
    x = (image / 2 + 0.5).clamp(0, 1)
    x = x.cpu().permute(0, 2, 3, 1).float().numpy()
    return (x * 255).round().astype("uint8")


def synthetic_frame(height, width, seed=0, shift=0):
    """Deterministic smooth test image + noise as limited-range YUV420P planes."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = np.meshgrid(np.arange(height, dtype=np.float32), np.arange(width, dtype=np.float32) + shift,
                         indexing="ij")
    base = np.stack([
        0.5 + 0.5 * np.sin(xx / 37.0) * np.cos(yy / 53.0),
        0.5 + 0.5 * np.sin((xx + yy) / 71.0),
        0.5 + 0.5 * np.cos(xx / 29.0 - yy / 41.0),
    ], axis=-1)
    noise = torch.rand((height, width, 3), generator=g).numpy() * 0.1
    rgb = np.clip((base * 0.9 + noise) * 255.0, 0, 255).astype(np.uint8)
    return rgb_to_yuv420(rgb)
