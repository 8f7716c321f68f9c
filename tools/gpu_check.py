"""Bring-up battery for the CUDA kernels: runs many small op checks against torch fp32 on the GPU, never stops at
the first failure, and writes a JSON report to gpurun_out/check_<tag>.json. Run on a B200 via gpurun.

    python tools/gpu_check.py [gemm] [attn] [norm] ...
"""
import json
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosd_b200 import ops  # noqa: E402
from videosd_b200._lib import check_fault  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
DEV = "cuda"
RESULTS = []


def rel_err(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def record(name, err, tol, extra=None):
    ok = bool(err == err and err <= tol)
    RESULTS.append({"name": name, "err": err, "tol": tol, "ok": ok, "extra": extra})
    print(("PASS" if ok else "FAIL"), name, f"err={err:.3e} tol={tol:.1e}", extra or "", flush=True)


def run(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
        check_fault()
    except Exception as e:  # noqa: BLE001
        RESULTS.append({"name": name, "ok": False, "exc": repr(e)})
        print("EXC ", name, repr(e), flush=True)
        traceback.print_exc()


def g(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def randn(shape, seed, scale=1.0):
    return (torch.randn(shape, generator=g(seed)) * scale).to(DEV)


# ------------------------------------------------------------------------------------------------ GEMM / conv
def ref_conv(x_nhwc, w_ohwi, taps, bias=None, rowvec=None, residual=None):
    nb, h, w, c = x_nhwc.shape
    n = w_ohwi.shape[0]
    xf = x_nhwc.float().permute(0, 3, 1, 2)
    if taps == 9:
        wf = w_ohwi.float().view(n, 3, 3, c).permute(0, 3, 1, 2)
        y = torch.nn.functional.conv2d(xf, wf, padding=1)
    else:
        y = torch.nn.functional.conv2d(xf, w_ohwi.float().view(n, c, 1, 1))
    y = y.permute(0, 2, 3, 1)
    if bias is not None:
        y = y + bias.view(1, 1, 1, n)
    if rowvec is not None:
        y = y + rowvec.view(nb, 1, 1, n)
    if residual is not None:
        y = y + residual.float()
    return y


def gemm_case(name, nb, h, w, c, n, taps, bias=False, rowvec=False, residual=False, out_f32=False, block_n=0, splits=0,
              tol=1e-2):
    def fn():
        x = randn((nb, h, w, c), 1).bfloat16()
        wt = randn((n, taps * c), 2, scale=(taps * c) ** -0.5).bfloat16()
        b = randn((n,), 3) if bias else None
        rv = randn((nb, n), 4) if rowvec else None
        res = randn((nb, h, w, n), 5).bfloat16() if residual else None
        y = ops.conv_gemm(x, wt, taps, bias=b, rowvec=rv, residual=res, out_f32=out_f32, block_n=block_n, splits=splits)
        torch.cuda.synchronize()
        ref = ref_conv(x, wt, taps, b, rv, res)
        record(name, rel_err(y, ref), tol, {"shape": [nb, h, w, c, n, taps], "block_n": block_n, "splits": splits})

    run(name, fn)


def check_gemm():
    # smallest possible: one tile, one k-block
    gemm_case("lin_128x64x32", 1, 1, 128, 64, 32, 1, block_n=32)
    gemm_case("lin_128x64x64", 1, 1, 128, 64, 64, 1, block_n=64)
    gemm_case("lin_128x128x128_k2", 1, 1, 128, 128, 128, 1, block_n=128)
    gemm_case("lin_256x320x320_auto", 1, 1, 256, 320, 320, 1)
    gemm_case("lin_4096x320x320_bn160", 1, 1, 4096, 320, 320, 1, block_n=160)
    gemm_case("lin_4096x320x320_bn256", 1, 1, 4096, 320, 320, 1, block_n=256)
    gemm_case("lin_4096x1280x320_bias_res", 1, 1, 4096, 1280, 320, 1, bias=True, residual=True)
    gemm_case("lin_77x768x320", 1, 1, 77, 768, 320, 1)
    gemm_case("lin_fp32out_n4", 1, 1, 4096, 320, 4, 1, bias=True, out_f32=True)
    gemm_case("conv1x1_64x64_320_640", 1, 64, 64, 320, 640, 1, bias=True)
    gemm_case("conv3x3_64x64_64_64", 1, 64, 64, 64, 64, 9, bias=True)
    gemm_case("conv3x3_64x64_320_320_all", 1, 64, 64, 320, 320, 9, bias=True, rowvec=True, residual=True)
    gemm_case("conv3x3_32x32_640_640", 1, 32, 32, 640, 640, 9, bias=True)
    gemm_case("conv3x3_16x16_1280_1280_splitauto", 1, 16, 16, 1280, 1280, 9, bias=True, rowvec=True)
    gemm_case("conv3x3_8x8_1280_1280_split4", 1, 8, 8, 1280, 1280, 9, bias=True, residual=True, splits=4)
    gemm_case("conv3x3_b2_8x8_1280", 2, 8, 8, 1280, 1280, 9, bias=True)
    gemm_case("conv3x3_b3_16x16_64", 3, 16, 16, 64, 64, 9, bias=True)
    gemm_case("conv3x3_odd_45x80_64", 1, 45, 80, 64, 64, 9, bias=True)
    gemm_case("conv3x3_odd_23x40_128_n3", 1, 23, 40, 128, 3, 9, bias=True, out_f32=True)
    gemm_case("conv3x3_512x512_64_64", 1, 512, 512, 64, 64, 9, bias=True)
    gemm_case("conv3x3_96x96_b4_320", 4, 96, 96, 320, 320, 9, bias=True)

    def geglu():
        m, c = 4096, 320
        x = randn((1, 1, m, c), 11).bfloat16()
        w_full = randn((8 * c, c), 12, scale=c ** -0.5)
        b_full = randn((8 * c,), 13)
        # interleave per 128-row tile: 64 value rows then the matching 64 gate rows
        half = 4 * c
        idx = []
        for t in range(half // 64):
            idx += list(range(t * 64, t * 64 + 64)) + list(range(half + t * 64, half + t * 64 + 64))
        idx = torch.tensor(idx, device=DEV)
        w_il = w_full[idx].bfloat16().contiguous()
        b_il = b_full[idx].contiguous()
        y = ops.conv_gemm(x, w_il, 1, bias=b_il, act=1, block_n=128)
        torch.cuda.synchronize()
        hfull = x.float().view(m, c) @ w_full.bfloat16().float().t() + b_full
        ref = hfull[:, :half] * torch.nn.functional.gelu(hfull[:, half:])
        record("geglu_4096x320", rel_err(y.view(m, half), ref), 1e-2)

    run("geglu_4096x320", geglu)


def bench_gemm():
    """Rough timings (CUDA events) of representative layers; not a benchmark of record."""
    cases = [
        ("lin 4096x320x320", 1, 1, 4096, 320, 320, 1),
        ("lin 4096x2560(geglu-in as plain)x320", 1, 1, 4096, 320, 2560, 1),
        ("lin 4096x320x1280", 1, 1, 4096, 1280, 320, 1),
        ("conv3 64x64 320->320", 1, 64, 64, 320, 320, 9),
        ("conv3 32x32 640->640", 1, 32, 32, 640, 640, 9),
        ("conv3 16x16 1280->1280", 1, 16, 16, 1280, 1280, 9),
        ("conv3 8x8 1280->1280", 1, 8, 8, 1280, 1280, 9),
        ("conv3 16x16 2560->1280", 1, 16, 16, 2560, 1280, 9),
        ("conv3 512x512 64->64", 1, 512, 512, 64, 64, 9),
        ("conv3 96x96x4 320->320", 4, 96, 96, 320, 320, 9),
        ("lin 8192x8192x8192", 1, 1, 8192, 8192, 8192, 1),
    ]
    for name, nb, h, w, c, n, taps in cases:
        try:
            x = randn((nb, h, w, c), 1).bfloat16()
            wt = randn((n, taps * c), 2, scale=(taps * c) ** -0.5).bfloat16()
            out = torch.empty((nb, h, w, n), device=DEV, dtype=torch.bfloat16)
            for _ in range(3):
                ops.conv_gemm(x, wt, taps, out=out)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            iters = 20
            e0.record()
            for _ in range(iters):
                ops.conv_gemm(x, wt, taps, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            fl = 2.0 * nb * h * w * n * taps * c
            print(f"TIME {name}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
            RESULTS.append({"name": "time " + name, "us": ms * 1e3, "tflops": fl / ms / 1e9, "ok": True})
        except Exception as e:  # noqa: BLE001
            print("EXC time", name, repr(e), flush=True)


def main():
    which = sys.argv[1:] or ["gemm"]
    tag = "_".join(which)
    t0 = time.time()
    print(torch.cuda.get_device_name(0), flush=True)
    for wname in which:
        fn = globals().get("check_" + wname) or globals().get(wname)
        if fn is None:
            print("unknown check", wname)
            continue
        fn()
    os.makedirs("gpurun_out", exist_ok=True)
    nfail = sum(1 for r in RESULTS if not r.get("ok"))
    with open(f"gpurun_out/check_{tag}.json", "w") as f:
        json.dump({"results": RESULTS, "failed": nfail, "seconds": time.time() - t0}, f, indent=1)
    print(f"DONE {len(RESULTS)} checks, {nfail} failed, {time.time()-t0:.1f}s", flush=True)
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
