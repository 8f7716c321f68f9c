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
              tol=1e-2, halo=False, pair=False, persist=False, relu=False, cluster=None):
    def fn():
        x = randn((nb, h, w, c), 1).bfloat16()
        wt = randn((n, taps * c), 2, scale=(taps * c) ** -0.5).bfloat16()
        b = randn((n,), 3) if bias else None
        rv = randn((nb, n), 4) if rowvec else None
        res = randn((nb, h, w, n), 5).bfloat16() if residual else None
        y = ops.conv_gemm(x, wt, taps, bias=b, rowvec=rv, residual=res, out_f32=out_f32, block_n=block_n, splits=splits,
                          halo=halo, pair=pair, persist=persist, act=16 if relu else 0, cluster=cluster)
        torch.cuda.synchronize()
        ref = ref_conv(x, wt, taps, b, rv, res)
        if relu:
            ref = torch.relu(ref)
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

    # 3x3 halo mode (8x16 tiles, three column-shifted halo tiles per channel block)
    gemm_case("halo_16x8_64_64", 1, 16, 8, 64, 64, 9, bias=True, halo=True, block_n=64)
    gemm_case("halo_64x64_320_320_all", 1, 64, 64, 320, 320, 9, bias=True, rowvec=True, residual=True, halo=True)
    gemm_case("halo_64x64_320_320_bn160", 1, 64, 64, 320, 320, 9, bias=True, halo=True, block_n=160)
    gemm_case("halo_32x32_640_640_split3", 1, 32, 32, 640, 640, 9, bias=True, halo=True, block_n=128, splits=3)
    gemm_case("halo_16x16_1280_1280_split6", 1, 16, 16, 1280, 1280, 9, bias=True, rowvec=True, halo=True, block_n=128, splits=6)
    gemm_case("halo_odd_45x80_64", 1, 45, 80, 64, 64, 9, bias=True, halo=True)
    gemm_case("halo_odd_23x40_128_n3", 1, 23, 40, 128, 3, 9, bias=True, out_f32=True, halo=True)
    gemm_case("halo_512x512_64_64", 1, 512, 512, 64, 64, 9, bias=True, halo=True)
    gemm_case("halo_96x96_b4_320", 4, 96, 96, 320, 320, 9, bias=True, halo=True)
    gemm_case("halo_b3_16x16_64", 3, 16, 16, 64, 64, 9, bias=True, halo=True)
    # CTA pairs (cta_group::2): 256-row tiles across two SMs, each CTA streams half of the weight tile
    gemm_case("pair_lin_256x1280_1280", 1, 1, 256, 1280, 1280, 1, bias=True, pair=True, block_n=128)
    gemm_case("pair_lin_4096x320_320_res", 1, 1, 4096, 320, 320, 1, bias=True, residual=True, pair=True, block_n=160)
    gemm_case("pair_lin_odd_tiles_77x768", 1, 1, 77, 768, 768, 1, bias=True, pair=True, block_n=64)
    gemm_case("pair_lin_1024x640_split2", 1, 1, 1024, 640, 640, 1, bias=True, pair=True, block_n=128, splits=2)
    gemm_case("pair_conv_64x64_320_320", 1, 64, 64, 320, 320, 9, bias=True, rowvec=True, residual=True, pair=True, block_n=96)
    gemm_case("pair_conv_16x16_1280_split4", 1, 16, 16, 1280, 1280, 9, bias=True, pair=True, block_n=128, splits=4)
    gemm_case("pair_halo_64x64_320_320", 1, 64, 64, 320, 320, 9, bias=True, residual=True, halo=True, pair=True, block_n=160)
    gemm_case("pair_halo_odd_45x80_64", 1, 45, 80, 64, 64, 9, bias=True, halo=True, pair=True)
    gemm_case("pair_b4_96x96_320", 4, 96, 96, 320, 320, 9, bias=True, pair=True, block_n=256)
    # split-K reduced inside a thread-block cluster (accumulators exchanged through distributed shared memory), and the same
    # shapes through the separate reduce kernel
    for ck in (True, False):
        t = "cluster" if ck else "sepreduce"
        gemm_case(f"{t}_conv_16x16_1280_split4", 1, 16, 16, 1280, 1280, 9, bias=True, rowvec=True, block_n=96, splits=4, cluster=ck)
        gemm_case(f"{t}_conv_16x16_1280_split6_res", 1, 16, 16, 1280, 1280, 9, bias=True, residual=True, block_n=128, splits=6, cluster=ck)
        gemm_case(f"{t}_conv_8x8_1280_split8", 1, 8, 8, 1280, 1280, 9, bias=True, residual=True, block_n=96, splits=8, cluster=ck)
        gemm_case(f"{t}_conv_8x8_2560_split7_bn160", 1, 8, 8, 2560, 1280, 9, bias=True, block_n=160, splits=7, cluster=ck)
        gemm_case(f"{t}_lin_256x1280_split3", 1, 1, 256, 1280, 1280, 1, bias=True, residual=True, block_n=64, splits=3, cluster=ck)
        gemm_case(f"{t}_lin_64x5120_split8_bn32", 1, 1, 64, 5120, 1280, 1, bias=True, residual=True, block_n=32, splits=8, cluster=ck)
        gemm_case(f"{t}_lin_1024x2560_split2_bn256", 1, 1, 1024, 2560, 640, 1, bias=True, block_n=256, splits=2, cluster=ck)
        gemm_case(f"{t}_halo_32x32_640_split3", 1, 32, 32, 640, 640, 9, bias=True, halo=True, block_n=128, splits=3, cluster=ck)
        gemm_case(f"{t}_odd_23x40_640_split5_relu", 1, 23, 40, 640, 640, 9, bias=True, relu=True, block_n=96, splits=5, cluster=ck)
        gemm_case(f"{t}_b2_12x20_1280_split4_f32", 2, 12, 20, 1280, 320, 9, bias=True, out_f32=True, block_n=64, splits=4, cluster=ck)
        gemm_case(f"{t}_n4_fp32out_split4", 1, 64, 64, 320, 4, 9, bias=True, out_f32=True, block_n=32, splits=4, cluster=ck)
    # persistent weight-stationary 3x3 convolution (TAESD shapes): one CTA per SM walks the 8x16 tiles
    gemm_case("persist_64x64_64_64", 1, 64, 64, 64, 64, 9, bias=True, persist=True)
    gemm_case("persist_512x512_64_64_res_relu", 1, 512, 512, 64, 64, 9, bias=True, residual=True, relu=True, persist=True)
    gemm_case("persist_odd_45x80_64_64", 1, 45, 80, 64, 64, 9, bias=True, persist=True)
    gemm_case("persist_b3_32x24_64_64_nobias", 3, 32, 24, 64, 64, 9, persist=True)
    gemm_case("persist_128x128_64_32", 1, 128, 128, 64, 32, 9, bias=True, relu=True, persist=True)
    gemm_case("persist_16x8_64_64_one_tile", 1, 16, 8, 64, 64, 9, bias=True, persist=True)

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

    # LayerNorm folded into the consuming GEMM: row statistics gathered from the operand tiles in shared memory
    F = torch.nn.functional
    for (rows, c, n, bn) in [(4096, 320, 1024, 0), (4096, 320, 1024, 64), (1024, 640, 2048, 128), (256, 1280, 3072, 96), (64, 1280, 1536, 0),
                             (77, 320, 320, 32), (920, 640, 640, 256), (3600, 320, 512, 160)]:
        def ln_a(rows=rows, c=c, n=n, bn=bn):
            x = (randn((rows, c), 71) * 2 + 0.7).bfloat16()
            w = randn((n, c), 72, scale=c ** -0.5).bfloat16()
            gam, bet, b = randn((c,), 73) * 0.3 + 1, randn((c,), 74) * 0.3, randn((n,), 75)
            y = ops.linear_ln(x, w, gam, bet, bias=b, block_n=bn)
            ref = F.layer_norm(x.float(), (c,), gam, bet, 1e-5) @ w.float().t() + b
            record(f"ln_fold_rows_{rows}x{c}x{n}_bn{bn}", rel_err(y, ref), 1e-2)
            y2 = ops.linear_ln(x, w, gam, bet, bias=b, block_n=bn)
            record(f"ln_fold_rows_{rows}x{c}x{n}_bn{bn}_deterministic", float((y.float() - y2.float()).abs().max()), 0.0)
            y3 = ops.linear_ln(x, w, gam, bet, bias=b, block_n=bn, pair=True)      # CTA pairs (cta_group::2)
            record(f"ln_fold_rows_{rows}x{c}x{n}_bn{bn}_pair", rel_err(y3, ref), 1e-2)
        run("ln_fold_rows", ln_a)
    for (rows, c, bn) in [(4096, 320, 0), (4096, 320, 256), (1024, 640, 160), (256, 1280, 64), (64, 1280, 32), (920, 640, 96), (3600, 320, 128)]:
        def ln_b(rows=rows, c=c, bn=bn):
            x = (randn((rows, c), 76) * 1.5 - 0.4).bfloat16()
            w = randn((c, c), 77, scale=c ** -0.5).bfloat16()
            gam, bet = randn((c,), 78) * 0.3 + 1, randn((c,), 79) * 0.3
            y = ops.linear_ln(x, w, gam, bet, swapped=True, block_n=bn)
            ref = (F.layer_norm(x.float(), (c,), gam, bet, 1e-5) @ w.float().t()).t()
            record(f"ln_fold_cols_{rows}x{c}_bn{bn}", rel_err(y[:, :rows], ref), 1e-2)
        run("ln_fold_cols", ln_b)

    # producer side: a GEMM leaves the row statistics of what it stores; chained into a LayerNorm-folded consumer
    for (rows, c, bn, splits, pair) in [(4096, 320, 160, 1, False), (4096, 320, 64, 1, True), (1024, 640, 128, 1, False),
                                        (256, 1280, 96, 3, False), (256, 1280, 128, 1, False), (64, 1280, 64, 4, False),
                                        (64, 1280, 32, 8, False), (920, 640, 96, 2, False), (77, 320, 32, 1, False)]:
        def chain(rows=rows, c=c, bn=bn, splits=splits, pair=pair):
            x = randn((rows, c), 91).bfloat16()
            w1 = randn((c, c), 92, scale=c ** -0.5).bfloat16()
            b1 = randn((c,), 93)
            res = (randn((rows, c), 94) + 0.5).bfloat16()
            h, st = ops.linear_stats(x, w1, bias=b1, residual=res, block_n=bn, splits=splits, pair=pair)
            href = x.float() @ w1.float().t() + b1 + res.float()
            record(f"rowstats_out_{rows}x{c}_bn{bn}_s{splits}_p{int(pair)}", rel_err(h, href), 1e-2)
            tot = st.sum(dim=1)
            record(f"rowstats_sum_{rows}x{c}_bn{bn}_s{splits}", rel_err(tot[:, 0], href.sum(-1)), 2e-3)
            record(f"rowstats_sumsq_{rows}x{c}_bn{bn}_s{splits}", rel_err(tot[:, 1], (href * href).sum(-1)), 2e-3)
            w2 = randn((2 * c, c), 95, scale=c ** -0.5).bfloat16()
            gam, bet = randn((c,), 96) * 0.3 + 1, randn((c,), 97) * 0.3
            y = ops.linear_ln(h, w2, gam, bet, stats=st)
            ref = F.layer_norm(h.float(), (c,), gam, bet, 1e-5) @ w2.float().t()
            record(f"ln_fold_chain_{rows}x{c}_bn{bn}_s{splits}", rel_err(y, ref), 1e-2)
            yt = ops.linear_ln(h, w2[:c].contiguous(), gam, bet, swapped=True, stats=st)
            record(f"ln_fold_chain_cols_{rows}x{c}_bn{bn}_s{splits}", rel_err(yt[:, :rows], ref[:, :c].t()), 1e-2)
        run("ln_fold_chain", chain)

    def ln_geglu():
        for (m, c, bn, pair) in [(4096, 320, 128, False), (256, 1280, 128, False), (64, 1280, 128, False), (4096, 320, 256, False),
                                 (4096, 320, 256, True), (1024, 640, 128, True), (256, 1280, 256, True), (64, 1280, 256, False)]:
            x = (randn((m, c), 81) * 2 + 0.3).bfloat16()
            w_full = randn((8 * c, c), 82, scale=c ** -0.5)
            b_full = randn((8 * c,), 83)
            gam, bet = randn((c,), 84) * 0.3 + 1, randn((c,), 85) * 0.3
            half = 4 * c
            idx = []
            for t in range(half // 64):
                idx += list(range(t * 64, t * 64 + 64)) + list(range(half + t * 64, half + t * 64 + 64))
            idx = torch.tensor(idx, device=DEV)
            y = ops.linear_ln(x, w_full[idx].bfloat16().contiguous(), gam, bet, bias=b_full[idx].contiguous(), act=1, block_n=bn, pair=pair)
            hfull = F.layer_norm(x.float(), (c,), gam, bet, 1e-5) @ w_full.bfloat16().float().t() + b_full
            ref = hfull[:, :half] * F.gelu(hfull[:, half:])
            record(f"ln_fold_geglu_{m}x{c}_bn{bn}_p{int(pair)}", rel_err(y, ref), 1e-2)
    run("ln_fold_geglu", ln_geglu)


def tuning_table_keys():
    """GEMM shape keys of every committed tuning table (videosd_b200/tuning/*.txt), deduplicated."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "videosd_b200", "tuning")
    keys = set()
    for fn in sorted(os.listdir(root)):
        if fn.endswith(".txt"):
            with open(os.path.join(root, fn)) as f:
                for ln in f:
                    if "|" in ln:
                        keys.add(ln.split()[0])
    return sorted(keys)


def check_tuner_sweep(keys=None):
    """For every GEMM shape of the committed tuning tables: EVERY configuration the autotuner may pick (block_n, split-K, CTAs
    per SM, k-blocks per stage, halo / pairs / persistent / in-cluster reduce -- the list comes from the library itself,
    vsd_op_gemm_candidates) against torch fp32. What the tuner can select is what is tested."""
    F = torch.nn.functional
    total = 0
    for key in (keys or tuning_table_keys()):
        def one(key=key):
            nonlocal total
            dims, t, n, a, f, r = key.split("|")
            nb, h, w, c = (int(v) for v in dims.split("x"))
            tcode, n, act_key, out_f32, has_res = int(t[1:]), int(n[1:]), int(a[1:]), int(f[1:]), int(r[1:])
            taps, stride2, pad = tcode % 100, (tcode // 100) % 10 == 1, 0 if tcode >= 1000 else 1
            if act_key & 128:
                return                                     # fp32 residual stream (AutoencoderKL attention): no operator-level wrapper
            base = act_key & 0xF
            act = base | (act_key & (16 | 32 | 64))        # the LayerNorm / row-statistics variants are swept as the plain GEMM
            x = randn((nb, h, w, c), 201).bfloat16()
            wt_f = randn((n, taps * c), 202, scale=(taps * c) ** -0.5)
            ho, wo = ((h + 2 * pad - 3 + (0 if pad else 1)) // 2 + 1, (w + 2 * pad - 3 + (0 if pad else 1)) // 2 + 1) if stride2 else (h, w)
            bias = randn((n,), 203)
            rowvec = randn((nb, n), 204) if (taps == 9 and not has_res and not out_f32 and base == 0 and not stride2) else None
            n_out = n // 2 if base == 1 else n
            res = randn((nb, ho, wo, n_out), 205).bfloat16() if has_res else None
            xf = x.float().permute(0, 3, 1, 2)
            if base == 1:                                  # GEGLU: value / gate rows interleaved per 128-row tile
                half = n // 2
                idx = []
                for tt in range(half // 64):
                    idx += list(range(tt * 64, tt * 64 + 64)) + list(range(half + tt * 64, half + tt * 64 + 64))
                idx = torch.tensor(idx, device=DEV)
                wt = wt_f[idx].bfloat16().contiguous()
                b_dev = bias[idx].contiguous()
                hfull = x.float().reshape(-1, c) @ wt_f.bfloat16().float().t() + bias
                ref = (hfull[:, :half] * F.gelu(hfull[:, half:])).view(nb, h, w, half)
            else:
                wt = wt_f.bfloat16().contiguous()
                b_dev = bias
                wf = wt.float()
                if taps == 9:
                    w4 = wf.view(n, 3, 3, c).permute(0, 3, 1, 2)
                    if stride2 and pad == 0:
                        y = F.conv2d(F.pad(xf, (0, 1, 0, 1)), w4, stride=2)
                    else:
                        y = F.conv2d(xf, w4, padding=1, stride=2 if stride2 else 1)
                else:
                    y = F.conv2d(xf, wf.view(n, c, 1, 1))
                ref = y.permute(0, 2, 3, 1) + bias.view(1, 1, 1, n)
                if rowvec is not None:
                    ref = ref + rowvec.view(nb, 1, 1, n)
                if res is not None:
                    ref = ref + res.float()
                if act_key & 16:
                    ref = torch.relu(ref)
                if base == 2:
                    ref = ref * torch.sigmoid(1.702 * ref)
            out = torch.empty((nb, ho, wo, n_out), device=DEV, dtype=torch.float32 if out_f32 else torch.bfloat16)
            cands = ops.gemm_candidates(x, wt, taps, out, bias=b_dev, rowvec=rowvec, residual=res, act=act, stride2=stride2, pad=pad)
            worst, worst_cfg = 0.0, None
            scale = float(ref.abs().max().clamp_min(1e-12))
            for cfg in cands:
                out.fill_(float("nan"))
                ops.conv_gemm_cfg(x, wt, taps, out, cfg, bias=b_dev, rowvec=rowvec, residual=res, act=act, stride2=stride2, pad=pad)
                d = (out.float() - ref).abs().max()
                e = float(d) / scale
                if not (e <= worst):                        # also catches NaN
                    worst, worst_cfg = e, cfg
            total += len(cands)
            record(f"sweep_{key}", worst if cands else float("nan"), 1e-2, {"candidates": len(cands), "worst_cfg": worst_cfg})
        run(f"sweep_{key}", one)
    print(f"SWEEP {total} configurations", flush=True)


def check_s2_sweep():
    """3x3 stride-2 convolutions through TMA element strides (VSD_TMA_S2=1 in the engine): every tuner candidate at the odd
    360x640 geometry (45x80 -> 23x40 -> 12x20 -> 6x10) and at 512 / 768, pad 1 (UNet / TAESD) and pad 0 (AutoencoderKL)."""
    keys = []
    for (h, w, c) in [(45, 80, 320), (23, 40, 640), (12, 20, 1280), (64, 64, 320), (32, 32, 640), (16, 16, 1280), (96, 96, 320),
                      (90, 160, 64), (180, 320, 64), (128, 128, 64)]:
        keys.append(f"1x{h}x{w}x{c}|t109|n{c}|a0|f0|r0")
    keys += ["2x45x80x320|t109|n320|a0|f0|r0", "1x45x80x128|t1109|n128|a0|f0|r0", "1x64x64x128|t1109|n128|a0|f0|r0"]
    check_tuner_sweep(keys)


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
    for name, nb, h, w, c, n, taps in cases + [(nm + " HALO", a, b_, c_, d, e, f) for (nm, a, b_, c_, d, e, f) in cases if f == 9 and b_ >= 16]:
        try:
            x = randn((nb, h, w, c), 1).bfloat16()
            wt = randn((n, taps * c), 2, scale=(taps * c) ** -0.5).bfloat16()
            out = torch.empty((nb, h, w, n), device=DEV, dtype=torch.bfloat16)
            hl = name.endswith("HALO")
            for _ in range(3):
                ops.conv_gemm(x, wt, taps, out=out, halo=hl)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            iters = 20
            e0.record()
            for _ in range(iters):
                ops.conv_gemm(x, wt, taps, out=out, halo=hl)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            fl = 2.0 * nb * h * w * n * taps * c
            print(f"TIME {name}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
            RESULTS.append({"name": "time " + name, "us": ms * 1e3, "tflops": fl / ms / 1e9, "ok": True})
        except Exception as e:  # noqa: BLE001
            print("EXC time", name, repr(e), flush=True)


# ------------------------------------------------------------------------------------------------ attention
def attn_case(name, batch, heads, d, nq, nk, k_slot=None, tol=1.5e-2, qscale=1.0):
    def fn():
        q = randn((batch, nq, heads, d), 21) * qscale
        k = randn((batch, nk, heads, d), 22)
        v = randn((batch, nk, heads, d), 23)
        qb, kb, vb = q.bfloat16(), k.bfloat16(), v.bfloat16()
        slot = nk if k_slot is None else k_slot
        qp = ops.pad_heads(qb.reshape(batch * nq, heads * d), heads, d)
        kfull = torch.zeros((batch, slot, heads * d), device=DEV, dtype=torch.bfloat16)
        kfull[:, :nk] = kb.reshape(batch, nk, heads * d)
        kp = ops.pad_heads(kfull.reshape(batch * slot, heads * d), heads, d)
        vfull = torch.zeros((batch, slot, heads * d), device=DEV, dtype=torch.bfloat16)
        vfull[:, :nk] = vb.reshape(batch, nk, heads * d)
        vt = vfull.reshape(batch * slot, heads * d).t().contiguous()
        o = ops.attention(qp, kp, vt, batch, heads, d, nq, nk, k_rows_per_img=slot, vt_cols_per_img=slot)
        torch.cuda.synchronize()
        # plain fp32 reference softmax(Q K^T / sqrt(d)) V (matmul + softmax; no fused library attention kernel in the checker)
        qf, kf, vf = qb.float().transpose(1, 2), kb.float().transpose(1, 2), vb.float().transpose(1, 2)
        ref = torch.empty_like(qf)
        for bi in range(batch):
            for hi in range(heads):
                sc = (qf[bi, hi] @ kf[bi, hi].t()) * (d ** -0.5)
                ref[bi, hi] = torch.softmax(sc, dim=-1) @ vf[bi, hi]
        ref = ref.transpose(1, 2).reshape(batch * nq, heads * d)
        record(name, rel_err(o, ref), tol, {"b": batch, "h": heads, "d": d, "nq": nq, "nk": nk})

    run(name, fn)


def check_attn():
    attn_case("attn_d64_128x128", 1, 1, 64, 128, 128)
    attn_case("attn_d40_128x128_h8", 1, 8, 40, 128, 128)
    attn_case("attn_d40_256x256", 1, 8, 40, 256, 256)
    attn_case("attn_d40_4096", 1, 8, 40, 4096, 4096)
    attn_case("attn_d80_1024", 1, 8, 80, 1024, 1024)
    attn_case("attn_d160_256", 1, 8, 160, 256, 256)
    attn_case("attn_d160_64", 1, 8, 160, 64, 64, k_slot=64)
    attn_case("attn_d40_cross77", 1, 8, 40, 4096, 77, k_slot=128)
    attn_case("attn_d80_cross77_b2", 2, 8, 80, 1024, 77, k_slot=128)
    attn_case("attn_d160_cross77", 1, 8, 160, 64, 77, k_slot=128)
    attn_case("attn_d40_b2_1024", 2, 8, 40, 1024, 1024)
    attn_case("attn_d40_odd_b2_920", 2, 8, 40, 920, 920)
    attn_case("attn_d40_9216", 1, 8, 40, 9216, 9216)
    # two-tile kernel edge cases: a lone first tile in the last CTA, partial last key block, batch > 1, d = 64, 2 key blocks
    attn_case("attn_d40_3600", 1, 8, 40, 3600, 3600)
    attn_case("attn_d40_b3_400", 3, 8, 40, 400, 400)
    attn_case("attn_d64_b2_1024", 2, 4, 64, 1024, 1024)
    attn_case("attn_d48_b2_129x300", 2, 2, 48, 129, 300, k_slot=304)
    attn_case("attn_d40_b4_2304", 4, 8, 40, 2304, 2304)
    attn_case("attn_d8_520", 1, 3, 8, 520, 520)
    # peaky rows: the running maximum moves by more than 2^8 between key blocks (the lazy O rescale path)
    attn_case("attn_d40_1024_peaky", 1, 8, 40, 1024, 1024, qscale=12.0)
    attn_case("attn_d40_b2_920_peaky", 2, 8, 40, 920, 920, qscale=30.0)



def check_attn_timing():
    """Rough timings (CUDA events, 10 back-to-back launches); not a benchmark of record."""
    def timing():
        for (b, h, d, n) in [(1, 8, 40, 4096), (1, 8, 80, 1024), (1, 8, 160, 256), (4, 8, 40, 9216), (1, 8, 40, 9216), (1, 8, 64, 4096)]:
            q = randn((b * n, h * d), 1).bfloat16()
            qp = ops.pad_heads(q, h, d)
            vt = q.t().contiguous()
            out = torch.empty((b * n, h * d), device=DEV, dtype=torch.bfloat16)
            for _ in range(3):
                ops.attention(qp, qp, vt, b, h, d, n, n, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.attention(qp, qp, vt, b, h, d, n, n, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            fl = 4.0 * b * h * n * n * d
            print(f"TIME attn b{b} h{h} d{d} n{n}: {ms*1e3:.1f} us {fl/ms/1e9:.1f} TFLOP/s", flush=True)

    run("attn_timing", timing)


# ------------------------------------------------------------------------------------------------ bandwidth kernels
def check_norm():
    for (nb, h, w, c, eps, silu) in [(1, 64, 64, 320, 1e-5, True), (1, 32, 32, 640, 1e-6, False), (2, 16, 16, 1280, 1e-5, True),
                                     (1, 8, 8, 2560, 1e-5, True), (1, 16, 16, 1920, 1e-5, True), (1, 32, 32, 960, 1e-5, True),
                                     (4, 96, 96, 320, 1e-5, True), (1, 23, 40, 640, 1e-5, True), (1, 45, 80, 960, 1e-5, True),
                                     (1, 8, 8, 1280, 1e-5, True), (2, 64, 64, 128, 1e-6, True), (1, 32, 32, 512, 1e-6, False),
                                     (1, 64, 64, 256, 1e-6, True), (4, 96, 96, 960, 1e-5, True), (1, 256, 256, 128, 1e-6, True),
                                     (3, 12, 20, 1280, 1e-5, False)]:
        def fn(nb=nb, h=h, w=w, c=c, eps=eps, silu=silu):
            x = (randn((nb, h, w, c), 31) * 2 + 0.5).bfloat16()
            gam, bet = randn((c,), 32) * 0.2 + 1, randn((c,), 33) * 0.2
            y = ops.groupnorm(x, gam, bet, 32, eps, silu)
            ref = torch.nn.functional.group_norm(x.float().permute(0, 3, 1, 2), 32, gam, bet, eps)
            if silu:
                ref = torch.nn.functional.silu(ref)
            record(f"groupnorm_{nb}x{h}x{w}x{c}", rel_err(y, ref.permute(0, 2, 3, 1)), 1e-2)
            y2 = ops.groupnorm(x, gam, bet, 32, eps, silu)      # statistics are reduced in a fixed order: bit-identical reruns
            record(f"groupnorm_{nb}x{h}x{w}x{c}_deterministic", float((y.float() - y2.float()).abs().max()), 0.0)
        run("groupnorm", fn)
    # strided input (a channel slice of a wider concat buffer)
    def gn_strided():
        buf = (randn((1, 32, 32, 1280), 34)).bfloat16()
        x = buf[..., 640:]
        gam, bet = randn((640,), 35) + 1, randn((640,), 36)
        y = ops.groupnorm(x, gam, bet, 32, 1e-5, True)
        ref = torch.nn.functional.silu(torch.nn.functional.group_norm(x.float().permute(0, 3, 1, 2), 32, gam, bet, 1e-5))
        record("groupnorm_strided", rel_err(y, ref.permute(0, 2, 3, 1)), 1e-2)
    run("groupnorm_strided", gn_strided)
    for (rows, c) in [(4096, 320), (1024, 640), (256, 1280), (77, 320), (36864, 320)]:
        def fn(rows=rows, c=c):
            x = (randn((rows, c), 41) * 3 + 1).bfloat16()
            gam, bet = randn((c,), 42) * 0.2 + 1, randn((c,), 43) * 0.2
            y = ops.layernorm(x, gam, bet)
            ref = torch.nn.functional.layer_norm(x.float(), (c,), gam, bet, 1e-5)
            record(f"layernorm_{rows}x{c}", rel_err(y, ref), 1e-2)
        run("layernorm", fn)


def check_misc():
    def ups():
        x = randn((2, 16, 16, 640), 51).bfloat16()
        y = ops.upsample_nearest(x, 32, 32)
        ref = torch.nn.functional.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
        record("upsample_2x", (y.float() - ref.permute(0, 2, 3, 1)).abs().max().item(), 0.0)
        x = randn((1, 12, 20, 320), 52).bfloat16()
        y = ops.upsample_nearest(x, 23, 40)
        ref = torch.nn.functional.interpolate(x.float().permute(0, 3, 1, 2), size=(23, 40), mode="nearest")
        record("upsample_size", (y.float() - ref.permute(0, 2, 3, 1)).abs().max().item(), 0.0)
    run("upsample", ups)

    def s2():
        for (nb, h, w, c, n) in [(1, 64, 64, 320, 320), (2, 45, 80, 64, 64), (1, 512, 512, 64, 64)]:
            x = randn((nb, h, w, c), 53).bfloat16()
            wt = randn((n, 9 * c), 54, scale=(9 * c) ** -0.5).bfloat16()
            b = randn((n,), 55)
            cols = ops.im2col_s2(x)
            y = ops.conv_gemm(cols, wt, 1, bias=b)
            ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().view(n, 3, 3, c).permute(0, 3, 1, 2),
                                             b, stride=2, padding=1).permute(0, 2, 3, 1)
            record(f"conv3x3_s2_{nb}x{h}x{w}x{c}", rel_err(y, ref), 1e-2)
    run("conv_s2", s2)

    def s2_tma():
        # the same stride-2 convolutions without im2col: TMA element strides (2, 2); symmetric and right/bottom-only padding
        F = torch.nn.functional
        for (nb, h, w, c, n, pad) in [(1, 64, 64, 320, 320, 1), (2, 45, 80, 64, 64, 1), (1, 512, 512, 64, 64, 1), (1, 16, 16, 1280, 1280, 1),
                                      (1, 64, 64, 128, 128, 0), (2, 32, 48, 256, 256, 0), (1, 33, 47, 64, 96, 1)]:
            x = randn((nb, h, w, c), 61).bfloat16()
            wt = randn((n, 9 * c), 62, scale=(9 * c) ** -0.5).bfloat16()
            b = randn((n,), 63)
            y = ops.conv_gemm(x, wt, 9, bias=b, stride2=True, pad=pad)
            xf = x.float().permute(0, 3, 1, 2)
            wf = wt.float().view(n, 3, 3, c).permute(0, 3, 1, 2)
            ref = (F.conv2d(xf, wf, b, stride=2, padding=1) if pad else F.conv2d(F.pad(xf, (0, 1, 0, 1)), wf, b, stride=2)).permute(0, 2, 3, 1)
            ok_shape = tuple(y.shape) == tuple(ref.shape)
            record(f"conv3x3_s2_tma_{nb}x{h}x{w}x{c}_pad{pad}", rel_err(y, ref) if ok_shape else 1.0, 1e-2, {"shape": list(y.shape)})
    run("conv_s2_tma", s2_tma)

    def small():
        # UNet conv_in: fp32 latents, 4 -> 320
        x = randn((2, 64, 64, 4), 56)
        w = randn((320, 3, 3, 4), 57, scale=1 / 6.0)
        b = randn((320,), 58)
        y = ops.conv3x3_small_cin(x, 0, w, b)
        ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), b, padding=1).permute(0, 2, 3, 1)
        record("conv_in_4_320", rel_err(y, ref), 5e-3)
        # TAESD encoder first conv from u8 RGB
        u8 = torch.randint(0, 256, (1, 128, 96, 3), generator=g(59), dtype=torch.uint8).to(DEV)
        w = randn((64, 3, 3, 3), 60, scale=0.2)
        b = randn((64,), 61)
        y = ops.conv3x3_small_cin(u8, 1, w, b)
        xin = ((2.0 * (u8.float() / 255.0) - 1.0) + 1) / 2
        ref = torch.nn.functional.conv2d(xin.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), b, padding=1).permute(0, 2, 3, 1)
        record("taesd_enc_conv0_u8", rel_err(y, ref), 5e-3)
        # TAESD decoder first conv with tanh prologue + ReLU
        z = randn((1, 64, 64, 4), 62) * 3
        w = randn((64, 3, 3, 4), 63, scale=0.2)
        b = randn((64,), 64)
        y = ops.conv3x3_small_cin(z, 2, w, b, relu=True)
        zin = torch.tanh(z / 3) * 3
        ref = torch.relu(torch.nn.functional.conv2d(zin.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), b, padding=1)).permute(0, 2, 3, 1)
        record("taesd_dec_conv0_tanh", rel_err(y, ref), 5e-3)
    run("small_cin", small)

    def sched():
        from oracle.scheduler import LCMSchedulerOracle
        s = LCMSchedulerOracle()
        s.set_timesteps(0.5, 4)
        x0, nz = randn((1, 64, 64, 4), 70), randn((1, 64, 64, 4), 71)
        sc = s.step_scalars(0)
        out = ops.add_noise(x0, nz, float(sc["sqrt_alpha"]), float(sc["sqrt_beta"]))
        ref = s.add_noise(x0.cpu(), nz.cpu(), s.timesteps[:1])
        record("add_noise_bitexact", (out.cpu() - ref).abs().max().item(), 0.0)
        for i in range(4):
            eps, x, z = randn((2, 96, 96, 4), 72 + i), randn((2, 96, 96, 4), 80 + i), randn((2, 96, 96, 4), 90 + i)
            sc = {k: float(v) for k, v in s.step_scalars(i).items()}
            xp, den = ops.lcm_step(eps, x, z, sc)
            rp, rd = s.step(eps.cpu(), i, x.cpu(), z.cpu())
            record(f"lcm_step{i}_prev_bitexact", (xp.cpu() - rp).abs().max().item(), 0.0)
            record(f"lcm_step{i}_den_bitexact", (den.cpu() - rd).abs().max().item(), 0.0)
    run("scheduler", sched)

    def colour():
        import numpy as np
        from oracle import imageproc
        for (h, w) in [(512, 512), (360, 640), (768, 768), (2, 4)]:
            gg = g(100 + h)
            y = torch.randint(0, 256, (2, h, w), generator=gg, dtype=torch.uint8)
            u = torch.randint(0, 256, (2, h // 2, w // 2), generator=gg, dtype=torch.uint8)
            v = torch.randint(0, 256, (2, h // 2, w // 2), generator=gg, dtype=torch.uint8)
            rgb = ops.yuv420_to_rgb(y.to(DEV), u.to(DEV), v.to(DEV)).cpu().numpy()
            ref = np.stack([imageproc.yuv420_to_rgb(y[i].numpy(), u[i].numpy(), v[i].numpy()) for i in range(2)])
            record(f"yuv420_to_rgb_{h}x{w}_bitexact", float(np.abs(rgb.astype(int) - ref.astype(int)).max()), 0.0)
            img = torch.rand((2, h, w, 4), generator=gg) * 1.4 - 0.2   # decoder output before *2-1, incl. out-of-range
            # exact .5 ties for round-half-even
            img[0, 0, 0, :3] = torch.tensor([0.5 / 255, 1.5 / 255, 2.5 / 255])
            r8, yy, uu, vv = ops.pack_rgb_yuv420(img.to(DEV), taesd_denorm=True)
            ref_rgb = imageproc.postprocess((img[..., :3] * 2 - 1).permute(0, 3, 1, 2))
            record(f"pack_rgb_{h}x{w}_bitexact", float(np.abs(r8.cpu().numpy().astype(int) - ref_rgb.astype(int)).max()), 0.0)
            worst = 0
            for i in range(2):
                ry, ru, rv = imageproc.rgb_to_yuv420(ref_rgb[i])
                worst = max(worst, np.abs(yy[i].cpu().numpy().astype(int) - ry).max(), np.abs(uu[i].cpu().numpy().astype(int) - ru).max(),
                            np.abs(vv[i].cpu().numpy().astype(int) - rv).max())
            record(f"rgb_to_yuv420_{h}x{w}_bitexact", float(worst), 0.0)
    run("colour", colour)


def check_controlnet():
    import numpy as np

    def sobel():
        gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                    "reference_golden.npz"))
        rgb = torch.from_numpy(gold["sobel_rgb_in"])[None].to(DEV)
        ctl = ops.sobel_control(rgb)
        got = (ctl[0, :, :, 0] * 255).round().cpu().numpy().astype(int)
        ref = gold["sobel_out"].astype(int)      # produced by the reference's SobelOperator
        d = np.abs(got - ref)
        record("sobel_vs_reference_golden_maxdiff", float(d.max()), 1.0, {"mismatch_frac": float((d > 0).mean())})
        record("sobel_vs_reference_golden_mismatch_frac", float((d > 0).mean()), 1e-3)
        assert torch.equal(ctl[..., 0], ctl[..., 1]) and torch.equal(ctl[..., 0], ctl[..., 2])
    run("sobel", sobel)

    def direct():
        for (h, w, cin, cout, stride) in [(64, 64, 16, 16, 1), (64, 96, 16, 32, 2), (33, 47, 32, 96, 2), (32, 32, 96, 96, 1), (32, 32, 96, 256, 2)]:
            x = randn((2, h, w, cin), 201).bfloat16()
            wt = randn((cout, 9 * cin), 202, scale=(9 * cin) ** -0.5).bfloat16()
            b = randn((cout,), 203)
            y = ops.conv3x3_direct(x, wt, b, stride, True)
            ref = torch.nn.functional.silu(torch.nn.functional.conv2d(
                x.float().permute(0, 3, 1, 2), wt.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2), b, stride=stride, padding=1)).permute(0, 2, 3, 1)
            record(f"direct_conv_{h}x{w}_{cin}_{cout}_s{stride}", rel_err(y, ref), 5e-3)
    run("direct_conv", direct)


def check_clip():
    """CLIP text encoder on the GPU vs the oracle restatement (itself pinned against transformers.CLIPTextModel)."""
    from oracle.clip import ClipTextOracle
    from videosd_b200 import weights
    from videosd_b200.engine import Engine

    def enc():
        sd = weights.random_clip_state_dict(5)
        ref_model = ClipTextOracle()
        ref_model.load_state_dict(sd)
        eng = Engine(0)
        eng.load_state_dict("text_encoder", sd)
        g = torch.Generator().manual_seed(1)
        ids = torch.randint(0, 49406, (3, 77), generator=g)
        ids[:, 0] = 49406
        ids[1, 9:] = 49407
        ids[2, 3:] = 49407
        ref = ref_model(ids)
        outs = []
        for b in range(3):
            got = eng.encode_prompt(ids[b].tolist())
            outs.append(got)
            record(f"clip_last_hidden_state_rel_{b}", ((got - ref[b]).norm() / ref[b].norm()).item(), 1e-2,
                   {"max_abs": (got - ref[b]).abs().max().item()})
        again = eng.encode_prompt(ids[0].tolist())
        record("clip_deterministic", 0.0 if torch.equal(again, outs[0]) else 1.0, 0.0)
        ids2 = ids[0].clone()
        ids2[40] = 123
        got2 = eng.encode_prompt(ids2.tolist())
        record("clip_causal_prefix_unchanged", 0.0 if torch.equal(got2[:40], outs[0][:40]) else 1.0, 0.0)
        record("clip_causal_suffix_changed", 0.0 if not torch.equal(got2[40:], outs[0][40:]) else 1.0, 0.0)
        t0 = time.time()
        for _ in range(20):
            eng.encode_prompt(ids[0].tolist())
        record("clip_encode_ms", (time.time() - t0) / 20 * 1e3, 1e9)
    run("clip_encode", enc)


def check_resize():
    """Center crop + Lanczos on the GPU vs Pillow itself (the reference's own CPU path, videopipeline.py:92-107)."""
    import numpy as np
    from PIL import Image

    from videosd_b200.videopipeline import VideoSDPipeline

    def vs_pil():
        geos = [(640, 480, 512, 512), (1280, 720, 640, 360), (640, 480, 768, 768), (320, 240, 512, 512), (500, 375, 512, 384),
                (641, 479, 256, 256), (640, 512, 512, 512), (1920, 1080, 512, 512), (512, 640, 512, 512), (300, 300, 512, 512)]
        for (iw, ih, w, h) in geos:
            rs = np.random.RandomState(iw + ih)
            src = rs.randint(0, 256, (2, ih, iw, 3)).astype(np.uint8)
            src[1, ::7] = 255; src[1, 3::11] = 0          # hard edges: exercises the clamp on ringing
            got = ops.crop_resize(torch.from_numpy(src).to(DEV), w, h).cpu().numpy()
            ref = np.stack([np.asarray(VideoSDPipeline._fit(Image.fromarray(s), w, h)) for s in src])
            d = np.abs(got.astype(int) - ref.astype(int))
            record(f"crop_lanczos_{iw}x{ih}_to_{w}x{h}_maxdiff", float(d.max()), 0.0)
    run("crop_resize_vs_pil", vs_pil)

    def timing():
        src = torch.randint(0, 256, (1, 720, 1280, 3), dtype=torch.uint8, device=DEV)
        for _ in range(3):
            ops.crop_resize(src, 512, 512)
        torch.cuda.synchronize()
        from videosd_b200 import resample
        plan = resample.resize_plan(1280, 720, 512, 512)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # time only the kernels: tables resident, as in the engine
        from ctypes import c_int
        from videosd_b200._lib import _p, check, cur_stream, lib
        x0, y0, cw, ch = plan["crop"]
        (hb, hk, hks), (vb, vk, vks) = plan["h"], plan["v"]
        t = lambda z: torch.from_numpy(z.astype("int32")).contiguous().to(DEV)  # noqa: E731
        hb, hk, vb, vk = t(hb), t(hk), t(vb), t(vk)
        tmp = torch.empty((1, ch, 512, 3), device=DEV, dtype=torch.uint8)
        out = torch.empty((1, 512, 512, 3), device=DEV, dtype=torch.uint8)
        a.record()
        for _ in range(50):
            check(lib().vsd_op_crop_resize(_p(src), c_int(1280), c_int(720), c_int(x0), c_int(y0), c_int(cw), c_int(ch), _p(tmp), _p(out),
                                           c_int(512), c_int(512), _p(hb), _p(hk), c_int(hks), _p(vb), _p(vk), c_int(vks), c_int(1),
                                           cur_stream()), "crop_resize")
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / 50 * 1e3
        t0 = time.time()
        img = Image.fromarray(src[0].cpu().numpy())
        for _ in range(10):
            VideoSDPipeline._fit(img, 512, 512)
        cpu_us = (time.time() - t0) / 10 * 1e6
        record("crop_lanczos_720p_to_512_us", us, 1e9, {"pil_cpu_us": cpu_us})
    run("crop_resize_timing", timing)


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
