"""End-to-end parity + timing of the engine against the fp32 oracle (run on a B200 via gpurun).

    python tools/gpu_pipeline_check.py [HxW[xB]] [--no-time]
Writes gpurun_out/pipeline_<H>x<W>x<B>.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import imageproc, pipeline  # noqa: E402
from oracle.scheduler import LCMSchedulerOracle, w_embedding  # noqa: E402
from oracle.weights import build_taesd, build_unet, random_context  # noqa: E402
from videosd_b200.engine import Engine  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def psnr(a, b):
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def main():
    spec = "512x512x1"
    for a in sys.argv[1:]:
        if a[0].isdigit():
            spec = a
    parts = [int(v) for v in spec.split("x")]
    H, W = parts[0], parts[1]
    B = parts[2] if len(parts) > 2 else 1
    do_time = "--no-time" not in sys.argv
    use_cn = "--controlnet" in sys.argv
    use_kl = "--kl" in sys.argv
    cn_scale = 0.7
    rep = {"H": H, "W": W, "B": B}
    t0 = time.time()
    unet, vae = build_unet(), build_taesd()
    kl_net = vae_noise = None
    if use_kl:
        from oracle.weights import KLAdapter, build_vae_kl
        kl_net = build_vae_kl()
        vae_noise = torch.randn((B, 4, H // 8, W // 8), generator=torch.Generator().manual_seed(99))
        vae = KLAdapter(kl_net, vae_noise)
    cn = None
    if use_cn:
        from oracle.weights import build_controlnet
        cn = build_controlnet()
    ctx = random_context(B)
    print(f"oracle built in {time.time()-t0:.1f}s", flush=True)
    eng = Engine(0)
    if os.environ.get("TUNE_FOR"):
        eng.set_autotune(int(os.environ["TUNE_FOR"]))
    t0 = time.time()
    eng.load_state_dict("unet", unet.state_dict())
    if use_kl:
        eng.load_state_dict("vae_kl", kl_net.state_dict())
    else:
        eng.load_state_dict("vae", vae.state_dict())
    if use_cn:
        eng.load_state_dict("controlnet", cn.state_dict())
    print(f"weights loaded in {time.time()-t0:.1f}s", flush=True)
    eng.configure(B, H, W)
    if use_kl:
        eng.set_vae("kl")
        eng.set_vae_noise(vae_noise)
    if use_cn:
        eng.set_controlnet(True, cn_scale)
    ts = eng.set_schedule(0.5, 4)
    for b in range(B):
        eng.set_context(b, ctx[b])
    eng.set_reference_noise()
    print("timesteps", ts, "launches/frame", eng.launches_per_frame(), "arena peak MB", eng.arena_peak_bytes() / 2 ** 20, flush=True)
    rep["launches_per_frame"] = eng.launches_per_frame()
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/tuning_{H}x{W}x{B}.txt", "w") as f:
        f.write(eng.tuning_report())

    if use_kl:
        kl_net.cuda()
        unet_g, vae_g = unet.cuda(), vae
    else:
        unet_g, vae_g = unet.cuda(), vae.cuda()
    cn_g = cn.cuda() if use_cn else None
    frames = [imageproc.synthetic_frame(H, W, seed=b, shift=17 * b) for b in range(B)]
    y = np.stack([f[0] for f in frames]); u = np.stack([f[1] for f in frames]); v = np.stack([f[2] for f in frames])
    rgb = np.stack([imageproc.yuv420_to_rgb(*f) for f in frames])
    t0 = time.time()
    ref = pipeline.lcm_img2img(unet_g, vae_g, rgb, ctx, steps=4, strength=0.5, device="cuda", controlnet=cn_g,
                               controlnet_scale=cn_scale)
    torch.cuda.synchronize()
    print(f"oracle (fp32 on GPU) frame in {time.time()-t0:.2f}s", flush=True)

    if use_cn:   # the control front end runs inside the frame plan: run one frame so its buffers are valid
        _y = np.empty_like(y); _u = np.empty_like(u); _v = np.empty_like(v)
        eng.infer_yuv420(y, u, v, _y, _u, _v)
    # ---- teacher-forced per-step parity: same input latents to both UNets
    sched = LCMSchedulerOracle(); sched.set_timesteps(0.5, 4)
    _, step_noise = pipeline.frame_noise(B, H // 8, W // 8, 4)
    rep["per_step"] = []
    for i in range(4):
        lat_in = ref["latents_in"][i].cpu()
        eps = eng.debug_unet(lat_in, i)
        e_eps = rel(eps, ref["eps"][i])
        lat, den = sched.step(eps, i, lat_in, step_noise[i])
        e_lat = rel(lat, ref["latents"][i])
        e_den = rel(den, ref["denoised_steps"][i])
        rep["per_step"].append({"step": i, "eps_rel": e_eps, "latents_rel": e_lat, "denoised_rel": e_den})
        print(f"step {i}: teacher-forced eps rel {e_eps:.3e}  latents rel {e_lat:.3e}  denoised rel {e_den:.3e}", flush=True)

    # ---- full frame through the graph
    oy, ou, ov = np.empty_like(y), np.empty_like(u), np.empty_like(v)
    eng.infer_yuv420(y, u, v, oy, ou, ov)
    eng.sync()
    rep["init_latents_rel"] = rel(eng.debug_read("init_latents"), ref["init_latents"])
    rep["noisy_rel"] = rel(eng.debug_read("noisy"), ref["noisy_latents"])
    print("init_latents rel", rep["init_latents_rel"], "noisy rel", rep["noisy_rel"], flush=True)
    rep["free_running"] = []
    for i in range(4):
        r = {"step": i, "eps_rel": rel(eng.debug_read("eps", i), ref["eps"][i]),
             "latents_rel": rel(eng.debug_read("latents", i), ref["latents"][i]),
             "denoised_rel": rel(eng.debug_read("denoised", i), ref["denoised_steps"][i])}
        rep["free_running"].append(r)
        print("free-running", r, flush=True)
    img = eng.debug_read("image", channels=4, spatial="image")[:, :3]
    rep["image_rel"] = rel(img if use_kl else img * 2 - 1, ref["image"])   # TAESD's buffer holds the image before its x*2-1 tail
    ref_rgb = ref["rgb"]
    ref_yuv = [imageproc.rgb_to_yuv420(ref_rgb[b]) for b in range(B)]
    p_y = psnr(oy, np.stack([r[0] for r in ref_yuv]))
    p_u = psnr(ou, np.stack([r[1] for r in ref_yuv]))
    p_v = psnr(ov, np.stack([r[2] for r in ref_yuv]))
    # also the RGB path
    rgb_out = np.empty_like(rgb)
    eng.infer_rgb(rgb, rgb_out)
    rep["psnr_rgb"] = psnr(rgb_out, ref_rgb)
    rep["psnr_yuv"] = [p_y, p_u, p_v]
    rep["rgb_max_abs_diff"] = int(np.abs(rgb_out.astype(int) - ref_rgb.astype(int)).max())
    rep["ref_rgb_mean_std"] = [float(ref_rgb.mean()), float(ref_rgb.std())]
    print("image rel", rep["image_rel"], "PSNR rgb", rep["psnr_rgb"], "yuv", rep["psnr_yuv"], "max u8 diff",
          rep["rgb_max_abs_diff"], "ref rgb mean/std", rep["ref_rgb_mean_std"], flush=True)
    # determinism: same frame twice
    oy2, ou2, ov2 = np.empty_like(y), np.empty_like(u), np.empty_like(v)
    eng.infer_yuv420(y, u, v, oy2, ou2, ov2)
    rep["deterministic"] = bool((oy2 == oy).all() and (ou2 == ou).all() and (ov2 == ov).all())
    print("deterministic", rep["deterministic"], flush=True)

    if do_time:
        ty = torch.from_numpy(y).pin_memory(); tu = torch.from_numpy(u).pin_memory(); tv = torch.from_numpy(v).pin_memory()
        py = torch.empty_like(ty).pin_memory(); pu = torch.empty_like(tu).pin_memory(); pv = torch.empty_like(tv).pin_memory()
        for _ in range(3):
            eng.infer_yuv420(ty, tu, tv, py, pu, pv)
        n = 20
        t0 = time.perf_counter()
        for _ in range(n):
            eng.infer_yuv420(ty, tu, tv, py, pu, pv)
        dt = (time.perf_counter() - t0) / n
        rep["e2e_ms_per_batch"] = dt * 1e3
        rep["e2e_fps"] = B / dt
        eng.upload_yuv420(ty, tu, tv); eng.sync()
        t0 = time.perf_counter()
        for _ in range(n):
            eng.run_yuv420()
        eng.sync()
        dt2 = (time.perf_counter() - t0) / n
        rep["device_ms_per_batch"] = dt2 * 1e3
        rep["device_fps"] = B / dt2
        print(f"TIMING e2e {dt*1e3:.2f} ms/batch ({B/dt:.1f} fps)  device-resident {dt2*1e3:.2f} ms/batch ({B/dt2:.1f} fps)", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/pipeline_{H}x{W}x{B}.json", "w") as f:
        json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
