"""-m gpu: parity of the configurations bench.py actually measures.

* frames in flight: six different frames run concurrently on six lanes (bench.py's default; one weight copy) through the reference-facing
  class; every output must be bit-equal to the same (frame, options) run alone with the same GEMM configurations, 50
  rounds in a row (shared split-K workspaces, cluster GroupNorm, concurrent graphs), and the alone-run must meet the
  oracle tolerance -- so the headline mode is covered by the same bar as the single-lane tests;
* BASELINE config 3: 768x768, frame batch 4, four contexts: per-step latents <= 1e-2 (teacher-forced), PSNR >= 40 dB;
* sessions merged into one batched launch by the dispatcher get the result they would get alone.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(model="SimianLuo/LCM_Dreamshaper_v7", controlnet="lllyasviel/control_v11p_sd15_canny", gpus=1, compile=False,
           random_init=True, device=0)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def psnr(a, b):
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def _frames(n, H, W):
    from oracle import imageproc

    return [imageproc.synthetic_frame(H, W, seed=40 + i, shift=11 * i) for i in range(n)]


def _same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


LANES = 6   # bench.py --lanes default: the configuration `value` / `e2e` are measured in


def test_frames_in_flight_bit_equal_to_alone_and_within_oracle_tolerance(oracle_models):
    from oracle import imageproc, pipeline
    from oracle.weights import random_context
    from videosd_b200.videopipeline import VideoSDPipeline

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    unet, vae = oracle_models
    H = W = 512
    cfg = dict(CFG, state_dicts={"unet": unet.state_dict(), "vae": vae.state_dict()})
    pipe = VideoSDPipeline.remote(frames_in_flight=LANES, noise_mode="reference_cpu", **cfg)
    frames = _frames(LANES, H, W)
    ctx = random_context(LANES, seed=3)
    opts = [dict(strength=0.5, steps=4, seed=42, prompt_embeds=ctx[i:i + 1]) for i in range(LANES)]
    alone = [pipe.infer_yuv420.remote(*frames[i], **opts[i]).result(timeout=900) for i in range(LANES)]
    assert not _same(alone[0], alone[1])
    for rnd in range(50):
        futs = [pipe.infer_yuv420.remote(*frames[i], **opts[i]) for i in range(LANES)]
        outs = [f.result(timeout=900) for f in futs]
        for i in range(LANES):
            assert _same(outs[i], alone[i]), (rnd, i)
    disp = pipe._obj.dispatcher
    assert sum(1 for lane in disp.lanes if lane.states) == LANES and disp.stats["frames"] == 51 * LANES    # every lane really ran
    # the frames the lanes produce meet the north star's bar against the fp32 oracle on the same weights / inputs / noise
    try:
        ug, vg = unet.cuda(), vae.cuda()
        for i in (1, 3):
            rgb = imageproc.yuv420_to_rgb(*frames[i])[None]
            with torch.no_grad():
                ref = pipeline.lcm_img2img(ug, vg, rgb, ctx[i:i + 1], steps=4, strength=0.5, device="cuda")
            out = pipe.infer_yuv420.remote(*frames[i], **opts[i]).result(timeout=900)
            assert _same(out, alone[i])
            eng = disp.last_state.engine
            assert rel(eng.debug_read("init_latents"), ref["init_latents"]) < 1e-2
            for k in range(4):
                assert rel(eng.debug_read("latents", k), ref["latents"][k]) < 2e-2, (i, k)
            ry, ru, rv = imageproc.rgb_to_yuv420(ref["rgb"][0])
            assert psnr(out[0][0], ry) >= 40.0 and psnr(out[1][0], ru) >= 40.0 and psnr(out[2][0], rv) >= 40.0
    finally:
        unet.cpu(); vae.cpu()


def test_lanes_with_controlnet_bit_equal_to_alone():
    """ControlNet branch (side stream, its own split-K workspace) with two frames in flight (LanePool + ControlNet)."""
    from videosd_b200.videopipeline import VideoSDPipeline

    H = W = 256
    pipe = VideoSDPipeline.remote(frames_in_flight=2, use_controlnet=True, **CFG)
    frames = _frames(2, H, W)
    opts = [dict(strength=0.5, steps=4, seed=7 + i, prompt=f"prompt {i}", controlnet_scale=0.8) for i in range(2)]
    alone = [pipe.infer_yuv420.remote(*frames[i], **opts[i]).result(timeout=900) for i in range(2)]
    for rnd in range(20):
        futs = [pipe.infer_yuv420.remote(*frames[i], **opts[i]) for i in range(2)]
        for i, f in enumerate(futs):
            assert _same(f.result(timeout=900), alone[i]), (rnd, i)
    weak = pipe.infer_yuv420.remote(*frames[0], **dict(opts[0], controlnet_scale=0.1)).result(timeout=900)
    assert not _same(weak, alone[0])                                  # the conditioning scale reaches the lanes


def test_sessions_merged_into_a_batch_get_their_own_frame():
    """8 sessions, one lane, max_batch 4: requests that wait are merged into batched launches; each session's output equals
    the output of its (frame, prompt, seed) submitted alone through the SAME batch size (slot 0 of a 1-frame launch uses
    other GEMM tiles than slot 2 of a 4-frame launch, so bit-equality is asserted per batch geometry, parity across them)."""
    import threading

    from videosd_b200.videopipeline import VideoSDPipeline

    H = W = 256
    pipe = VideoSDPipeline.remote(frames_in_flight=1, max_batch=4, **CFG)
    frames = _frames(8, H, W)
    opts = [dict(strength=0.5, steps=4, seed=100 + (i % 3), prompt=f"session prompt {i % 5}") for i in range(8)]
    alone = [pipe.infer_yuv420.remote(*frames[i], **opts[i]).result(timeout=900) for i in range(8)]
    results = [[] for _ in range(8)]

    def session(i):
        for _ in range(6):
            results[i].append(pipe.infer_yuv420.remote(*frames[i], **opts[i]).result(timeout=900))

    ths = [threading.Thread(target=session, args=(i,)) for i in range(8)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    disp = pipe._obj.dispatcher
    assert disp.stats["merged"] > 0                                   # batching really happened
    for i in range(8):
        for out in results[i]:
            # same session, any batch: within one bf16 rounding of the alone-run everywhere (PSNR >> 40 dB), and its own frame
            assert psnr(out[0], alone[i][0]) >= 45.0, i
            assert all(psnr(out[0], alone[j][0]) < psnr(out[0], alone[i][0]) for j in range(8) if j != i)


def test_sessions_on_two_lanes_with_batches_and_context_switches_complete():
    """BASELINE config 5 in small: 12 sessions over two lanes, batches of <= 4, every session switching its prompt every
    third frame. Regression test of the device hang of round 2 (a CTA-pair tcgen05.alloc issued before the peer CTA had
    started never returned; it needed several batch engines and per-slot context projections in flight to show): the run
    must complete under the device watchdog (conftest: VSD_WATCHDOG_S), and a session's frames must still be its own."""
    import threading

    from videosd_b200.videopipeline import VideoSDPipeline

    H = W = 256
    pipe = VideoSDPipeline.remote(frames_in_flight=2, max_batch=4, **CFG)
    frames = _frames(12, H, W)
    alone = [pipe.infer_yuv420.remote(*frames[i], strength=0.5, steps=4, seed=7, prompt="session prompt 0").result(timeout=900)
             for i in range(12)]
    results = [[] for _ in range(12)]

    def session(i):
        for f in range(12):
            results[i].append(pipe.infer_yuv420.remote(*frames[i], strength=0.5, steps=4, seed=7,
                                                       prompt=f"session prompt {(f // 3 + i) % 4}").result(timeout=900))

    ths = [threading.Thread(target=session, args=(i,)) for i in range(12)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    st = pipe._obj.dispatcher.stats
    assert st["frames"] == 12 + 12 * 12 and st["merged"] > 0 and st["context_switches"] > 12
    for i in range(12):
        assert len(results[i]) == 12
        for f in (0, 1, 2):          # the frames of the first prompt period of session 0 ... those with prompt 0: compare with `alone`
            if (f // 3 + i) % 4 == 0:
                assert psnr(results[i][f][0], alone[i][0]) >= 45.0, (i, f)


def test_config3_768x768_batch4_per_step_latents_and_psnr(oracle_models):
    """BASELINE.json configs[2]: 768x768, frame batch 4 (latent 4x4x96x96, 9216-key self-attention), four contexts."""
    from oracle import imageproc, pipeline
    from oracle.scheduler import LCMSchedulerOracle
    from oracle.weights import random_context
    from videosd_b200.engine import Engine

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    unet, vae = oracle_models
    H = W = 768
    B = 4
    eng = Engine(0)
    try:
        eng.load_state_dict("unet", unet.state_dict())
        eng.load_state_dict("vae", vae.state_dict())
        ug, vg = unet.cuda(), vae.cuda()
        ctx = random_context(B, seed=11)
        eng.configure(B, H, W)
        assert eng.set_schedule(0.5, 4) == [499, 379, 259, 139]
        for b in range(B):
            eng.set_context(b, ctx[b])
        eng.set_reference_noise()
        frames = [imageproc.synthetic_frame(H, W, seed=b, shift=13 * b) for b in range(B)]
        y, u, v = (np.stack([f[i] for f in frames]) for i in range(3))
        rgb = np.stack([imageproc.yuv420_to_rgb(*f) for f in frames])
        with torch.no_grad():
            ref = pipeline.lcm_img2img(ug, vg, rgb, ctx, steps=4, strength=0.5, device="cuda")
        oy, ou, ov = np.empty_like(y), np.empty_like(u), np.empty_like(v)
        eng.infer_yuv420(y, u, v, oy, ou, ov)
        eng.sync()
        assert rel(eng.debug_read("init_latents"), ref["init_latents"]) < 1e-2
        for i in range(4):
            assert rel(eng.debug_read("latents", i), ref["latents"][i]) < 2e-2, i      # free-running
        ref_yuv = [imageproc.rgb_to_yuv420(ref["rgb"][b]) for b in range(B)]
        for k, plane in enumerate((oy, ou, ov)):
            assert psnr(plane, np.stack([r[k] for r in ref_yuv])) >= 40.0
        sched = LCMSchedulerOracle(); sched.set_timesteps(0.5, 4)
        _, step_noise = pipeline.frame_noise(B, H // 8, W // 8, 4)
        for i in range(4):                                                             # teacher-forced: <= 1e-2 per step
            lat_in = ref["latents_in"][i].cpu()
            eps = eng.debug_unet(lat_in, i)
            lat, _ = sched.step(eps, i, lat_in, step_noise[i])
            assert rel(lat, ref["latents"][i]) <= 1e-2, (i, rel(lat, ref["latents"][i]))
    finally:
        eng.close()
        unet.cpu(); vae.cpu()
