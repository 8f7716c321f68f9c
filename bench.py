"""Benchmark of the per-frame LCM img2img hot path (BASELINE.json: frames/s, 512x512, 4-step LCM img2img, bf16).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference path's CPU restatement (fp32 oracle)

A "step" is one pass of the hot path over one frame batch: YUV420 in -> TAESD encode -> add noise -> 4 x (UNet,
LCM step) -> TAESD decode -> YUV420 out. Prints ONE JSON line (rank 0).
  value : frames/s with the input planes already resident in HBM (CUDA-event timed on the engine's stream)
  e2e   : frames/s through the public call Engine.infer_yuv420 with pinned HOST buffers (H2D + graph + D2H inside)
Multi-GPU is stream/frame-parallel: every rank runs its own replica on its own frames, no collective on the data
path ("scaling": "weak"); torch.distributed (NCCL) is used only for the start/stop barriers and the max-over-ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_FRAME_512 = 3476.8e9        # UNet 4 x 803.27 + TAESD enc 122.32 + dec 141.35 GFLOP
UNET_FLOP_PER_FRAME_512 = 3213.1e9


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs"), "tflops": d.get("bf16_tflops_sustained") or d.get("bf16_tflops"),
                "tflops_burst": d.get("bf16_tflops"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = []
        mx = None
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_frames(n, h, w):
    """n webcam-like YUV420P frames (smooth image translated per frame + seeded noise), pure numpy/torch."""
    frames = []
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    g = torch.Generator().manual_seed(0)
    for k in range(n):
        xs = xx + 3 * k
        rgb = np.stack([0.5 + 0.5 * np.sin(xs / 37.0) * np.cos(yy / 53.0), 0.5 + 0.5 * np.sin((xs + yy) / 71.0),
                        0.5 + 0.5 * np.cos(xs / 29.0 - yy / 41.0)], -1)
        rgb = np.clip((rgb * 0.9 + torch.rand((h, w, 3), generator=g).numpy() * 0.1) * 255, 0, 255).astype(np.int32)
        r, gg, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
        y = ((66 * r + 129 * gg + 25 * b + 128) >> 8) + 16
        m = (rgb.reshape(h // 2, 2, w // 2, 2, 3).sum(axis=(1, 3)) + 2) >> 2
        u = ((-38 * m[..., 0] - 74 * m[..., 1] + 112 * m[..., 2] + 128) >> 8) + 128
        v = ((112 * m[..., 0] - 94 * m[..., 1] - 18 * m[..., 2] + 128) >> 8) + 128
        frames.append((y.astype(np.uint8), u.astype(np.uint8), v.astype(np.uint8)))
    return frames


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# --------------------------------------------------------------------------------------------------- ours
def run_ours(args):
    rank, local_rank, world = dist_env()
    import torch.distributed as dist
    from videosd_b200 import weights
    from videosd_b200.engine import LanePool

    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created: point fd 1 at stderr until that
        # has happened, so that stdout carries exactly one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    H, W, B = args.height, args.width, args.batch
    L = max(1, args.lanes)
    pool = LanePool(local_rank, L)
    pool.load_state_dict("unet", weights.random_state_dict(weights.unet_param_shapes(), 1234))
    pool.load_state_dict("vae", weights.random_state_dict(weights.taesd_param_shapes(), 4321))
    pool.configure(B, H, W)
    # The pool's GEMM configurations are tuned for L frames in flight. The single-lane (latency mode) figure comes from a
    # second engine whose configurations are tuned for ONE frame in flight.
    from videosd_b200.engine import Engine
    solo = None
    if L > 1:
        # its own engine (own weight copy): it is timed alone, after the pool's lanes have drained
        solo = Engine(local_rank)
        solo.load_state_dict("unet", weights.random_state_dict(weights.unet_param_shapes(), 1234))
        solo.load_state_dict("vae", weights.random_state_dict(weights.taesd_param_shapes(), 4321))
        solo.set_autotune(1)
        solo.configure(B, H, W)
    ts = pool.set_schedule(args.strength, args.lcm_steps)
    ctx = torch.randn((B, 77, 768), generator=torch.Generator().manual_seed(7))
    for b in range(B):
        pool.set_context(b, ctx[b])
    pool.set_reference_noise()
    if solo is not None:
        solo.set_schedule(args.strength, args.lcm_steps)
        for b in range(B):
            solo.set_context(b, ctx[b])
        solo.set_reference_noise()
    else:
        solo = pool.lanes[0]
    eng = pool.lanes[0]

    nfr = 8
    frames = synthetic_frames(nfr * B, H, W)
    pinned = []
    for k in range(nfr):
        fs = frames[k * B:(k + 1) * B]
        pinned.append(tuple(torch.from_numpy(np.stack([f[i] for f in fs])).pin_memory() for i in range(3)))
    outs = [(torch.empty((B, H, W), dtype=torch.uint8).pin_memory(),
             torch.empty((B, H // 2, W // 2), dtype=torch.uint8).pin_memory(),
             torch.empty((B, H // 2, W // 2), dtype=torch.uint8).pin_memory()) for _ in range(L)]
    all_lanes = pool.lanes + ([solo] if solo is not pool.lanes[0] else [])
    outs.append(tuple(torch.empty_like(t).pin_memory() for t in outs[0]))
    stream_of = {id(e): torch.cuda.ExternalStream(e.stream, device=local_rank) for e in all_lanes}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, Wm = args.steps, max(args.warmup, 3)

    def device_run(lanes, k_steps):
        """k_steps graph replays spread round-robin over `lanes`, inputs resident in HBM; CUDA-event time (ms)."""
        n = len(lanes)
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        e0.record(stream_of[id(lanes[0])])
        for i in range(1, n):
            stream_of[id(lanes[i])].wait_event(e0)
        for k in range(k_steps):
            lanes[k % n].run_yuv420()
        for i in range(n):
            ends[i].record(stream_of[id(lanes[i])])
        for e in lanes:
            e.sync()
        return max(e0.elapsed_time(ev) for ev in ends)

    def e2e_run(lanes, k_steps):
        """Public call Engine.infer_yuv420 (pinned host planes in and out, synchronous) from one thread per lane."""
        n = len(lanes)
        lat = [[] for _ in range(n)]

        def worker(i):
            for k in range(i, k_steps, n):
                t1 = time.perf_counter()
                lanes[i].infer_yuv420(*pinned[k % nfr], *outs[all_lanes.index(lanes[i])])
                lat[i].append((time.perf_counter() - t1) * 1e3)

        ths = [threading.Thread(target=worker, args=(i,)) for i in range(n)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return (time.perf_counter() - t0) * 1e3, [v for l in lat for v in l]

    # ---- warm-up (>= 3 steps per lane), then the timed regions
    for e in all_lanes:
        e.upload_yuv420(*pinned[0])
        for _ in range(Wm):
            e.run_yuv420()
        e.sync()
    e2e_run(pool.lanes, Wm * L)
    e2e_run([solo], Wm)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms = device_run(pool.lanes, K)                 # value: all lanes
    barrier()
    e2e_ms, lat = e2e_run(pool.lanes, K)               # e2e: all lanes
    barrier()
    dev1_ms = device_run([solo], K)                    # single lane (one frame in flight): latency-optimal mode
    e2e1_ms, lat1 = e2e_run([solo], K)
    clocks = sampler.stop()
    barrier()
    oy = outs[0][0]
    checksum = int(oy.to(torch.int64).sum())  # the device->host result is really read

    times = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = float(times[0]), float(times[1])
    if rank == 0:
        peaks = measured_peaks()
        frames_total = world * K * B
        fps = frames_total / (dev_ms_max / 1e3)
        e2e_fps = frames_total / (e2e_ms_max / 1e3)
        flop_per_frame = FLOP_PER_FRAME_512 * (H * W) / (512.0 * 512.0)  # conv/linear part scales with pixels
        achieved = (fps / world) * flop_per_frame / 1e12                  # per-GPU TFLOP/s over the whole step
        line = {
            "metric": "frames/s (512^2 LCM 4-step img2img)", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"LCM SD1.5 (Dreamshaper-v7 arch, random-init) img2img {H}x{W}, {len(ts)} steps "
                                   f"(timesteps {ts}), strength {args.strength}, TAESD VAE, frame batch {B}, {L} frames in "
                                   f"flight per GPU (lanes sharing one weight copy), YUV420 in/out",
                       "global_batch": world * B, "parallelism": f"frame-parallel x{world} (no collectives)",
                       "frames_in_flight_per_gpu": L,
                       "l2_policy": "no flush: 1.72 GB of UNet weights streamed every pass exceed the 126 MB L2"},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * H * W * 3 // 2,
                    "d2h_bytes_per_step": B * H * W * 3 // 2, "p50_ms": float(np.percentile(lat, 50)),
                    "p95_ms": float(np.percentile(lat, 95)), "checksum": checksum},
            "single_lane": {"value": world * K * B / (dev1_ms / 1e3), "e2e": world * K * B / (e2e1_ms / 1e3),
                            "p50_ms": float(np.percentile(lat1, 50)), "p95_ms": float(np.percentile(lat1, 95)),
                            "note": "one frame in flight per GPU (rank-0 timing)"},
            "gpu_launches": int(eng.launches_per_frame()) * K * world,   # kernel nodes of the frame graph x timed steps of `value`
            "launches_per_frame": int(eng.launches_per_frame()),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops"],
                         "traffic": 12.5e9 if (H, W, B) == (512, 512, 1) else None,   # DRAM bytes per frame, ncu (profiles/r01_summary.md)
                         "note": f"whole-frame algorithmic FLOPs ({flop_per_frame/1e9:.1f} GFLOP/frame) / device time, "
                                 f"of {peaks['source']} sustained bf16 peak; dominant kernel conv_gemm_kernel (tcgen05)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(H, W, args.strength, args.lcm_steps, max_frames=1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------- CPU arm
def cpu_baseline(H, W, strength, lcm_steps, max_frames=1, budget_s=240.0, want_frames=None):
    """Times the fp32 oracle (the CPU restatement of the reference path) on the host cores. The oracle is used
    here only as the baseline being measured, never on the product path."""
    from oracle import imageproc, pipeline
    from oracle.weights import build_taesd, build_unet, random_context

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    unet, vae = build_unet(), build_taesd()
    ctx = random_context(1)
    y, u, v = imageproc.synthetic_frame(H, W)
    t0 = time.perf_counter()
    pipeline.frame_yuv420(unet, vae, y, u, v, ctx, steps=lcm_steps, strength=strength)   # warm-up
    t_warm = time.perf_counter() - t0
    n = max_frames if want_frames is None else want_frames
    n = max(1, min(n, int(budget_s / max(t_warm, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n):
        pipeline.frame_yuv420(unet, vae, y, u, v, ctx, steps=lcm_steps, strength=strength)
    dt = (time.perf_counter() - t0) / n
    return {"value": 1.0 / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} full {H}x{W} {lcm_steps}-step frame(s) after 1 warm-up, fp32 torch on "
                      f"{torch.get_num_threads()} host threads, YUV420 in -> YUV420 out", "s_per_frame": dt}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    H, W = args.height, args.width
    cb = cpu_baseline(H, W, args.strength, args.lcm_steps, want_frames=args.steps, budget_s=240.0)
    fps = cb["value"]
    ts_note = f"{args.lcm_steps} steps, strength {args.strength}"
    line = {
        "impl": "reference", "metric": "frames/s (512^2 LCM 4-step img2img)", "value": fps, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": 1, "ms_per_step": 1e3 / fps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"LCM SD1.5 (Dreamshaper-v7 arch, random-init) img2img {H}x{W}, {ts_note}, TAESD VAE, "
                               f"frame batch 1, YUV420 in/out — CPU restatement of the reference path (diffusers is "
                               f"not installable here)", "global_batch": 1, "parallelism": "host threads"},
        "cpu_baseline": cb,
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=80)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--strength", type=float, default=0.5)
    ap.add_argument("--lcm-steps", type=int, default=4)
    ap.add_argument("--lanes", type=int, default=4, help="frames in flight per GPU (lanes sharing one weight copy)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
