"""Benchmark of the per-frame LCM img2img hot path (BASELINE.json: frames/s, 512x512, 4-step LCM img2img, bf16).

    python bench.py --gpus N --steps K --warmup W                     # headline: configs[1], 512x512, B = 1, 4 frames in flight
    python bench.py --config c768b4 ...                               # configs[2]: 768x768, frame batch 4
    python bench.py --config sessions --gpus N ...                    # configs[4]: 32 streams over N GPUs, batch 4, prompt switches
    python bench.py --config paced ...                                # configs[1] paced: 300 frames submitted every 33.3 ms
    python bench.py --impl reference --gpus N --steps K ...           # the reference path's CPU restatement (fp32 oracle)

A "step" is one pass of the hot path over one frame batch: YUV420 in -> TAESD encode -> add noise -> 4 x (UNet,
LCM step) -> TAESD decode -> YUV420 out. Prints ONE JSON line (rank 0).
  value : frames/s with the input planes already resident in HBM (CUDA-event timed on the engines' streams)
  e2e   : frames/s through the reference-facing call `VideoSDPipeline.remote(...).infer_yuv420.remote(...)` with HOST
          planes in and out (staging copy, H2D, frame graph, D2H inside the timed region)
Multi-GPU is stream/frame-parallel: every rank runs its own replica on its own frames, no collective on the data
path ("scaling": "weak"); torch.distributed (NCCL) is used only for the start/stop barriers and the max-over-ranks.
"""
import argparse
import hashlib
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
os.environ.setdefault("VSD_WATCHDOG_S", "60")    # a stuck device aborts the bench with a message instead of hanging the driver
os.environ.setdefault("VIDEOSD_NO_RAY", "1")     # the bench reads engine-level counters behind the handle: in-process actors

# SURVEY.md 8(d) / Appendix B, algorithmic FLOPs per frame (2 x MAC; norms, activations and data movement count 0)
FLOP_PER_FRAME = {"c512": 3476.8e9, "c768b4": 9185.8e9}     # UNet 4 passes + TAESD encode + decode
UNET_FLOP_PER_FRAME = {"c512": 3213.1e9, "c768b4": 4 * 8592.5e9 / 4}
PIPE_CFG = dict(model="SimianLuo/LCM_Dreamshaper_v7", controlnet="lllyasviel/control_v11p_sd15_canny", gpus=1, compile=False,
                random_init=True)


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
        try:
            for line in self._proc.stdout:
                parts = [p.strip() for p in line.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
                if self._stop.is_set():
                    break
        except Exception:  # noqa: BLE001
            pass

    def start(self):
        # ONE long-running nvidia-smi in loop mode (-lms 200, the recipe's form): spawning a process per sample costs a full
        # NVML initialisation over all GPUs of the box every 200 ms per rank, which the lane threads of 8 ranks feel
        self._proc = None
        try:
            self._proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                           "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                          stderr=subprocess.DEVNULL, text=True)
        except Exception:  # noqa: BLE001
            return
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if getattr(self, "_proc", None) is not None:
            try:
                self._proc.terminate()
            except Exception:  # noqa: BLE001
                pass
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


def workload(H, W, B, lcm_steps, strength):
    """The configuration both arms run (BASELINE.json configs[1] / [2]); identical text in both arms' JSON lines."""
    return (f"LCM SD1.5 (Dreamshaper-v7 arch, random-init) img2img {H}x{W}, {lcm_steps} steps, strength {strength}, "
            f"TAESD VAE, frame batch {B}, one synthetic webcam stream, YUV420 in/out")


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# --------------------------------------------------------------------------------------------------- ours
def init_dist(local_rank, world):
    import torch.distributed as dist

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
    return dist


def traffic_bytes(config):
    """DRAM bytes per frame (dram__bytes_read.sum + dram__bytes_write.sum over the frame's kernels) from the committed ncu
    launch list of this configuration (profiles/r02_traffic.json), or None when that configuration was not captured."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(config)
    return None


def golden_hash(key):
    p = os.path.join(ROOT, "tests", "golden", "bench_expected.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(key)
    return None


def planes_sha256(planes):
    h = hashlib.sha256()
    for a in planes:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


class Feeder:
    """Keeps `depth` requests outstanding on a pipeline handle and records submit -> result latency per frame."""

    def __init__(self, handle, depth, opts):
        self.handle, self.depth, self.opts = handle, depth, opts

    def run(self, frames, k_steps, interval_s=None):
        lat = []
        lock = threading.Lock()
        sem = threading.Semaphore(self.depth)
        pending = []

        def on_done(t_submit, fut):
            with lock:
                lat.append((time.perf_counter() - t_submit) * 1e3)
            sem.release()

        t0 = time.perf_counter()
        for k in range(k_steps):
            if interval_s is not None:                       # paced source (a 30 fps webcam): never submit early
                wait = t0 + k * interval_s - time.perf_counter()
                if wait > 0:
                    time.sleep(wait)
            sem.acquire()
            ts = time.perf_counter()
            aw = self.handle.infer_yuv420.remote(*frames[k % len(frames)], **self.opts)
            aw._fut.add_done_callback(lambda f, ts=ts: on_done(ts, f))
            pending.append(aw)
        outs = [p.result(timeout=600) for p in pending]
        return (time.perf_counter() - t0) * 1e3, lat, outs


def lane_engines(handle, B, H, W):
    """The engines behind a pipeline handle that are configured for (B, H, W): one per lane that has run a frame."""
    disp = handle._obj.dispatcher
    return [lane.states[(B, H, W)].engine for lane in disp.lanes if (B, H, W) in lane.states]


def run_ours(args):
    rank, local_rank, world = dist_env()
    dist = init_dist(local_rank, world)
    from videosd_b200.videopipeline import VideoSDPipeline

    cfg = args.config
    H, W, B = (768, 768, 4) if cfg == "c768b4" else (args.height, args.width, args.batch)
    L = max(1, args.lanes if cfg != "c768b4" else (args.lanes_c768 or 1))
    K, Wm = args.steps, max(args.warmup, 3)
    opts = dict(strength=args.strength, steps=args.lcm_steps, seed=42, prompt="pixar, cg")
    pipe = VideoSDPipeline.remote(frames_in_flight=L, device=local_rank, **PIPE_CFG)       # the class server.py imports
    solo = VideoSDPipeline.remote(frames_in_flight=1, device=local_rank, **PIPE_CFG) if L > 1 else pipe   # shares the weights

    nfr = 8
    raw = synthetic_frames(nfr * B, H, W)
    frames = [tuple(np.stack([f[i] for f in raw[k * B:(k + 1) * B]]) for i in range(3)) for k in range(nfr)]
    feed, feed1 = Feeder(pipe, L, opts), Feeder(solo, 1, opts)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up through the public call (>= 3 steps per lane: plans, graphs, staging buffers), then the timed regions
    feed.run(frames, Wm * L)
    feed1.run(frames, Wm)
    engines, engines1 = lane_engines(pipe, B, H, W), lane_engines(solo, B, H, W)
    assert len(engines) == L and len(engines1) == 1, (len(engines), len(engines1))
    streams = {id(e): torch.cuda.ExternalStream(e.stream, device=local_rank) for e in engines + engines1}
    ts = engines[0].timesteps

    def device_run(lanes, k_steps):
        """k_steps graph replays spread round-robin over `lanes`, inputs resident in HBM; CUDA-event time (ms)."""
        n = len(lanes)
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        e0.record(streams[id(lanes[0])])
        for i in range(1, n):
            streams[id(lanes[i])].wait_event(e0)
        for k in range(k_steps):
            lanes[k % n].run_yuv420()
        for i in range(n):
            ends[i].record(streams[id(lanes[i])])
        for e in lanes:
            e.sync()
        return max(e0.elapsed_time(ev) for ev in ends)

    for e in engines + engines1:
        for _ in range(2):
            e.run_yuv420()
        e.sync()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms = device_run(engines, K)                       # value: all lanes
    barrier()
    e2e_ms, lat, outs = feed.run(frames, K)               # e2e: the reference-facing call, L requests outstanding
    barrier()
    dev1_ms = device_run(engines1, K)                     # one frame in flight: the per-stream / latency mode
    e2e1_ms, lat1, _ = feed1.run(frames, K)
    paced = None
    if cfg in ("c512", "paced") and B == 1:
        # SURVEY.md 8(d) config 2, latency mode: a 30 fps source, frame k submitted at k x 33.3 ms, submit -> planes on host
        n_paced = 300 if cfg == "paced" else args.paced_frames
        if n_paced > 0:
            _, lat_p, _ = feed1.run(frames, n_paced, interval_s=1.0 / 30.0)
            paced = {"frames": n_paced, "interval_ms": 33.3, "p50_ms": float(np.percentile(lat_p, 50)),
                     "p95_ms": float(np.percentile(lat_p, 95)), "max_ms": float(np.max(lat_p)),
                     "note": "one stream, frames submitted every 33.3 ms through VideoSDPipeline.infer_yuv420.remote; "
                             "latency = submit -> output planes on the host"}
    clocks = sampler.stop()
    barrier()
    sha = planes_sha256(outs[0])                           # frame 0 of the timed e2e run: the D2H result is really read
    gkey = f"{H}x{W}x{B}_n{L}"
    want = golden_hash(gkey)
    misses = sum(e.tuning_misses() for e in engines)

    times = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    ok = torch.tensor([1.0 if (want is None or want == sha) else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dev_ms_max, e2e_ms_max = float(times[0]), float(times[1])
    if rank == 0:
        peaks = measured_peaks()
        frames_total = world * K * B
        fps = frames_total / (dev_ms_max / 1e3)
        e2e_fps = frames_total / (e2e_ms_max / 1e3)
        ckey = "c768b4" if (H, W) == (768, 768) else "c512"
        scale = (H * W) / (768.0 * 768.0 if ckey == "c768b4" else 512.0 * 512.0)
        flop_per_frame = FLOP_PER_FRAME[ckey] * scale
        unet_flop = UNET_FLOP_PER_FRAME[ckey] * scale
        achieved = (fps / world) * flop_per_frame / 1e12                  # per-GPU TFLOP/s over the whole step
        lpf = int(engines[0].launches_per_frame())
        line = {
            "metric": f"frames/s ({H}x{W} LCM {len(ts)}-step img2img)", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload(H, W, B, args.lcm_steps, args.strength), "name": cfg, "global_batch": B},
            "impl_detail": {"timesteps": ts, "frames_in_flight_per_gpu": L, "parallelism": f"frame-parallel x{world} (no collectives)",
                            "note": f"{L} lanes sharing one weight copy keep {L} frames of the stream in flight per GPU",
                            "l2_policy": "no flush: 1.72 GB of UNet weights streamed every pass exceed the 126 MB L2"},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * H * W * 3 // 2,
                    "d2h_bytes_per_step": B * H * W * 3 // 2, "p50_ms": float(np.percentile(lat, 50)),
                    "p95_ms": float(np.percentile(lat, 95)),
                    "api": "VideoSDPipeline.remote(frames_in_flight=%d).infer_yuv420.remote(y, u, v, ...)" % L,
                    "output_sha256": sha, "output_matches_golden": (None if want is None else bool(ok[0] > 0.5)),
                    "golden_key": gkey},
            "single_lane": {"value": world * K * B / (dev1_ms / 1e3), "e2e": world * K * B / (e2e1_ms / 1e3),
                            "p50_ms": float(np.percentile(lat1, 50)), "p95_ms": float(np.percentile(lat1, 95)),
                            "roofline_frac": (K * B / (dev1_ms / 1e3)) * flop_per_frame / 1e12 / peaks["tflops"],
                            "note": "one frame in flight per GPU (rank-0 timing)"},
            "gpu_launches": lpf * K * world,   # kernel nodes of the frame graph x timed steps of `value`
            "launches_per_frame": lpf,
            "tuning": {"table_misses": misses, "note": "GEMM configurations come from the committed videosd_b200/tuning tables; "
                                                       "a miss is timed on the device at plan build"},
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops"], "traffic": traffic_bytes(cfg if cfg != "paced" else "c512"),
                         "unet_frac": (fps / world) * unet_flop / 1e12 / peaks["tflops"],
                         "note": f"whole-frame algorithmic FLOPs ({flop_per_frame/1e9:.1f} GFLOP/frame; UNet only "
                                 f"{unet_flop/1e9:.1f}) / device time, of {peaks['source']} sustained bf16 peak; dominant kernel "
                                 f"conv_gemm_kernel (tcgen05); traffic = DRAM bytes per frame from the committed ncu launch list"},
        }
        if paced is not None:
            line["paced_30fps"] = paced
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(H, W, args.strength, args.lcm_steps, max_frames=1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_sessions(args):
    """BASELINE.json configs[4]: `--streams` concurrent synthetic WebRTC sessions over the GPUs of the box, sessions pinned
    per GPU (parallel.shard_streams), per-GPU batching of <= 4 by the product dispatcher, every stream switching its prompt
    context every `--switch-every` frames among 8 contexts. One frame in flight per session (the reference's recv() never
    queues a second frame behind a running one, server.py:132-137)."""
    rank, local_rank, world = dist_env()
    dist = init_dist(local_rank, world)
    from videosd_b200.parallel import shard_streams
    from videosd_b200.videopipeline import VideoSDPipeline

    H, W = args.height, args.width
    g = torch.Generator().manual_seed(7)
    contexts = {f"context {k}": torch.randn((77, 768), generator=g) for k in range(args.n_contexts)}
    pipe = VideoSDPipeline.remote(frames_in_flight=args.session_lanes, max_batch=args.max_batch, device=local_rank,
                                  prompt_encoder=lambda p: contexts[p if isinstance(p, str) else p[0]], **PIPE_CFG)
    mine = shard_streams(args.streams, world, rank)
    raw = synthetic_frames(max(len(mine), 1), H, W)
    K = args.steps
    lat = []
    lock = threading.Lock()

    def session(j, sid, n_frames):
        y, u, v = raw[j]
        for f in range(n_frames):
            prompt = f"context {(f // args.switch_every + sid) % args.n_contexts}"   # datachannel prompt update, server.py:168-197
            t1 = time.perf_counter()
            pipe.infer_yuv420.remote(y, u, v, strength=args.strength, steps=args.lcm_steps, seed=42 + sid, prompt=prompt).result(timeout=900)
            with lock:
                lat.append((time.perf_counter() - t1) * 1e3)

    def run_all(n_frames):
        ths = [threading.Thread(target=session, args=(j, sid, n_frames)) for j, sid in enumerate(mine)]
        t0 = time.perf_counter()
        [t.start() for t in ths]
        [t.join() for t in ths]
        return time.perf_counter() - t0

    run_all(max(args.warmup, 3))
    lat.clear()
    disp = pipe._obj.dispatcher
    disp.stats.clear()
    sampler = ClockSampler(local_rank)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.start()
    dt = run_all(K)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    t = torch.tensor([dt, float(len(mine) * K)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt_max, frames_total = float(tm[0]), float(t[1])
    else:
        dt_max, frames_total = dt, float(t[1])
    if rank == 0:
        peaks = measured_peaks()
        fps = frames_total / dt_max
        flop = FLOP_PER_FRAME["c512"] * (H * W) / (512.0 * 512.0)
        st = disp.stats
        line = {
            "metric": f"frames/s ({H}x{W} LCM {args.lcm_steps}-step img2img), {args.streams} concurrent streams", "value": fps,
            "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt_max / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.streams} concurrent synthetic streams {H}x{W} over {world} GPU(s), sessions pinned per "
                                   f"GPU, per-GPU frame batching <= {args.max_batch} x {args.session_lanes} batches in flight, prompt "
                                   f"context switch every {args.switch_every} frames among {args.n_contexts} contexts, "
                                   f"{K} frames per stream, YUV420 in/out through VideoSDPipeline.infer_yuv420.remote",
                       "name": "sessions", "global_batch": args.streams, "parallelism": f"stream-parallel x{world} (no collectives)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": H * W * 3 // 2, "d2h_bytes_per_step": H * W * 3 // 2,
                    "p50_ms": float(np.percentile(lat, 50)), "p95_ms": float(np.percentile(lat, 95)),
                    "note": "value == e2e here: the timed region is wall clock over all session threads (host planes in and out)"},
            "fps_per_stream": fps / max(args.streams, 1),
            "rank0_dispatcher": {"launches": st["launches"], "frames": st["frames"], "frames_merged_into_batches": st["merged"],
                                 "mean_batch": st["frames"] / max(st["launches"], 1), "context_switches": st["context_switches"]},
            "gpu_launches": int(st["launches"]) * world * int(lane_engines_any(pipe)),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": (fps / world) * flop / 1e12, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": (fps / world) * flop / 1e12 / peaks["tflops"], "traffic": None},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def lane_engines_any(handle):
    """Kernel nodes per launch of the most used engine behind `handle` (for the gpu_launches claim of the sessions mode)."""
    best = 0
    for lane in handle._obj.dispatcher.lanes:
        for st in lane.states.values():
            best = max(best, int(st.engine.launches_per_frame()))
    return best


# --------------------------------------------------------------------------------------------------- CPU arm
def cpu_baseline(H, W, strength, lcm_steps, max_frames=1, budget_s=240.0, want_frames=None, want_warmup=1):
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
    warm = 1
    while warm < want_warmup and (warm + 1) * t_warm < 0.25 * budget_s:                   # further warm-up frames if they are cheap
        pipeline.frame_yuv420(unet, vae, y, u, v, ctx, steps=lcm_steps, strength=strength)
        warm += 1
    n = max_frames if want_frames is None else want_frames
    n = max(1, min(n, int(0.75 * budget_s / max(t_warm, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n):
        pipeline.frame_yuv420(unet, vae, y, u, v, ctx, steps=lcm_steps, strength=strength)
    dt = (time.perf_counter() - t0) / n
    return {"value": 1.0 / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} full {H}x{W} {lcm_steps}-step frame(s) after {warm} warm-up, fp32 torch on "
                      f"{torch.get_num_threads()} host threads, YUV420 in -> YUV420 out", "s_per_frame": dt,
            "frames": n, "warmup_frames": warm}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    H, W, B = (768, 768, 4) if args.config == "c768b4" else (args.height, args.width, args.batch)
    cb = cpu_baseline(H, W, args.strength, args.lcm_steps, want_frames=args.steps, budget_s=240.0, want_warmup=args.warmup)
    fps = cb["value"]
    line = {
        "impl": "reference", "metric": f"frames/s ({H}x{W} LCM {args.lcm_steps}-step img2img)", "value": fps, "unit": "frames/s",
        "n_gpus": world, "steps": cb["frames"], "warmup": cb["warmup_frames"], "ms_per_step": 1e3 / fps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(H, W, B, args.lcm_steps, args.strength), "name": args.config, "global_batch": B},
        "impl_detail": {"parallelism": "host threads", "note": "CPU restatement of the reference path, one frame at a time "
                        "(diffusers is not installable here); frames/s of single frames"},
        "cpu_baseline": cb,
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (frame batches; frames per stream in the sessions config)")
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c512", choices=["c512", "c768b4", "sessions", "paced"])
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--strength", type=float, default=0.5)
    ap.add_argument("--lcm-steps", type=int, default=4)
    ap.add_argument("--lanes", type=int, default=6, help="frames in flight per GPU (lanes sharing one weight copy); 6 keeps p50 <= 50 ms")
    ap.add_argument("--lanes-c768", type=int, default=0, help="batches in flight for --config c768b4 (default 1)")
    ap.add_argument("--paced-frames", type=int, default=90, help="frames of the paced 30 fps latency run inside the default config")
    ap.add_argument("--streams", type=int, default=32)
    ap.add_argument("--max-batch", type=int, default=4)
    ap.add_argument("--session-lanes", type=int, default=3, help="batches in flight per GPU in the sessions config (2: 146 fps, 3: 162, 4: 158 at 32 streams)")
    ap.add_argument("--switch-every", type=int, default=60)
    ap.add_argument("--n-contexts", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {"c512": 160, "paced": 160, "c768b4": 40, "sessions": 120}[args.config] if args.impl == "ours" else 1
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "sessions":
        run_sessions(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
