"""BASELINE.json configs 4/5: many concurrent synthetic WebRTC streams over the GPUs of one box.

Every rank (GPU) owns `streams_per_gpu` sessions (videosd_b200.parallel.shard_streams), batches their frames
`batch` at a time (SessionRouter), keeps `lanes` batches in flight (LanePool) and switches each stream's prompt
context every `switch_every` frames among `n_contexts` cached contexts (vsd_set_context re-projects the cross-attention
K/V of the slot; nothing else is recomputed). Reports aggregate frames/s and per-batch latency; rank 0 prints one JSON line.

    python tools/multi_session_sim.py --streams 32 --batch 4 --frames 120            # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/multi_session_sim.py --streams 32 --batch 4
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosd_b200 import weights  # noqa: E402
from videosd_b200.engine import LanePool  # noqa: E402
from videosd_b200.parallel import SessionRouter, aggregate_fps, shard_streams  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=32)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--lanes", type=int, default=2)
    ap.add_argument("--frames", type=int, default=120, help="frames per stream")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--switch-every", type=int, default=60)
    ap.add_argument("--n-contexts", type=int, default=8)
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    H = W = args.size
    B = args.batch
    mine = shard_streams(args.streams, world, rank)            # sessions pinned to this GPU
    router = SessionRouter(num_gpus=1, max_batch=B)
    pool = LanePool(local_rank, args.lanes)
    pool.load_state_dict("unet", weights.random_state_dict(weights.unet_param_shapes(), 1234))
    pool.load_state_dict("vae", weights.random_state_dict(weights.taesd_param_shapes(), 4321))
    pool.configure(B, H, W)
    pool.set_schedule(0.5, 4)
    g = torch.Generator().manual_seed(7)
    contexts = [torch.randn((77, 768), generator=g) for _ in range(args.n_contexts)]
    for b in range(B):
        pool.set_context(b, contexts[0])
    pool.set_reference_noise()
    rs = np.random.RandomState(rank)
    frame = (rs.randint(16, 235, (B, H, W)).astype(np.uint8), rs.randint(16, 240, (B, H // 2, W // 2)).astype(np.uint8),
             rs.randint(16, 240, (B, H // 2, W // 2)).astype(np.uint8))
    pin = [tuple(torch.from_numpy(a).pin_memory() for a in frame) for _ in pool.lanes]
    outs = [tuple(torch.empty_like(t).pin_memory() for t in p) for p in pin]
    batches = router.batches([f"s{s}" for s in mine])[0] if mine else []   # groups of <= B co-located sessions
    # each lane serves a disjoint subset of the session groups; a group's slots hold its sessions' current contexts
    work = [[] for _ in pool.lanes]
    for i, grp in enumerate(batches):
        work[i % len(pool.lanes)].append(grp)
    lat, switches = [], [0]
    lock = threading.Lock()

    def lane_worker(li):
        eng = pool.lanes[li]
        cur = {}   # slot -> context id currently projected
        for f in range(args.frames):
            for grp in work[li]:
                for slot, sess in enumerate(grp):
                    want = (f // args.switch_every + int(sess[1:])) % args.n_contexts   # the stream's prompt at frame f
                    if cur.get(slot) != want:
                        eng.set_context(slot, contexts[want])                           # prompt switch (datachannel, server.py:168-197)
                        cur[slot] = want
                        with lock:
                            switches[0] += 1
                t1 = time.perf_counter()
                eng.infer_yuv420(*pin[li], *outs[li])
                with lock:
                    lat.append((time.perf_counter() - t1) * 1e3)

    for e in pool.lanes:   # warm-up (graph capture)
        e.infer_yuv420(*pin[0], *outs[0])
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    ths = [threading.Thread(target=lane_worker, args=(i,)) for i in range(len(pool.lanes))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    frames_done = sum(len(grp) for lw in work for grp in lw) * args.frames
    fps = aggregate_fps(frames_done, dt, device="cuda" if world > 1 else None)
    if rank == 0:
        print(json.dumps({"config": f"{args.streams} streams {H}x{W}, {world} GPU(s), per-GPU batch {B}, {args.lanes} batches in flight, "
                                    f"context switch every {args.switch_every} frames among {args.n_contexts}",
                          "frames_per_s_aggregate": fps, "streams_per_gpu": len(mine), "fps_per_stream": fps / max(args.streams, 1),
                          "batch_latency_ms_p50": float(np.percentile(lat, 50)), "batch_latency_ms_p95": float(np.percentile(lat, 95)),
                          "context_switches_rank0": switches[0], "seconds": dt}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
