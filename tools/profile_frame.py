"""Sets up the engine with seeded random weights and runs N frames (eager or graph) — the command ncu wraps.
    python tools/profile_frame.py [--frames N] [--eager] [--size HxWxB]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosd_b200 import weights  # noqa: E402
from videosd_b200.engine import Engine  # noqa: E402


def main():
    frames, eager, size = 2, "--eager" in sys.argv, "512x512x1"
    for i, a in enumerate(sys.argv):
        if a == "--frames":
            frames = int(sys.argv[i + 1])
        if a == "--size":
            size = sys.argv[i + 1]
    H, W, B = (int(v) for v in size.split("x"))
    eng = Engine(0)
    eng.load_state_dict("unet", weights.random_state_dict(weights.unet_param_shapes(), 1234))
    eng.load_state_dict("vae", weights.random_state_dict(weights.taesd_param_shapes(), 4321))
    eng.configure(B, H, W)
    cache = f"gpurun_out/tune_cache_{size}.txt"
    if os.path.exists(cache):        # produced by an earlier run outside the profiler
        print("tuning cache entries loaded:", eng.tuning_load(open(cache).read()), flush=True)
    eng.set_schedule(0.5, 4)
    if "--save-tuning" in sys.argv:
        os.makedirs("gpurun_out", exist_ok=True)
        open(cache, "w").write(eng.tuning_report())
    ctx = torch.randn((77, 768), generator=torch.Generator().manual_seed(7))
    for b in range(B):
        eng.set_context(b, ctx)
    eng.set_reference_noise()
    rs = np.random.RandomState(0)
    y = rs.randint(16, 235, (B, H, W)).astype(np.uint8)
    u = rs.randint(16, 240, (B, H // 2, W // 2)).astype(np.uint8)
    v = rs.randint(16, 240, (B, H // 2, W // 2)).astype(np.uint8)
    eng.upload_yuv420(y, u, v)
    eng.sync()
    torch.cuda.synchronize()
    print("SETUP_DONE", flush=True)
    for _ in range(frames):
        if eager:
            eng.debug_run_eager(True)
        else:
            eng.run_yuv420()
            eng.sync()
    print("launches/frame", eng.launches_per_frame())
    for i, a in enumerate(sys.argv):
        if a == "--sections":
            depth = int(sys.argv[i + 1])
            rows = eng.debug_profile_sections(depth, 10)
            tot = sum(r[2] for r in rows)
            print(f"SECTIONS depth={depth} sum={tot/1e3:.3f} ms")
            for t, n, us in rows:
                print(f"  {t:70s} kernels={n:4d} {us:9.2f} us")


if __name__ == "__main__":
    main()
