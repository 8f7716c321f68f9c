"""Aggregate throughput of N engines (contexts) sharing one GPU, each running frames back-to-back from its own thread."""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosd_b200 import weights  # noqa: E402
from videosd_b200.engine import LanePool  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
TUNE_FOR = int(sys.argv[2]) if len(sys.argv) > 2 else None   # frames in flight the GEMM autotuner optimises for (default N)
H = W = 512
usd = weights.random_state_dict(weights.unet_param_shapes(), 1234)
vsd = weights.random_state_dict(weights.taesd_param_shapes(), 4321)
ctx = torch.randn((77, 768), generator=torch.Generator().manual_seed(7))
pool = LanePool(0, N, tune_for=TUNE_FOR)
pool.load_state_dict("unet", usd); pool.load_state_dict("vae", vsd)
pool.configure(1, H, W); pool.set_schedule(0.5, 4); pool.set_context(0, ctx); pool.set_reference_noise()
engs = pool.lanes
rs = np.random.RandomState(0)
y = rs.randint(16, 235, (1, H, W)).astype(np.uint8); u = rs.randint(16, 240, (1, H // 2, W // 2)).astype(np.uint8); v = u.copy()
for e in engs:
    e.upload_yuv420(y, u, v); e.run_yuv420(); e.sync()
K = 40


def work(e):
    for _ in range(K):
        e.run_yuv420()
    e.sync()


for n in range(1, N + 1):
    ths = [threading.Thread(target=work, args=(engs[i],)) for i in range(n)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    print(f"{n} concurrent lane(s): {n*K/dt:.1f} fps aggregate, {dt/K*1e3:.2f} ms per frame per lane", flush=True)
