import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosd_b200 import ops
from videosd_b200._lib import lib, _p
dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
lib().vsd_debug_set_gn_stamps(_p(dbg))
names = ["prologue(param loads)+pdl_wait", "load chunk + block reduce", "grid barrier", "finalize stats", "apply + store"]
for (nb, h, w, c) in [(1, 64, 64, 320), (1, 32, 32, 640), (1, 16, 16, 1280), (1, 8, 8, 1280), (1, 64, 64, 960)]:
    x = torch.randn((nb, h, w, c), device="cuda").bfloat16()
    g = torch.ones(c, device="cuda"); b = torch.zeros(c, device="cuda")
    y = torch.empty_like(x)
    for _ in range(3):
        ops.groupnorm(x, g, b, 32, 1e-5, True, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.groupnorm(x, g, b, 32, 1e-5, True, out=y)
    e1.record(); torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    ph = [d[i + 1] - d[i] for i in range(5)]
    print(f"{nb}x{h}x{w}x{c}: {e0.elapsed_time(e1)*1e3/20:.1f} us/call (eager) | block0 total {d[5]-d[0]} cyc | " + " | ".join(f"{n}={v}" for n, v in zip(names, ph)), flush=True)
