import os, sys, torch
sys.path.insert(0, os.getcwd())
from videosd_b200 import ops
F = torch.nn.functional
def run(nb, h, w, c, n, poison):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((nb, h, w, c), device="cuda", generator=g).bfloat16()
    wt = (torch.randn((n, 9 * c), device="cuda", generator=g) * (9 * c) ** -0.5).bfloat16()
    if poison:   # leave NaNs behind in shared memory: a GEMM over NaN operands uses the same ring buffers
        xn = torch.full((1, 64, 64, c), float("nan"), device="cuda").bfloat16()
        for _ in range(3):
            ops.conv_gemm(xn, wt, 9)
    y = ops.conv_gemm(x, wt, 9, stride2=True, pad=1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().view(n, 3, 3, c).permute(0, 3, 1, 2), stride=2, padding=1).permute(0, 2, 3, 1)
    bad = ~torch.isfinite(y.float())
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item() if not bad.any() else float("nan")
    rows = bad.any(dim=-1)[0].nonzero()[:6].tolist() if bad.any() else []
    print(f"{h}x{w}x{c}->{n} poison={poison}: err {err:.3e} nonfinite {int(bad.sum())} first bad (h,w): {rows}", flush=True)
for shp in [(1, 45, 80, 320, 320), (1, 44, 80, 320, 320), (1, 23, 40, 640, 640)]:
    for poison in (False, True):
        run(*shp, poison)
