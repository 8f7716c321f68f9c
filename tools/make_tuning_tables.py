"""Regenerates the committed GEMM tuning tables (videosd_b200/tuning/*.txt) on a B200 (run through gpurun).

For every (height, width, batch, frames in flight) the engine is built with the table loading switched off, the autotuner
times every candidate configuration of every GEMM shape of the plan (UNet + TAESD, and where listed the ControlNet and
AutoencoderKL plans), and its choices are written to gpurun_out/tuning/<H>x<W>x<B>_n<N>.txt. Copy them to
videosd_b200/tuning/ and commit: Engine.configure loads them by default.

    VSD_TUNING_TABLES=0 python tools/make_tuning_tables.py [HxWxB:n[:cn][:kl] ...]
"""
import os
import sys
import time

os.environ["VSD_TUNING_TABLES"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videosd_b200 import weights  # noqa: E402
from videosd_b200.engine import Engine  # noqa: E402

DEFAULT = ["512x512x1:1:cn:kl", "512x512x1:6", "512x512x1:4", "512x512x1:2", "512x512x1:3", "768x768x4:1", "768x768x1:1", "512x512x4:2", "512x512x4:1",
           "512x512x2:1", "512x512x2:2", "512x512x3:2", "512x512x2:3", "512x512x3:3", "512x512x4:3", "360x640x1:1", "256x256x1:1:cn:kl", "256x256x3:1", "256x256x2:1", "256x256x4:1", "128x128x1:1", "64x64x1:1:cn"]


def main():
    specs = [a for a in sys.argv[1:] if a[0].isdigit()] or DEFAULT
    out_dir = os.path.join("gpurun_out", "tuning")
    os.makedirs(out_dir, exist_ok=True)
    need_cn = any(":cn" in s for s in specs)
    need_kl = any(":kl" in s for s in specs)
    for spec in specs:
        parts = spec.split(":")
        H, W, B = (int(v) for v in parts[0].split("x"))
        n = int(parts[1]) if len(parts) > 1 else 1
        t0 = time.time()
        eng = Engine(0)
        eng.set_autotune(n)
        eng.load_state_dict("unet", weights.random_state_dict(weights.unet_param_shapes(), 1234))
        eng.load_state_dict("vae", weights.random_state_dict(weights.taesd_param_shapes(), 4321))
        if ":cn" in spec and need_cn:
            eng.load_state_dict("controlnet", weights.random_state_dict(weights.controlnet_param_shapes(), 9876))
        if ":kl" in spec and need_kl:
            eng.load_state_dict("vae_kl", weights.random_state_dict(weights.autoencoder_kl_param_shapes(), 2222))
        eng.configure(B, H, W)
        eng.set_schedule(0.5, 4)                      # tunes the UNet + TAESD plan
        if ":cn" in spec:
            eng.set_controlnet(True, 1.0)
            eng.set_schedule(0.5, 4)                  # + the ControlNet branch
            eng.set_controlnet(False, 1.0)
        if ":kl" in spec:
            eng.set_vae("kl")
            eng.set_schedule(0.5, 4)                  # + AutoencoderKL encoder / decoder
            eng.set_vae("taesd")
        rep = eng.tuning_report()
        lines = sorted(ln for ln in rep.splitlines() if ln.strip())
        path = os.path.join(out_dir, f"{H}x{W}x{B}_n{n}.txt")
        with open(path, "w") as f:
            f.write("\n".join(lines) + "\n")
        print(f"{path}: {len(lines)} shapes, {eng.tuning_misses()} timed, {time.time() - t0:.1f}s", flush=True)
        eng.close()


if __name__ == "__main__":
    main()
