"""Host-side LCM schedule for the engine: timesteps and the per-step scalar coefficients.

Mirrors LCMScheduler_X in the reference (diffusert/lcm/lcm_controlnet.py): scaled-linear betas (:793-815, with
beta_start=0.00085, beta_end=0.012 from :86-95), the strength-aware timestep table (:905-938), the boundary-condition
scalings (:940-946) and the alpha/beta products used by step (:995-1036) and add_noise (:1046-1071). Only scalars are
computed here (torch fp32, same operation order as the reference so they round identically); the tensor math runs in
the CUDA kernels (csrc/bw_kernels.cu: add_noise_kernel, lcm_step_kernel).
"""
import numpy as np
import torch

NUM_TRAIN_TIMESTEPS = 1000
LCM_ORIGIN_STEPS = 50


class LCMSchedule:
    def __init__(self, beta_start=0.00085, beta_end=0.012):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, NUM_TRAIN_TIMESTEPS, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)

    @staticmethod
    def timesteps(strength, num_inference_steps):
        if num_inference_steps > NUM_TRAIN_TIMESTEPS:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than {NUM_TRAIN_TIMESTEPS}")
        c = NUM_TRAIN_TIMESTEPS // LCM_ORIGIN_STEPS
        origin = np.asarray(list(range(1, int(LCM_ORIGIN_STEPS * strength) + 1))) * c - 1
        skip = max(len(origin) // num_inference_steps, 1)
        return [int(t) for t in origin[::-skip][:num_inference_steps]]

    def step_scalars(self, timesteps):
        """-> float32 array [len(timesteps)][6]: sqrt_a, sqrt_1ma, c_skip, c_out, sqrt_a_prev, sqrt_1ma_prev."""
        out = np.zeros((len(timesteps), 6), dtype=np.float32)
        ts = torch.tensor(timesteps, dtype=torch.int64)
        for i in range(len(timesteps)):
            t = ts[i]
            prev_t = ts[i + 1] if i + 1 < len(timesteps) else t
            a_t = self.alphas_cumprod[t]
            a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else torch.tensor(1.0)
            c_skip = 0.5 ** 2 / ((t / 0.1) ** 2 + 0.5 ** 2)
            c_out = (t / 0.1) / ((t / 0.1) ** 2 + 0.5 ** 2) ** 0.5
            out[i] = [float(a_t.sqrt()), float((1 - a_t).sqrt()), float(c_skip), float(c_out), float(a_p.sqrt()),
                      float((1 - a_p).sqrt())]
        return out

    def add_noise_coeffs(self, t0):
        a = self.alphas_cumprod[t0]
        return float(a ** 0.5), float((1 - a) ** 0.5)


def guidance_embedding(guidance_scale=7.5, dim=256):
    """get_w_embedding (lcm_controlnet.py:347-368) for one sample -> float32[dim]."""
    w = torch.tensor([guidance_scale]) * 1000.0
    half = dim // 2
    e = torch.log(torch.tensor(10000.0)) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float32) * -e)
    e = w.to(torch.float32)[:, None] * e[None, :]
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    return e[0].numpy().astype(np.float32)


def reference_cpu_noise(batch, h8, w8, num_timesteps):
    """The noise the reference draws per frame on a CPU device: the CPU global RNG is reset to the
    state of a fresh generator (videopipeline.py:126), then randn(B,4,h,w) for the init noise (lcm_controlnet.py:331)
    and one randn per step inside scheduler.step (:1033). Returned NCHW."""
    g = torch.Generator()
    init = torch.randn((batch, 4, h8, w8), generator=g)
    steps = [torch.randn((batch, 4, h8, w8), generator=g) for _ in range(num_timesteps)] if num_timesteps > 1 else []
    return init, steps
