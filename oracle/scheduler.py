"""Oracle restatement of LCMScheduler_X (TEST INFRASTRUCTURE ONLY).

Follows diffusert/lcm/lcm_controlnet.py:
  betas / alphas_cumprod        :793-815   (scaled_linear, beta_start=0.00085, beta_end=0.012 from :86-95)
  set_timesteps                 :905-938   (strength-aware LCM timestep table, lcm_origin_steps=50)
  boundary-condition scalings   :940-946   (sigma_data 0.5, t/0.1)
  step                          :948-1043  (epsilon prediction, no clipping, noise for multi-step)
  add_noise                     :1046-1071
All arithmetic is done with torch fp32 tensors in the same operation order as the reference so that scalar
constants round identically.
"""
import numpy as np
import torch


class LCMSchedulerOracle:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
        self.num_train_timesteps = num_train_timesteps
        # "scaled_linear": linspace in sqrt-space, then squared (lcm_controlnet.py:798-809)
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)  # set_alpha_to_one=True default (:822-824)
        self.timesteps = None

    # lcm_controlnet.py:905-938
    def set_timesteps(self, strength, num_inference_steps, lcm_origin_steps=50):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps larger than the training schedule")
        c = self.num_train_timesteps // lcm_origin_steps
        origin = np.asarray(list(range(1, int(lcm_origin_steps * strength) + 1))) * c - 1
        skipping_step = max(len(origin) // num_inference_steps, 1)
        ts = origin[::-skipping_step][:num_inference_steps]
        self.timesteps = torch.from_numpy(ts.copy().astype(np.int64))
        return self.timesteps

    # lcm_controlnet.py:940-946
    @staticmethod
    def boundary_scalings(t):
        sigma_data = 0.5
        c_skip = sigma_data ** 2 / ((t / 0.1) ** 2 + sigma_data ** 2)
        c_out = (t / 0.1) / ((t / 0.1) ** 2 + sigma_data ** 2) ** 0.5
        return c_skip, c_out

    def step_scalars(self, timeindex):
        """All scalars of one step as 0-d fp32 tensors (what the reference multiplies the tensors by)."""
        t = self.timesteps[timeindex]
        prev_index = timeindex + 1
        prev_t = self.timesteps[prev_index] if prev_index < len(self.timesteps) else t
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        c_skip, c_out = self.boundary_scalings(t)
        return {
            "t": int(t), "prev_t": int(prev_t),
            "sqrt_alpha": a_t.sqrt(), "sqrt_beta": b_t.sqrt(),
            "sqrt_alpha_prev": a_prev.sqrt(), "sqrt_beta_prev": b_prev.sqrt(),
            "c_skip": c_skip, "c_out": c_out,
        }

    # lcm_controlnet.py:948-1043 (epsilon parameterisation). `noise` is what torch.randn(shape) returned.
    def step(self, model_output, timeindex, sample, noise=None):
        s = self.step_scalars(timeindex)
        pred_x0 = (sample - s["sqrt_beta"] * model_output) / s["sqrt_alpha"]
        denoised = s["c_out"] * pred_x0 + s["c_skip"] * sample
        if len(self.timesteps) > 1:
            if noise is None:
                noise = torch.randn(model_output.shape)
            prev_sample = s["sqrt_alpha_prev"] * denoised + s["sqrt_beta_prev"] * noise
        else:
            prev_sample = denoised
        return prev_sample, denoised

    # lcm_controlnet.py:1046-1071
    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.dim() < original_samples.dim():
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original_samples + sb * noise


# lcm_controlnet.py:347-368
def w_embedding(w, embedding_dim=256, dtype=torch.float32):
    assert w.dim() == 1
    w = w * 1000.0
    half = embedding_dim // 2
    e = torch.log(torch.tensor(10000.0)) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=dtype) * -e)
    e = w.to(dtype)[:, None] * e[None, :]
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    if embedding_dim % 2 == 1:
        e = torch.nn.functional.pad(e, (0, 1))
    return e
