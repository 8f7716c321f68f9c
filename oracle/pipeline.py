"""Oracle frame pipeline (TEST INFRASTRUCTURE ONLY): sequencing of one LCM img2img frame as in
diffusert/lcm/lcm_controlnet.py:380-618 with ControlNet residuals = 0, behind the argument handling
of diffusert/videopipeline.py:75-128, with the reference's RNG order on a CPU device:

  per frame: CPU global RNG := state of a fresh torch.Generator()   (videopipeline.py:126)
             init noise  = randn(B,4,h,w)                            (lcm_controlnet.py:331, generator not forwarded)
             step i noise = randn(B,4,h,w) inside scheduler.step     (:1033), drawn on every step incl. the last
"""
import numpy as np
import torch

from . import imageproc
from .scheduler import LCMSchedulerOracle, w_embedding


def frame_noise(batch, h8, w8, num_timesteps):
    """Returns (init_noise, [step noises]) exactly as the reference draws them on a CPU device."""
    g = torch.Generator()  # fresh generator == the state videopipeline.py:126 restores every frame
    init = torch.randn((batch, 4, h8, w8), generator=g)
    steps = []
    if num_timesteps > 1:
        steps = [torch.randn((batch, 4, h8, w8), generator=g) for _ in range(num_timesteps)]
    return init, steps


@torch.no_grad()
def lcm_img2img(unet, vae, rgb_u8, context, steps=4, strength=0.5, guidance_scale=7.5, noise=None, taps=None,
                device="cpu", controlnet=None, controlnet_scale=1.0):
    """rgb_u8: (B,H,W,3) u8 (already cropped/resized to the working size). context: (B,77,768) fp32.
    Returns dict with 'image' (B,3,H,W) fp32, 'rgb' (B,H,W,3) u8, 'latents' (list per step), 'denoised'.
    `device` only moves the fp32 module math (tests run the checker on the GPU for speed); RNG stays on the CPU."""
    rgb_u8 = np.asarray(rgb_u8)
    x = imageproc.preprocess(rgb_u8).to(device)                       # :457
    context = context.to(device)
    B, _, H, W = x.shape
    sched = LCMSchedulerOracle()
    timesteps = sched.set_timesteps(strength, steps, 50)              # :493
    init_latents = vae.encode(x) * vae.scaling_factor                 # :298-303
    h8, w8 = init_latents.shape[-2:]
    if noise is None:
        noise = frame_noise(B, h8, w8, len(timesteps))
    init_noise, step_noise = noise
    init_noise = init_noise.to(device)
    step_noise = [z.to(device) for z in step_noise]
    latents = sched.add_noise(init_latents, init_noise, timesteps[:1].repeat(B))   # :334
    w = torch.tensor(guidance_scale).repeat(B)
    w_emb = w_embedding(w, 256).to(device)                             # :517-520
    out = {"timesteps": timesteps.tolist(), "init_latents": init_latents, "noisy_latents": latents, "latents": [],
           "eps": [], "denoised_steps": [], "latents_in": []}
    control = None
    if controlnet is not None:                                         # videopipeline.py:109, lcm_controlnet.py:459-469
        from .controlnet import control_image_tensor, sobel_edges
        control = torch.cat([control_image_tensor(sobel_edges(rgb_u8[b])) for b in range(B)], 0).to(device)
        out["control"] = control
    denoised = None
    for i, t in enumerate(timesteps):                                  # :532-582
        ts = torch.full((B,), int(t), dtype=torch.long, device=device)
        out["latents_in"].append(latents)
        if controlnet is not None:                                     # :553-566, guess_mode=True, keep = 1.0
            down_res, mid_res = controlnet(latents, ts, context, control, conditioning_scale=controlnet_scale, guess_mode=True)
            eps = unet(latents, ts, w_emb, context, down_res, mid_res)
        else:
            eps = unet(latents, ts, w_emb, context)
        latents, denoised = sched.step(eps, i, latents, step_noise[i] if step_noise else None)
        out["eps"].append(eps)
        out["latents"].append(latents)
        out["denoised_steps"].append(denoised)
        if taps is not None:
            taps(i, eps, latents, denoised)
    image = vae.decode(denoised / vae.scaling_factor)                 # :594-596
    out["denoised"] = denoised
    out["image"] = image
    out["rgb"] = imageproc.postprocess(image)                         # :609-611
    return out


@torch.no_grad()
def frame_yuv420(unet, vae, y, u, v, context, **kw):
    """YUV420P planes in -> YUV420P planes out (the north star's full per-frame path), batch 1."""
    rgb = imageproc.yuv420_to_rgb(y, u, v)
    out = lcm_img2img(unet, vae, rgb[None], context, **kw)
    out["yuv"] = imageproc.rgb_to_yuv420(out["rgb"][0])
    return out
