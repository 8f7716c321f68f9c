"""Generates tests/golden/reference_golden.npz by running THE REFERENCE'S OWN CODE in this container.

The reference module diffusert/lcm/lcm_controlnet.py (read from /root/reference, never copied) cannot be imported
as is because `diffusers` is not installed. This script installs a stub `diffusers` namespace that supplies only
the base classes / helpers the file imports, then:

  1. drives the reference's live scheduler `LCMScheduler_X` (set_timesteps / step / add_noise) and the pipeline's
     `get_w_embedding` on seeded inputs                                 -> pins oracle/scheduler.py
  2. runs the reference's own `LatentConsistencyModelPipeline_controlnet.__call__` control flow (prompt_embeds path,
     ControlNet stub returning zero residuals) over the oracle's restated UNet / TAESD modules with
     the RNG reset of diffusert/videopipeline.py:126                      -> pins oracle/pipeline.py sequencing + RNG order
  3. runs the reference's own `SobelOperator` (lcm/canny_gpu.py, importable as is) and the same `__call__` with the
     oracle's restated ControlNetModel plugged in (conditioning scale 0.7, guess mode, per-step keep)
                                                                          -> pins oracle/controlnet.py sobel_edges + the ControlNet call sequencing

What this does NOT pin: the arithmetic inside UNet2DConditionModel / AutoencoderTiny / VaeImageProcessor, which lives in
the absent third-party package (those restatements are checked by parameter-count and key-name identities only).

Run (only where /root/reference exists):  python tests/golden/make_reference_golden.py
"""
import dataclasses
import importlib.util
import os
import sys
import types

import numpy as np
import PIL.Image
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/diffusert/lcm/lcm_controlnet.py"

from oracle import imageproc  # noqa: E402
from oracle.weights import build_taesd, build_unet, random_context  # noqa: E402


def install_stub_diffusers():
    import inspect

    d = types.ModuleType("diffusers")

    class ConfigMixin:
        pass

    class SchedulerMixin:
        pass

    class DiffusionPipeline:
        def __init__(self):
            pass

        def register_modules(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

        @property
        def _execution_device(self):
            return torch.device("cpu")

    class _Base(torch.nn.Module):
        pass

    class ControlNetModel(_Base):
        """Zero-residual stand-in: the hot path's parity target is the residual-free UNet."""
        config = types.SimpleNamespace(global_pool_conditions=False)
        dtype = torch.float32

        def forward(self, sample, ts, encoder_hidden_states=None, controlnet_cond=None, conditioning_scale=1.0,
                    guess_mode=False, return_dict=False):
            return [None] * 12, None

    class logging:  # noqa: N801
        @staticmethod
        def get_logger(name):
            import logging as _l

            return _l.getLogger(name)

    d.AutoencoderKL = type("AutoencoderKL", (_Base,), {})
    d.UNet2DConditionModel = type("UNet2DConditionModel", (_Base,), {})
    d.ConfigMixin, d.SchedulerMixin, d.DiffusionPipeline, d.ControlNetModel, d.logging = (
        ConfigMixin, SchedulerMixin, DiffusionPipeline, ControlNetModel, logging)

    cu = types.ModuleType("diffusers.configuration_utils")

    def register_to_config(init):
        sig = inspect.signature(init)

        def wrapped(self, *a, **kw):
            bound = sig.bind(self, *a, **kw)
            bound.apply_defaults()
            cfg = dict(bound.arguments)
            cfg.pop("self")
            self.config = types.SimpleNamespace(**cfg)
            init(self, *a, **kw)

        return wrapped

    cu.register_to_config = register_to_config

    ip = types.ModuleType("diffusers.image_processor")

    class VaeImageProcessor:
        """The reference calls this diffusers class; the stub routes to the oracle's restatement (Appendix D.1/D.2)."""

        def __init__(self, vae_scale_factor=8, do_convert_rgb=False, do_normalize=True):
            self.do_normalize = do_normalize

        def preprocess(self, image, height=None, width=None):
            arr = np.array(image.convert("RGB"))
            x = imageproc.preprocess(arr)
            return x if self.do_normalize else (x + 1) / 2

        def postprocess(self, image, output_type="pil", do_denormalize=None):
            u8 = imageproc.postprocess(image)
            return [PIL.Image.fromarray(a) for a in u8]

    ip.VaeImageProcessor = VaeImageProcessor
    ip.PipelineImageInput = object

    sd = types.ModuleType("diffusers.pipelines.stable_diffusion")

    @dataclasses.dataclass
    class StableDiffusionPipelineOutput:
        images: list
        nsfw_content_detected: object

    sd.StableDiffusionPipelineOutput = StableDiffusionPipelineOutput
    sc = types.ModuleType("diffusers.pipelines.stable_diffusion.safety_checker")
    sc.StableDiffusionSafetyChecker = type("StableDiffusionSafetyChecker", (), {})
    ut = types.ModuleType("diffusers.utils")

    class BaseOutput:
        pass

    ut.BaseOutput = BaseOutput
    tu = types.ModuleType("diffusers.utils.torch_utils")
    tu.randn_tensor = lambda shape, generator=None, device=None, dtype=None: torch.randn(
        shape, generator=generator, device=device, dtype=dtype)
    tu.is_compiled_module = lambda m: False
    mc = types.ModuleType("diffusers.pipelines.controlnet.multicontrolnet")
    mc.MultiControlNetModel = type("MultiControlNetModel", (), {})
    mods = {"diffusers": d, "diffusers.configuration_utils": cu, "diffusers.image_processor": ip,
            "diffusers.pipelines": types.ModuleType("diffusers.pipelines"),
            "diffusers.pipelines.stable_diffusion": sd, "diffusers.pipelines.stable_diffusion.safety_checker": sc,
            "diffusers.utils": ut, "diffusers.utils.torch_utils": tu,
            "diffusers.pipelines.controlnet": types.ModuleType("diffusers.pipelines.controlnet"),
            "diffusers.pipelines.controlnet.multicontrolnet": mc}
    sys.modules.update(mods)
    return d


def load_reference():
    install_stub_diffusers()
    spec = importlib.util.spec_from_file_location("ref_lcm_controlnet", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class UNetAdapter(torch.nn.Module):
    """Gives the oracle UNet the diffusers call signature the reference uses (lcm_controlnet.py:568-577)."""
    dtype = torch.float32
    config = types.SimpleNamespace(in_channels=4, sample_size=96)

    def __init__(self, net):
        super().__init__()
        self.net = net
        self.calls = []

    def forward(self, sample, ts, timestep_cond=None, encoder_hidden_states=None, cross_attention_kwargs=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None, return_dict=False):
        assert all(r is None for r in down_block_additional_residuals) and mid_block_additional_residual is None
        out = self.net(sample, ts, timestep_cond, encoder_hidden_states)
        self.calls.append((sample.clone(), ts.clone(), out.clone()))
        return (out,)


class VaeAdapter(torch.nn.Module):
    config = types.SimpleNamespace(scaling_factor=1.0, block_out_channels=(64, 64, 64, 64))

    def __init__(self, net):
        super().__init__()
        self.net = net

    def encode(self, x):
        return types.SimpleNamespace(latents=self.net.encode(x))

    def decode(self, z, return_dict=False):
        return (self.net.decode(z),)


def main():
    ref = load_reference()
    out = {}
    # ---------------------------------------------------------------- 1. scheduler + w-embedding
    sch = ref.LCMScheduler_X(beta_start=0.00085, beta_end=0.0120, beta_schedule="scaled_linear", prediction_type="epsilon")
    out["alphas_cumprod"] = sch.alphas_cumprod.numpy()
    tables = [(0.5, 4), (0.6, 4), (0.8, 4), (1.0, 4), (0.4, 20), (0.5, 1), (0.5, 12), (0.05, 4), (0.1, 12)]
    out["table_cfg"] = np.array(tables, dtype=np.float64)
    for k, (st, n) in enumerate(tables):
        sch.set_timesteps(st, n, 50)
        out[f"table_{k}"] = sch.timesteps.numpy()
    g = torch.Generator().manual_seed(123)
    sample = torch.randn((2, 4, 8, 8), generator=g)
    model_out = torch.randn((2, 4, 8, 8), generator=g)
    out["step_sample"], out["step_model_out"] = sample.numpy(), model_out.numpy()
    sch.set_timesteps(0.5, 4, 50)
    for i, t in enumerate(sch.timesteps):
        torch.manual_seed(1000 + i)  # the reference draws torch.randn(shape) from the global CPU RNG (:1033)
        prev, den = sch.step(model_out, i, t, sample, return_dict=False)
        out[f"step_prev_{i}"], out[f"step_den_{i}"] = prev.numpy(), den.numpy()
    sch.set_timesteps(0.5, 1, 50)
    prev, den = sch.step(model_out, 0, sch.timesteps[0], sample, return_dict=False)
    out["single_prev"], out["single_den"] = prev.numpy(), den.numpy()
    sch.set_timesteps(0.5, 4, 50)
    noise = torch.randn((2, 4, 8, 8), generator=g)
    out["add_noise_noise"] = noise.numpy()
    out["add_noise_out"] = sch.add_noise(sample, noise, sch.timesteps[:1].repeat(2)).numpy()
    w = torch.tensor(7.5).repeat(2)
    out["w_embedding"] = ref.LatentConsistencyModelPipeline_controlnet.get_w_embedding(None, w, embedding_dim=256).numpy()

    # ---------------------------------------------------------------- 2. the reference __call__ over the oracle modules
    unet, vae = build_unet(), build_taesd()
    ua = UNetAdapter(unet)
    import diffusers

    pipe = ref.LatentConsistencyModelPipeline_controlnet(
        vae=VaeAdapter(vae), text_encoder=None, tokenizer=None, controlnet=diffusers.ControlNetModel(), unet=ua,
        scheduler=None, safety_checker=None, feature_extractor=None)
    H = W = 64
    y, u, v = imageproc.synthetic_frame(H, W, seed=3)
    rgb = imageproc.yuv420_to_rgb(y, u, v)
    ctx = random_context(1, seed=11)
    # videopipeline.py:28-32,110-112,126
    cpu_state = torch.Generator(device="cpu").get_state()
    np.random.seed(42)
    torch.manual_seed(42).set_state(cpu_state)
    img = PIL.Image.fromarray(rgb)
    res = pipe(prompt=None, prompt_embeds=ctx, height=H, width=W, num_inference_steps=4, image=img, control_image=img,
               controlnet_conditioning_scale=1, generator=None, strength=0.5)
    out["pipe_rgb_in"] = rgb
    out["pipe_ctx_seed"] = np.array([11])
    out["pipe_rgb_out"] = np.array(res.images[0])
    out["pipe_timesteps"] = np.array([int(c[1][0]) for c in ua.calls])
    for i, (lat, ts, eps) in enumerate(ua.calls):
        out[f"pipe_latents_in_{i}"] = lat.numpy()
        out[f"pipe_eps_{i}"] = eps.numpy()
    # ---------------------------------------------------------------- 3. Sobel operator + __call__ with a real ControlNet
    # (SURVEY.md 8(f) next-row #1). canny_gpu.py is importable as is (torch / torchvision / PIL only).
    spec = importlib.util.spec_from_file_location("ref_canny_gpu", "/root/reference/diffusert/lcm/canny_gpu.py")
    canny = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(canny)
    sobel = canny.SobelOperator(device="cpu")
    y2, u2, v2 = imageproc.synthetic_frame(96, 128, seed=9)
    rgb2 = imageproc.yuv420_to_rgb(y2, u2, v2)
    out["sobel_rgb_in"] = rgb2
    out["sobel_out"] = np.array(sobel(PIL.Image.fromarray(rgb2), 0.11, 0.8))        # videopipeline.py:109
    out["sobel_out_64"] = np.array(sobel(img, 0.11, 0.8))

    from oracle.weights import build_controlnet
    cn = build_controlnet()

    class ControlNetAdapter(diffusers.ControlNetModel):
        """diffusers ControlNetModel call signature (lcm_controlnet.py:558-566) over the oracle's restated module."""

        def __init__(self, net):
            super().__init__()
            self.net = net
            self.calls = []

        def forward(self, sample, ts, encoder_hidden_states=None, controlnet_cond=None, conditioning_scale=1.0,
                    guess_mode=False, return_dict=False):
            self.calls.append((float(conditioning_scale), bool(guess_mode), controlnet_cond.clone()))
            return self.net(sample, ts, encoder_hidden_states, controlnet_cond, conditioning_scale, guess_mode)

    class UNetAdapterCN(UNetAdapter):
        def forward(self, sample, ts, timestep_cond=None, encoder_hidden_states=None, cross_attention_kwargs=None,
                    down_block_additional_residuals=None, mid_block_additional_residual=None, return_dict=False):
            o = self.net(sample, ts, timestep_cond, encoder_hidden_states, down_block_additional_residuals,
                         mid_block_additional_residual)
            self.calls.append((sample.clone(), ts.clone(), o.clone()))
            return (o,)

    ua2 = UNetAdapterCN(unet)
    cna = ControlNetAdapter(cn)
    pipe2 = ref.LatentConsistencyModelPipeline_controlnet(
        vae=VaeAdapter(vae), text_encoder=None, tokenizer=None, controlnet=cna, unet=ua2, scheduler=None,
        safety_checker=None, feature_extractor=None)
    canny_image = sobel(img, 0.11, 0.8)
    np.random.seed(42)
    torch.manual_seed(42).set_state(cpu_state)
    res2 = pipe2(prompt=None, prompt_embeds=ctx, height=H, width=W, num_inference_steps=4, image=img,
                 control_image=canny_image, controlnet_conditioning_scale=0.7, generator=None, strength=0.5)
    out["cn_scale"] = np.array([0.7])
    out["cn_rgb_out"] = np.array(res2.images[0])
    out["cn_control"] = cna.calls[0][2].numpy()
    out["cn_call_scales"] = np.array([c[0] for c in cna.calls])
    out["cn_guess_mode"] = np.array([c[1] for c in cna.calls])
    for i, (lat, ts, eps) in enumerate(ua2.calls):
        out[f"cn_latents_in_{i}"] = lat.numpy()
        out[f"cn_eps_{i}"] = eps.numpy()
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote reference_golden.npz:", {k: getattr(v, "shape", None) for k, v in out.items() if k.startswith("pipe")})


if __name__ == "__main__":
    main()
