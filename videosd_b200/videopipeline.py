"""Drop-in for the reference's frame-transform boundary, diffusert/videopipeline.py.

Same surface as the reference class (`VideoSDPipeline.remote(**config)`, `await pipe.infer.remote(pil_img, **options)
-> PIL.Image`, ctor kwargs model / controlnet / gpus / compile / device, missing model or controlnet -> KeyError
re-raised, videopipeline.py:16-32; infer signature :75-88), but everything between the cropped/resized RGB frame and
the output RGB frame runs in the CUDA engine (videosd_b200/engine.py -> libvideosd.so) instead of
diffusers + PyTorch. `infer_yuv420` additionally moves the YUV420<->RGB conversions that the reference leaves to
PyAV/libswscale in server.py:108,117 inside the boundary.

SURVEY.md 8(f) "next" rows inside the boundary: the canny ControlNet branch + Sobel (`use_controlnet=True`), the center
crop + Lanczos resize of arbitrary-size frames (GPU kernels, bit-identical to the reference's PIL calls) and the CLIP text
encoder (`text_encoder=True` / a checkpoint with text_encoder/ + tokenizer/): prompts are tokenized on the host
(videosd_b200/tokenizer.py) and encoded on the GPU once per prompt change. Without text-encoder weights a
`prompt_encoder=callable(list[str]) -> (77, 768)` can be plugged in; without either a deterministic pseudo-embedding
derived from the prompt text is used.
"""
import asyncio
import concurrent.futures
import hashlib
import itertools
import os

import numpy as np
import torch
from PIL import Image

from . import weights as _weights
from .engine import Engine

try:  # the reference uses Ray actors; Ray is optional here
    import ray  # noqa: F401

    _HAVE_RAY = True
except Exception:  # noqa: BLE001
    _HAVE_RAY = False


class _Awaitable:
    """Result of `.remote(...)`: awaitable from an asyncio loop (like a Ray ObjectRef) and `.result()`-able."""

    def __init__(self, fut):
        self._fut = fut

    def __await__(self):
        return asyncio.wrap_future(self._fut).__await__()

    def result(self, timeout=None):
        return self._fut.result(timeout)


class _RemoteMethod:
    def __init__(self, handle, name):
        self._handle, self._name = handle, name

    def remote(self, *args, **kwargs):
        h = self._handle
        return _Awaitable(h._pool.submit(lambda: getattr(h._obj, self._name)(*args, **kwargs)))


class _ActorHandle:
    """Minimal stand-in for a Ray actor handle: one worker thread per instance => methods run serially,
    exactly one in-flight infer per GPU (server.py:132-137)."""

    def __init__(self, cls, args, kwargs):
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="videosd-gpu")
        self._obj = self._pool.submit(lambda: cls(*args, **kwargs)).result()

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return _RemoteMethod(self, name)


_device_counter = itertools.count()


def _pseudo_prompt_embedding(prompt):
    text = prompt if isinstance(prompt, str) else "\n".join(prompt)
    seed = int.from_bytes(hashlib.sha256(text.encode()).digest()[:8], "little") % (2 ** 63)
    return torch.randn((77, 768), generator=torch.Generator().manual_seed(seed))


class VideoSDPipeline:
    def __init__(self, *args, **kwargs):
        try:
            self.model_name = kwargs["model"]
            self.controlnet_name = kwargs["controlnet"]
        except KeyError:
            print("Model name and controlnet model must be specified")
            raise
        if "device" in kwargs:
            self.device = int(kwargs["device"])
        else:
            n = max(torch.cuda.device_count(), 1)
            self.device = 0 if _HAVE_RAY and os.environ.get("RAY_ACTOR") else next(_device_counter) % n
        self.prompt_encoder = kwargs.get("prompt_encoder")
        # The reference runs the canny ControlNet before every UNet pass; the north star's hot path excludes it, so it is
        # opt-in here: use_controlnet=True adds ~35 % FLOPs per step.
        self.use_controlnet = bool(kwargs.get("use_controlnet", False))
        self.use_text_encoder = bool(kwargs.get("text_encoder", False))
        # vae="kl": the pipeline's declared AutoencoderKL instead of the AutoencoderTiny the reference loads
        # (videopipeline.py:67-69); SURVEY.md 8(f) next-row #4
        self.vae_kind = str(kwargs.get("vae", "taesd"))
        self.tokenizer = None
        self.noise_mode = kwargs.get("noise_mode", "reference_cuda")
        self.engine = Engine(self.device)          # raises if the CUDA library / a B200 is missing: no fallback
        self.load_model(self.model_name, self.controlnet_name, kwargs.get("random_init", False))
        self._cn_scale = None
        self._shape = None
        self._noise_key = None
        self._prompt_key = None

    @classmethod
    def remote(cls, *args, **kwargs):
        return _ActorHandle(cls, args, kwargs)

    def load_model(self, model_name, controlnet_model=None, random_init=False):
        """Loads UNet + TAESD weights. `model_name` may be a local diffusers directory holding unet/ and vae/
        safetensors; with random_init (or VIDEOSD_RANDOM_INIT=1) seeded random weights of the architecture are used
        (benchmarks / tests; no checkpoints can be downloaded in this environment)."""
        if os.path.isdir(str(model_name)) and os.path.isdir(os.path.join(model_name, "unet")):
            self.engine.load_state_dict("unet", _weights.load_safetensors_dir(os.path.join(model_name, "unet")))
            self.engine.load_state_dict("vae", _weights.load_safetensors_dir(os.path.join(model_name, "vae")))
            if self.use_controlnet:
                self.engine.load_state_dict("controlnet", _weights.load_safetensors_dir(str(controlnet_model)))
            if self.vae_kind == "kl":
                self.engine.load_state_dict("vae_kl", _weights.load_safetensors_dir(os.path.join(model_name, "vae_kl")))
            if os.path.isdir(os.path.join(model_name, "text_encoder")):
                from . import tokenizer as _tok
                self.engine.load_state_dict("text_encoder", _weights.load_safetensors_dir(os.path.join(model_name, "text_encoder")))
                self.tokenizer = _tok.load(os.path.join(model_name, "tokenizer"))
                self.use_text_encoder = True
        elif random_init or os.environ.get("VIDEOSD_RANDOM_INIT") == "1":
            self.engine.load_state_dict("unet", _weights.random_state_dict(_weights.unet_param_shapes(), 1234))
            self.engine.load_state_dict("vae", _weights.random_state_dict(_weights.taesd_param_shapes(), 4321))
            if self.use_controlnet:
                self.engine.load_state_dict("controlnet",
                                            _weights.random_state_dict(_weights.controlnet_param_shapes(), 9876))
            if self.vae_kind == "kl":
                self.engine.load_state_dict("vae_kl", _weights.random_state_dict(_weights.autoencoder_kl_param_shapes(), 2222))
            if self.use_text_encoder:
                from . import tokenizer as _tok
                self.engine.load_state_dict("text_encoder", _weights.random_clip_state_dict(2468))
                self.tokenizer = _tok.load(None)
        else:
            raise FileNotFoundError(
                f"model '{model_name}' is not a local diffusers directory (unet/, vae/ with .safetensors). "
                "There is no network here; pass random_init=True for seeded random weights.")
        return self.engine

    # ------------------------------------------------------------------------------------------------
    def _prepare(self, batch, height, width, strength, steps, guidance_scale, seed, prompt, prompt_embeds=None,
                 controlnet_scale=1.0):
        if self._shape != (batch, height, width):
            self.engine.configure(batch, height, width)
            if self.vae_kind == "kl":
                self.engine.set_vae("kl")
            self._shape = (batch, height, width)
            self._noise_key = self._prompt_key = None
        if self.use_controlnet and self._cn_scale != float(controlnet_scale):
            self.engine.set_controlnet(True, float(controlnet_scale))   # lcm_controlnet.py:553-556 (keep = 1.0)
            self._cn_scale = float(controlnet_scale)
        ts = self.engine.set_schedule(strength, steps, 7.5)   # guidance_scale from the UI is dropped by the reference (F8)
        nkey = (seed, len(ts), self.noise_mode)
        if nkey != self._noise_key:
            h8, w8 = height // 8, width // 8
            if self.noise_mode == "reference_cpu":
                self.engine.set_reference_noise()
            else:
                # reference on a CUDA device: torch.manual_seed(seed) seeds the device Philox that draws the init
                # noise in the model dtype (lcm_controlnet.py:331); step noise always comes from the re-armed CPU RNG
                g = torch.Generator(device=f"cuda:{self.device}").manual_seed(int(seed))
                init = torch.randn((batch, 4, h8, w8), generator=g, device=f"cuda:{self.device}", dtype=torch.float16)
                # On a CUDA device the init noise never touches the CPU RNG (randn_tensor(generator=None, device=cuda),
                # lcm_controlnet.py:503-513, :331), so step noise i is draw i of the re-armed CPU global generator
                # (videopipeline.py:126; scheduler.step :1033) -- no draw is skipped.
                gc = torch.Generator()
                st = [torch.randn((batch, 4, h8, w8), generator=gc) for _ in range(len(ts))] if len(ts) > 1 else []
                if self.vae_kind == "kl":
                    # latent_dist.sample() draws from the same device generator BEFORE the init noise (lcm_controlnet.py:298-331)
                    g = torch.Generator(device=f"cuda:{self.device}").manual_seed(int(seed))
                    vn = torch.randn((batch, 4, h8, w8), generator=g, device=f"cuda:{self.device}", dtype=torch.float16)
                    init = torch.randn((batch, 4, h8, w8), generator=g, device=f"cuda:{self.device}", dtype=torch.float16)
                    self.engine.set_vae_noise(vn.float().cpu())
                self.engine.set_noise(init.float().cpu(), st)
            self._noise_key = nkey
        pkey = prompt if isinstance(prompt, str) else tuple(prompt)
        if prompt_embeds is not None or pkey != self._prompt_key:
            if prompt_embeds is not None:
                emb = torch.as_tensor(prompt_embeds).reshape(-1, 77, 768)[0]
            elif self.prompt_encoder is not None:
                emb = torch.as_tensor(self.prompt_encoder(prompt)).reshape(-1, 77, 768)[0]
            elif self.use_text_encoder:
                # lcm_controlnet.py:144-179: tokenizer(prompt, padding="max_length", max_length=77) -> text_encoder(ids)[0];
                # a list of prompts is a batch there, and server.py sends one prompt per stream: the first entry conditions it
                text = prompt if isinstance(prompt, str) else prompt[0]
                emb = self.engine.encode_prompt(self.tokenizer(text))
            else:
                emb = _pseudo_prompt_embedding(prompt)
            for b in range(batch):
                self.engine.set_context(b, emb)
            self._prompt_key = None if prompt_embeds is not None else pkey
        return ts

    @staticmethod
    def _fit(img, width, height):
        """Center-crop to the target aspect ratio, then Lanczos-resize (videopipeline.py:92-107)."""
        if img.width / img.height > width / height:
            new_w = img.height * (width / height)
            box = ((img.width - new_w) / 2, 0, (img.width + new_w) / 2, img.height)
        else:
            new_h = img.width * (height / width)
            box = (0, (img.height - new_h) / 2, img.width, (img.height + new_h) / 2)
        img = img.crop(box)
        return img.resize((width, height), resample=Image.Resampling.LANCZOS)

    def infer(self, img, prompt=["pixar, cg"], height=360, width=640, strength=0.4, steps=20, guidance_scale=7.5,
              ref=False, style_fidelity=0.0, controlnet=False, seed=42, controlnet_scale=1, prompt_embeds=None,
              **_ignored):
        width, height = int(width), int(height)
        width -= width % 8
        height -= height % 8
        img = img.convert("RGB")
        self._prepare(1, height, width, float(strength), int(steps), guidance_scale, int(seed), prompt, prompt_embeds,
                      controlnet_scale)
        rgb_in = np.ascontiguousarray(np.asarray(img, dtype=np.uint8))
        rgb_out = np.empty((height, width, 3), dtype=np.uint8)
        if (img.width, img.height) == (width, height):
            self.engine.infer_rgb(rgb_in, rgb_out)
        else:
            # center crop + Lanczos resize on the GPU, bit-identical to the reference's PIL calls (videopipeline.py:92-107)
            key = (img.width, img.height, width, height, 1)
            if getattr(self.engine, "_resize_key", None) != key:
                self.engine.set_resize(img.width, img.height)
            self.engine.infer_rgb_resized(rgb_in, rgb_out)
        return Image.fromarray(rgb_out)

    def infer_yuv420(self, y, u, v, prompt=["pixar, cg"], strength=0.4, steps=20, seed=42, prompt_embeds=None,
                     controlnet_scale=1, height=None, width=None, **_ignored):
        """Fast path: YUV420P planes (u8 numpy, (B,H,W) or (H,W)) -> planes at the working size. With height/width
        different from the planes' size the frame is colour-converted, center-cropped and Lanczos-resized on the GPU
        (same result as frame.to_image() + the reference's PIL crop/resize, videopipeline.py:92-107)."""
        y, u, v = (np.ascontiguousarray(a) for a in (y, u, v))
        if y.ndim == 2:
            y, u, v = y[None], u[None], v[None]
        b, h, w = y.shape
        oh, ow = (h if height is None else int(height)), (w if width is None else int(width))
        self._prepare(b, oh, ow, float(strength), int(steps), 7.5, int(seed), prompt, prompt_embeds, controlnet_scale)
        oy = np.empty((b, oh, ow), np.uint8)
        ou, ov = np.empty((b, oh // 2, ow // 2), np.uint8), np.empty((b, oh // 2, ow // 2), np.uint8)
        if (oh, ow) == (h, w):
            self.engine.infer_yuv420(y, u, v, oy, ou, ov)
        else:
            key = (w, h, ow, oh, b)
            if getattr(self.engine, "_resize_key", None) != key:
                self.engine.set_resize(w, h)
            self.engine.infer_yuv420_resized(y, u, v, oy, ou, ov)
        return oy, ou, ov
