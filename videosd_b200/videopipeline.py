"""Drop-in for the reference's frame-transform boundary, diffusert/videopipeline.py.

Same surface as the reference class (`VideoSDPipeline.remote(**config)`, `await pipe.infer.remote(pil_img, **options)
-> PIL.Image`, ctor kwargs model / controlnet / gpus / compile / device, missing model or controlnet -> KeyError
re-raised, videopipeline.py:16-32; infer signature :75-88), but everything between the cropped/resized RGB frame and
the output RGB frame runs in the CUDA engine (videosd_b200/engine.py -> libvideosd.so) instead of
diffusers + PyTorch. `infer_yuv420` additionally moves the YUV420<->RGB conversions that the reference leaves to
PyAV/libswscale in server.py:108,117 inside the boundary.

Beyond the reference's one-frame-per-actor model (server.py:132-137), two optional ctor kwargs (also accepted as
config.yaml keys, since server.py splats the whole file into the ctor, :321):
  frames_in_flight=N  N lanes on one copy of the weights: up to N `infer` calls of one actor run concurrently (the handle
                      returned by `.remote()` runs N worker threads; under Ray the actor is created with max_concurrency=N)
  max_batch=B         frames of different sessions that wait while all lanes are busy are merged into one batched launch
Scheduling lives in videosd_b200/dispatcher.py. Actors created in ONE process on the same GPU (Ray absent, `gpus: 4` on a
smaller box) share one copy of the weights.

SURVEY.md 8(f) "next" rows inside the boundary: the canny ControlNet branch + Sobel (`use_controlnet=True`), the center
crop + Lanczos resize of arbitrary-size frames (GPU kernels, bit-identical to the reference's PIL calls) and the CLIP text
encoder (a checkpoint with text_encoder/ + tokenizer/, or `text_encoder=True` with random weights): prompts are tokenized on
the host (videosd_b200/tokenizer.py) and encoded on the GPU once per prompt change. With a real checkpoint a prompt can
only be conditioned through the text encoder, a `prompt_encoder=callable(prompt) -> (77, 768)` or explicit `prompt_embeds`;
anything else raises (no silent stand-in). Only `random_init=True` pipelines (benchmarks / tests: the weights carry no
meaning) fall back to a deterministic pseudo-embedding of the prompt text.
"""
import asyncio
import concurrent.futures
import hashlib
import itertools
import os
import threading

import numpy as np
import torch
from PIL import Image

from . import weights as _weights
from .dispatcher import FrameDispatcher, FrameRequest, embedding_key
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
    """Minimal stand-in for a Ray actor handle when Ray is absent: `frames_in_flight` worker threads per instance (default
    1 => methods run serially, exactly one in-flight infer per actor as in server.py:132-137)."""

    def __init__(self, cls, args, kwargs):
        n = max(1, int(kwargs.get("frames_in_flight", 1))) * max(1, int(kwargs.get("max_batch", 1)))
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=n, thread_name_prefix="videosd-gpu")
        self._obj = self._pool.submit(lambda: cls(*args, **kwargs)).result()

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return _RemoteMethod(self, name)


_device_counter = itertools.count()
_shared_lock = threading.Lock()
_shared_roots = {}     # (device, model, controlnet, options) -> (weight-owning Engine, tokenizer): actors of one process share weights


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
            # one Ray actor sees its GPU as device 0 (num_gpus=1 sets CUDA_VISIBLE_DEVICES); in-process actors go round-robin
            n = max(torch.cuda.device_count(), 1)
            self.device = 0 if os.environ.get("VIDEOSD_RAY_ACTOR") == "1" else next(_device_counter) % n
        self.prompt_encoder = kwargs.get("prompt_encoder")
        # The reference runs the canny ControlNet before every UNet pass; the north star's hot path excludes it, so it is
        # opt-in here: use_controlnet=True adds ~35 % FLOPs per step.
        self.use_controlnet = bool(kwargs.get("use_controlnet", False))
        self.use_text_encoder = bool(kwargs.get("text_encoder", False))
        # vae="kl": the pipeline's declared AutoencoderKL instead of the AutoencoderTiny the reference loads
        # (videopipeline.py:67-69); SURVEY.md 8(f) next-row #4
        self.vae_kind = str(kwargs.get("vae", "taesd"))
        self.random_init = bool(kwargs.get("random_init", False)) or os.environ.get("VIDEOSD_RANDOM_INIT") == "1"
        self.tokenizer = None
        self.noise_mode = kwargs.get("noise_mode", "reference_cuda")
        self.frames_in_flight = max(1, int(kwargs.get("frames_in_flight", 1)))
        self.max_batch = max(1, int(kwargs.get("max_batch", 1)))
        key = (self.device, str(self.model_name), str(self.controlnet_name), self.use_controlnet, self.use_text_encoder,
               self.vae_kind, self.random_init)
        self._state_dicts = kwargs.get("state_dicts")   # tests: {"unet": sd, "vae": sd, ...} in diffusers layout instead of files
        with _shared_lock:
            shared = _shared_roots.get(key) if self._state_dicts is None else None
            if shared is None:
                self.engine = Engine(self.device)      # raises if the CUDA library / a B200 is missing: no fallback
                self.load_model(self.model_name, self.controlnet_name, self.random_init)
                if self._state_dicts is None:
                    _shared_roots[key] = (self.engine, self.tokenizer, self.use_text_encoder)
            else:                                      # another actor of this process already holds these weights on this GPU
                root, self.tokenizer, self.use_text_encoder = shared
                self.engine = Engine(self.device, parent=root)
        self.dispatcher = FrameDispatcher(self.engine, self.frames_in_flight, self.max_batch, self.noise_mode,
                                          self.use_controlnet, self.vae_kind)
        self._emb_cache = {}
        self._emb_lock = threading.Lock()

    @classmethod
    def remote(cls, *args, **kwargs):
        """`VideoSDPipeline.remote(**config)` (server.py:321). With Ray: a `ray.remote(num_gpus=1, num_cpus=4)` actor like the
        reference's (videopipeline.py:11), threaded when frames_in_flight > 1. Without Ray: an in-process handle with the
        same `.infer.remote(...)` -> awaitable surface."""
        if _HAVE_RAY and os.environ.get("VIDEOSD_NO_RAY") != "1":
            import ray as _ray
            n = max(1, int(kwargs.get("frames_in_flight", 1))) * max(1, int(kwargs.get("max_batch", 1)))
            actor = _ray.remote(num_gpus=1, num_cpus=4, max_concurrency=n)(_RayActor)
            return actor.remote(*args, **kwargs)
        return _ActorHandle(cls, args, kwargs)

    def load_model(self, model_name, controlnet_model=None, random_init=False):
        """Loads UNet + TAESD weights (replaces from_pretrained, videopipeline.py:49-72). `model_name`: a local diffusers
        directory holding unet/ and vae/ (and optionally text_encoder/ + tokenizer/, vae_kl/) with .safetensors files;
        `controlnet_model`: a directory with the ControlNet's .safetensors (read when use_controlnet). With random_init (or
        VIDEOSD_RANDOM_INIT=1) seeded random weights of the architecture are used (benchmarks / tests; no checkpoints can
        be downloaded in this environment)."""
        if self._state_dicts is not None:
            for prefix, sd in self._state_dicts.items():
                self.engine.load_state_dict(prefix, sd)
        elif os.path.isdir(str(model_name)) and os.path.isdir(os.path.join(model_name, "unet")):
            self.engine.load_state_dict("unet", _weights.load_safetensors_dir(os.path.join(model_name, "unet")))
            self.engine.load_state_dict("vae", _weights.load_safetensors_dir(os.path.join(model_name, "vae")))
            if self.use_controlnet:
                self.engine.load_state_dict("controlnet", _weights.load_safetensors_dir(str(controlnet_model)))
            if self.vae_kind == "kl":
                self.engine.load_state_dict("vae_kl", _weights.load_safetensors_dir(os.path.join(model_name, "vae_kl")))
            if os.path.isdir(os.path.join(model_name, "text_encoder")):
                from . import tokenizer as _tok
                self.tokenizer = _tok.load(os.path.join(model_name, "tokenizer"))   # raises without vocab.json / merges.txt
                self.engine.load_state_dict("text_encoder", _weights.load_safetensors_dir(os.path.join(model_name, "text_encoder")))
                self.use_text_encoder = True
            self.random_init = False
        elif random_init:
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
                self.tokenizer = _tok.load(None, allow_hash=True)    # random text tower: a vocabulary-free tokenizer is enough
        else:
            raise FileNotFoundError(
                f"model '{model_name}' is not a local diffusers directory (unet/, vae/ with .safetensors). "
                "There is no network here; pass random_init=True for seeded random weights.")
        return self.engine

    # ------------------------------------------------------------------------------------------------
    def _embedding(self, prompt, prompt_embeds):
        """-> (cache key, (77, 768) embedding) of the request's conditioning (lcm_controlnet.py:143-179)."""
        if prompt_embeds is not None:
            emb = torch.as_tensor(prompt_embeds).reshape(-1, 77, 768)[0]
            return embedding_key(emb), emb
        pkey = ("prompt", prompt if isinstance(prompt, str) else tuple(prompt))
        with self._emb_lock:
            emb = self._emb_cache.get(pkey)
            if emb is None:
                if self.prompt_encoder is not None:
                    emb = torch.as_tensor(self.prompt_encoder(prompt)).reshape(-1, 77, 768)[0]
                elif self.use_text_encoder:
                    # tokenizer(prompt, padding="max_length", max_length=77) -> text_encoder(ids)[0]; a list of prompts is a
                    # batch there, and server.py sends one prompt per stream: the first entry conditions it
                    text = prompt if isinstance(prompt, str) else prompt[0]
                    emb = self.engine.encode_prompt(self.tokenizer(text))
                elif self.random_init:
                    emb = _pseudo_prompt_embedding(prompt)
                else:
                    raise RuntimeError(
                        "this checkpoint has no text_encoder/ + tokenizer/: pass prompt_embeds=, or construct the pipeline "
                        "with prompt_encoder=callable(prompt) -> (77, 768). A prompt is never replaced by a stand-in embedding.")
                if len(self._emb_cache) >= 64:
                    self._emb_cache.pop(next(iter(self._emb_cache)))
                self._emb_cache[pkey] = emb
        return pkey, emb

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
        pkey, emb = self._embedding(prompt, prompt_embeds)
        # a frame that is not already the working size is center-cropped + Lanczos-resized on the GPU, bit-identical to the
        # reference's PIL calls (videopipeline.py:92-107)
        rgb_in = np.ascontiguousarray(np.asarray(img, dtype=np.uint8))[None]
        req = FrameRequest("rgb", rgb_in, 1, img.height, img.width, height, width, float(strength), int(steps), int(seed), pkey,
                           emb, float(controlnet_scale))
        return Image.fromarray(self.dispatcher.run(req)[0])

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
        pkey, emb = self._embedding(prompt, prompt_embeds)
        req = FrameRequest("yuv", (y, u, v), b, h, w, oh, ow, float(strength), int(steps), int(seed), pkey, emb,
                           float(controlnet_scale))
        return self.dispatcher.run(req)


class _RayActor(VideoSDPipeline):
    """The class handed to ray.remote: inside a Ray actor the GPU assigned by num_gpus=1 is device 0 (videopipeline.py:11,20)."""

    def __init__(self, *args, **kwargs):
        os.environ["VIDEOSD_RAY_ACTOR"] = "1"
        super().__init__(*args, **kwargs)
