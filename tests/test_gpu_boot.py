"""-m gpu: the replacement of `from_pretrained` (reference diffusert/videopipeline.py:51-69): a diffusers-layout checkpoint
directory on disk (unet/, vae/ with .safetensors) boots VideoSDPipeline and produces the same frames as the same tensors
handed over in memory; a checkpoint without text_encoder/ + tokenizer/ refuses prompts instead of inventing a context."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _save(sd, path):
    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    out = {}
    for k, v in sd.items():     # matrices as bf16 (what the engine computes in), vectors and the time-embedding path as fp32
        low = v.ndim == 4 or (v.ndim == 2 and "time_emb" not in k)
        out[k] = (v.to(torch.bfloat16) if low else v.float()).contiguous()
    save_file(out, os.path.join(path, "diffusion_pytorch_model.safetensors"))
    return {k: v.float() for k, v in out.items()}


def test_boot_from_a_diffusers_safetensors_directory(tmp_path, oracle_models):
    from PIL import Image

    from oracle.weights import random_context
    from videosd_b200.videopipeline import VideoSDPipeline

    unet, vae = oracle_models
    root = tmp_path / "LCM_Dreamshaper_v7"
    sd_u = _save(unet.state_dict(), str(root / "unet"))
    sd_v = _save(vae.state_dict(), str(root / "vae"))
    pipe = VideoSDPipeline(model=str(root), controlnet="lllyasviel/control_v11p_sd15_canny", gpus=1, compile=False, device=0)
    assert pipe.random_init is False and pipe.use_text_encoder is False
    img = Image.fromarray((np.random.RandomState(5).rand(512, 512, 3) * 255).astype(np.uint8))
    kw = dict(width=512, height=512, strength=0.5, steps=4, seed=42)   # a size with a committed tuning table: same kernels in both
    with pytest.raises(RuntimeError, match="never replaced"):      # no text tower in this checkpoint: no stand-in context
        pipe.infer(img, prompt="pixar, cg", **kw)
    ctx = random_context(1, seed=2)
    a = np.asarray(pipe.infer(img, prompt_embeds=ctx, **kw))
    mem = VideoSDPipeline(model="in-memory", controlnet="c", device=0, state_dicts={"unet": sd_u, "vae": sd_v})
    b = np.asarray(mem.infer(img, prompt_embeds=ctx, **kw))
    assert a.std() > 1.0 and np.array_equal(a, b)
    with pytest.raises(FileNotFoundError):                          # neither a directory nor random_init
        VideoSDPipeline(model=str(tmp_path / "missing"), controlnet="c", device=0)
