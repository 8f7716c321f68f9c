"""Parameter tables (diffusers state-dict names and shapes) of the two networks on the hot path, a random-init
generator for benchmarks (no checkpoints can be downloaded here), and a safetensors loader for real checkpoints.

Names follow SURVEY.md Appendix A.7: UNet2DConditionModel of SimianLuo/LCM_Dreamshaper_v7 (SD1.5 topology +
time_embedding.cond_proj) and AutoencoderTiny (madebyollin/taesd), which the reference loads at
diffusert/videopipeline.py:49-72.
"""
import math
import os

import torch


def _resnet(p, cin, cout, out):
    out[p + ".norm1.weight"] = (cin,); out[p + ".norm1.bias"] = (cin,)
    out[p + ".conv1.weight"] = (cout, cin, 3, 3); out[p + ".conv1.bias"] = (cout,)
    out[p + ".time_emb_proj.weight"] = (cout, 1280); out[p + ".time_emb_proj.bias"] = (cout,)
    out[p + ".norm2.weight"] = (cout,); out[p + ".norm2.bias"] = (cout,)
    out[p + ".conv2.weight"] = (cout, cout, 3, 3); out[p + ".conv2.bias"] = (cout,)
    if cin != cout:
        out[p + ".conv_shortcut.weight"] = (cout, cin, 1, 1); out[p + ".conv_shortcut.bias"] = (cout,)


def _transformer(p, c, out, ctx=768):
    out[p + ".norm.weight"] = (c,); out[p + ".norm.bias"] = (c,)
    out[p + ".proj_in.weight"] = (c, c, 1, 1); out[p + ".proj_in.bias"] = (c,)
    t = p + ".transformer_blocks.0"
    for n in ("norm1", "norm2", "norm3"):
        out[f"{t}.{n}.weight"] = (c,); out[f"{t}.{n}.bias"] = (c,)
    for a, kv in (("attn1", c), ("attn2", ctx)):
        out[f"{t}.{a}.to_q.weight"] = (c, c)
        out[f"{t}.{a}.to_k.weight"] = (c, kv)
        out[f"{t}.{a}.to_v.weight"] = (c, kv)
        out[f"{t}.{a}.to_out.0.weight"] = (c, c); out[f"{t}.{a}.to_out.0.bias"] = (c,)
    out[f"{t}.ff.net.0.proj.weight"] = (8 * c, c); out[f"{t}.ff.net.0.proj.bias"] = (8 * c,)
    out[f"{t}.ff.net.2.weight"] = (c, 4 * c); out[f"{t}.ff.net.2.bias"] = (c,)
    out[p + ".proj_out.weight"] = (c, c, 1, 1); out[p + ".proj_out.bias"] = (c,)


def unet_param_shapes():
    w = (320, 640, 1280, 1280)
    o = {}
    o["conv_in.weight"] = (320, 4, 3, 3); o["conv_in.bias"] = (320,)
    o["time_embedding.linear_1.weight"] = (1280, 320); o["time_embedding.linear_1.bias"] = (1280,)
    o["time_embedding.linear_2.weight"] = (1280, 1280); o["time_embedding.linear_2.bias"] = (1280,)
    o["time_embedding.cond_proj.weight"] = (320, 256)
    cin = 320
    skips = [320]
    for i, c in enumerate(w):
        for j in range(2):
            _resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else c, c, o)
            if i < 3:
                _transformer(f"down_blocks.{i}.attentions.{j}", c, o)
            skips.append(c)
        if i < 3:
            o[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (c, c, 3, 3)
            o[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (c,)
            skips.append(c)
        cin = c
    _resnet("mid_block.resnets.0", 1280, 1280, o)
    _transformer("mid_block.attentions.0", 1280, o)
    _resnet("mid_block.resnets.1", 1280, 1280, o)
    prev = 1280
    for i, c in enumerate(reversed(w)):
        for j in range(3):
            s = skips.pop()
            _resnet(f"up_blocks.{i}.resnets.{j}", (prev if j == 0 else c) + s, c, o)
            if i > 0:
                _transformer(f"up_blocks.{i}.attentions.{j}", c, o)
        if i < 3:
            o[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (c, c, 3, 3)
            o[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (c,)
        prev = c
    o["conv_norm_out.weight"] = (320,); o["conv_norm_out.bias"] = (320,)
    o["conv_out.weight"] = (4, 320, 3, 3); o["conv_out.bias"] = (4,)
    return o


def controlnet_param_shapes():
    """diffusers ControlNetModel (lllyasviel/control_v11p_sd15_canny), SURVEY.md Appendix A.8."""
    w = (320, 640, 1280, 1280)
    o = {}
    o["conv_in.weight"] = (320, 4, 3, 3); o["conv_in.bias"] = (320,)
    o["time_embedding.linear_1.weight"] = (1280, 320); o["time_embedding.linear_1.bias"] = (1280,)
    o["time_embedding.linear_2.weight"] = (1280, 1280); o["time_embedding.linear_2.bias"] = (1280,)
    e = "controlnet_cond_embedding"
    o[f"{e}.conv_in.weight"] = (16, 3, 3, 3); o[f"{e}.conv_in.bias"] = (16,)
    ws = (16, 32, 96, 256)
    for i in range(3):
        o[f"{e}.blocks.{2*i}.weight"] = (ws[i], ws[i], 3, 3); o[f"{e}.blocks.{2*i}.bias"] = (ws[i],)
        o[f"{e}.blocks.{2*i+1}.weight"] = (ws[i + 1], ws[i], 3, 3); o[f"{e}.blocks.{2*i+1}.bias"] = (ws[i + 1],)
    o[f"{e}.conv_out.weight"] = (320, 256, 3, 3); o[f"{e}.conv_out.bias"] = (320,)
    cin = 320
    chans = [320]
    for i, c in enumerate(w):
        for j in range(2):
            _resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else c, c, o)
            if i < 3:
                _transformer(f"down_blocks.{i}.attentions.{j}", c, o)
            chans.append(c)
        if i < 3:
            o[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (c, c, 3, 3)
            o[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (c,)
            chans.append(c)
        cin = c
    _resnet("mid_block.resnets.0", 1280, 1280, o)
    _transformer("mid_block.attentions.0", 1280, o)
    _resnet("mid_block.resnets.1", 1280, 1280, o)
    for k, c in enumerate(chans):
        o[f"controlnet_down_blocks.{k}.weight"] = (c, c, 1, 1); o[f"controlnet_down_blocks.{k}.bias"] = (c,)
    o["controlnet_mid_block.weight"] = (1280, 1280, 1, 1); o["controlnet_mid_block.bias"] = (1280,)
    return o


def taesd_param_shapes():
    o = {}

    def block(p):
        for k in (0, 2, 4):
            o[f"{p}.conv.{k}.weight"] = (64, 64, 3, 3); o[f"{p}.conv.{k}.bias"] = (64,)

    e = "encoder.layers"
    o[f"{e}.0.weight"] = (64, 3, 3, 3); o[f"{e}.0.bias"] = (64,)
    block(f"{e}.1")
    n = 2
    for _ in range(3):
        o[f"{e}.{n}.weight"] = (64, 64, 3, 3); n += 1
        for _ in range(3):
            block(f"{e}.{n}"); n += 1
    o[f"{e}.{n}.weight"] = (4, 64, 3, 3); o[f"{e}.{n}.bias"] = (4,)
    d = "decoder.layers"
    o[f"{d}.0.weight"] = (64, 4, 3, 3); o[f"{d}.0.bias"] = (64,)
    n = 2
    for _ in range(3):
        for _ in range(3):
            block(f"{d}.{n}"); n += 1
        n += 1  # nn.Upsample
        o[f"{d}.{n}.weight"] = (64, 64, 3, 3); n += 1
    block(f"{d}.{n}"); n += 1
    o[f"{d}.{n}.weight"] = (3, 64, 3, 3); o[f"{d}.{n}.bias"] = (3,)
    return o


def autoencoder_kl_param_shapes():
    """diffusers AutoencoderKL (SD1.5 VAE) state-dict keys: 83 653 863 parameters."""
    out = {}

    def conv(p, cin, cout, k=3):
        out[p + ".weight"] = (cout, cin, k, k)
        out[p + ".bias"] = (cout,)

    def norm(p, c):
        out[p + ".weight"] = (c,)
        out[p + ".bias"] = (c,)

    def resnet(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cin, cout)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout)
        if cin != cout:
            conv(p + ".conv_shortcut", cin, cout, 1)

    def mid(p, c):
        a = p + ".attentions.0"
        norm(a + ".group_norm", c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            out[f"{a}.{n}.weight"] = (c, c)
            out[f"{a}.{n}.bias"] = (c,)
        resnet(p + ".resnets.0", c, c)
        resnet(p + ".resnets.1", c, c)

    w = (128, 256, 512, 512)
    conv("encoder.conv_in", 3, w[0])
    for i in range(4):
        cin = w[max(i - 1, 0)]
        resnet(f"encoder.down_blocks.{i}.resnets.0", cin, w[i])
        resnet(f"encoder.down_blocks.{i}.resnets.1", w[i], w[i])
        if i < 3:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", w[i], w[i])
    mid("encoder.mid_block", 512)
    norm("encoder.conv_norm_out", 512)
    conv("encoder.conv_out", 512, 8)
    conv("decoder.conv_in", 4, 512)
    mid("decoder.mid_block", 512)
    rev = (512, 512, 256, 128)
    for i in range(4):
        cin = rev[max(i - 1, 0)]
        resnet(f"decoder.up_blocks.{i}.resnets.0", cin, rev[i])
        resnet(f"decoder.up_blocks.{i}.resnets.1", rev[i], rev[i])
        resnet(f"decoder.up_blocks.{i}.resnets.2", rev[i], rev[i])
        if i < 3:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", rev[i], rev[i])
    norm("decoder.conv_norm_out", 128)
    conv("decoder.conv_out", 128, 3)
    conv("quant_conv", 8, 8, 1)
    conv("post_quant_conv", 4, 4, 1)
    return out


def clip_param_shapes():
    """transformers CLIPTextModel state-dict keys (SD1.5 text tower): 123 060 480 parameters."""
    out = {"text_model.embeddings.token_embedding.weight": (49408, 768),
           "text_model.embeddings.position_embedding.weight": (77, 768)}
    for i in range(12):
        p = f"text_model.encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            out[p + f"self_attn.{n}.weight"] = (768, 768)
            out[p + f"self_attn.{n}.bias"] = (768,)
        for n in ("layer_norm1", "layer_norm2"):
            out[p + n + ".weight"] = (768,)
            out[p + n + ".bias"] = (768,)
        out[p + "mlp.fc1.weight"] = (3072, 768)
        out[p + "mlp.fc1.bias"] = (3072,)
        out[p + "mlp.fc2.weight"] = (768, 3072)
        out[p + "mlp.fc2.bias"] = (768,)
    out["text_model.final_layer_norm.weight"] = (768,)
    out["text_model.final_layer_norm.bias"] = (768,)
    return out


def random_clip_state_dict(seed):
    """Seeded random text tower: embeddings N(0, 0.02) (CLIP's own init), linears U(+-1/sqrt(fan_in)), LayerNorm affine
    (1 + 0.1 N, 0.1 N) so that gamma / beta are exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in clip_param_shapes().items():
        if "embedding" in name:
            sd[name] = torch.randn(shape, generator=g) * 0.02
        elif "layer_norm" in name:
            sd[name] = (1.0 if name.endswith("weight") else 0.0) + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 2:
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[1])
        else:
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return sd


def random_state_dict(shapes, seed):
    """torch.nn default-style init (U(+-1/sqrt(fan_in)); norm affine = (1, 0)) from a seeded generator."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes.items():
        if ".norm" in name or name.startswith("conv_norm_out"):
            sd[name] = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
            continue
        if len(shape) == 1:
            fan_in = 64
        else:
            fan_in = int(math.prod(shape[1:]))
        bound = 1.0 / math.sqrt(fan_in)
        sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def load_safetensors_dir(path):
    """Loads a diffusers component directory (unet/ or vae/) containing *.safetensors into a dict of fp32 tensors."""
    from safetensors.torch import load_file

    sd = {}
    files = [f for f in sorted(os.listdir(path)) if f.endswith(".safetensors")]
    if not files:
        raise FileNotFoundError(f"no .safetensors file under {path}")
    for f in files:
        for k, v in load_file(os.path.join(path, f)).items():
            sd[k] = v.float()
    return sd
