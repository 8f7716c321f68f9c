"""Oracle restatement of diffusers' UNet2DConditionModel as configured by SimianLuo/LCM_Dreamshaper_v7
(TEST INFRASTRUCTURE ONLY). Spec: SURVEY.md Appendix A.1-A.4 [diffusers-knowledge]; the reference calls it at
diffusert/lcm/lcm_controlnet.py:568-577 as unet(latents, ts, timestep_cond=w_emb, encoder_hidden_states=ctx).

Module / parameter names reproduce the diffusers state-dict keys (Appendix A.7) so real checkpoints would load.
Plain torch.nn, fp32, NCHW.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def timestep_sinusoid(t, dim=320):
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): returns [cos, sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t[:, None].float() * freqs[None, :]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim=320, dim=1280, cond_dim=256):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)
        self.cond_proj = nn.Linear(cond_dim, in_dim, bias=False)

    def forward(self, sample, condition):
        sample = sample + self.cond_proj(condition)
        return self.linear_2(F.silu(self.linear_1(sample)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb=1280, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, emb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(emb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim, ctx_dim=None, heads=8):
        super().__init__()
        self.heads = heads
        kv_dim = ctx_dim if ctx_dim is not None else dim
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(kv_dim, dim, bias=False)
        self.to_v = nn.Linear(kv_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])  # [Linear, Dropout(0)]

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b, n, c = x.shape
        d = c // self.heads
        q = self.to_q(x).view(b, n, self.heads, d).transpose(1, 2)
        k = self.to_k(ctx).view(b, -1, self.heads, d).transpose(1, 2)
        v = self.to_v(ctx).view(b, -1, self.heads, d).transpose(1, 2)
        s = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
        o = torch.matmul(torch.softmax(s, dim=-1), v)
        o = o.transpose(1, 2).reshape(b, n, c)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        u, g = self.proj(x).chunk(2, dim=-1)  # first half value, second half gate
        return u * F.gelu(g)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, ctx_dim=768, heads=8):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, ctx_dim, heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, dim, ctx_dim=768, heads=8, groups=32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, ctx_dim, heads)])
        self.proj_out = nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        res = x
        x = self.proj_in(self.norm(x))
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
        for blk in self.transformer_blocks:
            x = blk(x, ctx)
        x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
        return self.proj_out(x) + res


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x, size=None):
        if size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=size, mode="nearest")
        return self.conv(x)


class DownBlock(nn.Module):
    def __init__(self, cin, cout, attn, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin, cout), ResnetBlock2D(cout, cout)])
        if attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout), Transformer2DModel(cout)])
        else:
            self.attentions = None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, emb, ctx):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, emb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c), ResnetBlock2D(c, c)])
        self.attentions = nn.ModuleList([Transformer2DModel(c)])

    def forward(self, x, emb, ctx):
        x = self.resnets[0](x, emb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, emb)


class UpBlock(nn.Module):
    def __init__(self, prev_c, skip_cs, cout, attn, add_up):
        super().__init__()
        rs = []
        for j, sc in enumerate(skip_cs):
            cin = (prev_c if j == 0 else cout) + sc
            rs.append(ResnetBlock2D(cin, cout))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList([Transformer2DModel(cout) for _ in skip_cs]) if attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, emb, ctx, upsample_size=None):
        for j, r in enumerate(self.resnets):
            s = skips.pop()
            x = r(torch.cat([x, s], dim=1), emb)
            if self.attentions is not None:
                x = self.attentions[j](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x, upsample_size)
        return x


class UNetLCM(nn.Module):
    """SD1.5 topology, block widths (320,640,1280,1280), time_cond_proj_dim=256 (Appendix A.1)."""

    widths = (320, 640, 1280, 1280)

    def __init__(self):
        super().__init__()
        w = self.widths
        self.conv_in = nn.Conv2d(4, w[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(w[0], w[0] * 4, 256)
        downs = []
        cin = w[0]
        for i, c in enumerate(w):
            downs.append(DownBlock(cin, c, attn=(i < 3), add_down=(i < 3)))
            cin = c
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(w[-1])
        # skip widths in push order: conv_in, then per block (res, res, down)
        skip = [w[0]]
        for i, c in enumerate(w):
            skip += [c, c] + ([c] if i < 3 else [])
        ups = []
        prev = w[-1]
        rev = list(reversed(w))
        for i, c in enumerate(rev):
            sk = [skip.pop() for _ in range(3)]
            ups.append(UpBlock(prev, sk, c, attn=(i > 0), add_up=(i < 3)))
            prev = c
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(32, w[0], eps=1e-5)
        self.conv_out = nn.Conv2d(w[0], 4, 3, padding=1)

    def time_embed(self, timesteps, timestep_cond):
        return self.time_embedding(timestep_sinusoid(timesteps, self.widths[0]), timestep_cond)

    def forward(self, sample, timesteps, timestep_cond, encoder_hidden_states, down_residuals=None, mid_residual=None):
        """down_residuals / mid_residual: ControlNet outputs added to the 12 skips and to the mid-block output
        (diffusers UNet2DConditionModel.forward: down_block_additional_residuals / mid_block_additional_residual)."""
        emb = self.time_embed(timesteps, timestep_cond)
        # nearest-2x only reproduces the skip sizes when H, W are multiples of 2^3 (forward_upsample_size)
        need_sizes = any(s % 8 != 0 for s in sample.shape[-2:])
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips += outs
        x = self.mid_block(x, emb, encoder_hidden_states)
        if down_residuals is not None:
            skips = [s_ + r_ for s_, r_ in zip(skips, down_residuals)]
        if mid_residual is not None:
            x = x + mid_residual
        for i, blk in enumerate(self.up_blocks):
            n = len(blk.resnets)
            size = None
            if need_sizes and blk.upsamplers is not None:
                size = skips[-n - 1].shape[2:]
            x = blk(x, skips, emb, encoder_hidden_states, size)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return x
