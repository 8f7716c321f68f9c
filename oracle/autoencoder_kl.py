"""Oracle restatement of diffusers' AutoencoderKL as shipped with Stable Diffusion 1.5 (TEST INFRASTRUCTURE ONLY).
SURVEY.md 8(f) row 4: the pipeline's declared VAE type (lcm_controlnet.py:69) and what the checkpoints ship; the reference
calls it as `self.vae.encode(image).latent_dist.sample(generator) * scaling_factor` (lcm_controlnet.py:55-58, :298-313) and
`self.vae.decode(denoised / scaling_factor)` (:594-596). Parity unpinned (diffusers is not installable here): restated from
knowledge of diffusers ~0.23 [diffusers-knowledge], checked structurally (83 653 863 parameters, state-dict keys).

config: in/out 3, latent 4, block_out_channels (128, 256, 512, 512), layers_per_block 2, norm_num_groups 32 (eps 1e-6),
SiLU, mid-block self-attention (1 head of 512), scaling_factor 0.18215. Plain torch.nn, fp32, NCHW.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

SCALING = 0.18215
WIDTHS = (128, 256, 512, 512)


class Resnet(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (self.conv_shortcut(x) if self.conv_shortcut is not None else x) + h


class Attention(nn.Module):
    """diffusers Attention(heads=1, dim_head=512, norm_num_groups=32, residual_connection=True, bias=True)."""

    def __init__(self, c):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, c, eps=1e-6)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x.view(b, c, h * w)).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        p = torch.softmax(q @ k.transpose(1, 2) * c ** -0.5, dim=-1)
        o = self.to_out[0](p @ v)
        return o.transpose(1, 2).reshape(b, c, h, w) + x


class Mid(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attentions = nn.ModuleList([Attention(c)])
        self.resnets = nn.ModuleList([Resnet(c, c), Resnet(c, c)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Down(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))      # Downsample2D(padding=0): pad right / bottom only


class Up(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, down):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(cin, cout), Resnet(cout, cout)])
        self.downsamplers = nn.ModuleList([Down(cout)]) if down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.downsamplers[0](x) if self.downsamplers is not None else x


class UpBlock(nn.Module):
    def __init__(self, cin, cout, up):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(cin, cout), Resnet(cout, cout), Resnet(cout, cout)])
        self.upsamplers = nn.ModuleList([Up(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.upsamplers[0](x) if self.upsamplers is not None else x


class Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        w = WIDTHS
        self.conv_in = nn.Conv2d(3, w[0], 3, padding=1)
        self.down_blocks = nn.ModuleList([DownBlock(w[max(i - 1, 0)], w[i], i < 3) for i in range(4)])
        self.mid_block = Mid(w[3])
        self.conv_norm_out = nn.GroupNorm(32, w[3], eps=1e-6)
        self.conv_out = nn.Conv2d(w[3], 8, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(self.mid_block(x))))


class Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        w = WIDTHS
        self.conv_in = nn.Conv2d(4, w[3], 3, padding=1)
        self.mid_block = Mid(w[3])
        rev = (512, 512, 256, 128)
        self.up_blocks = nn.ModuleList([UpBlock(rev[max(i - 1, 0)], rev[i], i < 3) for i in range(4)])
        self.conv_norm_out = nn.GroupNorm(32, w[0], eps=1e-6)
        self.conv_out = nn.Conv2d(w[0], 3, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLOracle(nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = Encoder()
        self.decoder = Decoder()
        self.quant_conv = nn.Conv2d(8, 8, 1)
        self.post_quant_conv = nn.Conv2d(4, 4, 1)

    @torch.no_grad()
    def encode_moments(self, image):
        return self.quant_conv(self.encoder(image))

    @torch.no_grad()
    def encode(self, image, noise):
        """image in [-1, 1] (B,3,H,W); noise (B,4,H/8,W/8) -> latent_dist.sample() * scaling_factor."""
        m = self.encode_moments(image)
        mean, logvar = m[:, :4], m[:, 4:].clamp(-30.0, 20.0)
        return (mean + torch.exp(0.5 * logvar) * noise) * SCALING

    @torch.no_grad()
    def decode(self, latents):
        """latents as the scheduler leaves them -> image in ~[-1, 1]."""
        return self.decoder(self.post_quant_conv(latents / SCALING))
