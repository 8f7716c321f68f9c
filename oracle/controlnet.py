"""Oracle restatement of diffusers' ControlNetModel (lllyasviel/control_v11p_sd15_canny) and of the reference's Sobel edge
operator (TEST INFRASTRUCTURE ONLY).

* ControlNetOracle: SURVEY.md Appendix A.8 [diffusers-knowledge]; called by the reference every step at
  diffusert/lcm/lcm_controlnet.py:558-566 with guess_mode=True (:399,:447) and conditioning_scale =
  controlnet_scale * keep (:553-556). State-dict keys follow diffusers (conv_in, time_embedding,
  controlnet_cond_embedding.{conv_in,blocks.N,conv_out}, down_blocks.*, mid_block.*, controlnet_down_blocks.N,
  controlnet_mid_block).
* sobel_control_image: diffusert/lcm/canny_gpu.py:6-44 (SobelOperator) followed by the control-image preprocessing of
  lcm_controlnet.py:218-248 (VaeImageProcessor(do_convert_rgb=True, do_normalize=False)): PIL "L" -> 3 equal channels in [0,1].
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from PIL import Image

from .unet import DownBlock, MidBlock, timestep_sinusoid


class TimestepEmbeddingNoCond(nn.Module):
    def __init__(self, in_dim=320, dim=1280):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, sample):
        return self.linear_2(F.silu(self.linear_1(sample)))


class ControlNetConditioningEmbedding(nn.Module):
    def __init__(self, out_channels=320, cond_channels=3, widths=(16, 32, 96, 256)):
        super().__init__()
        self.conv_in = nn.Conv2d(cond_channels, widths[0], 3, padding=1)
        blocks = []
        for i in range(len(widths) - 1):
            blocks.append(nn.Conv2d(widths[i], widths[i], 3, padding=1))
            blocks.append(nn.Conv2d(widths[i], widths[i + 1], 3, padding=1, stride=2))
        self.blocks = nn.ModuleList(blocks)
        self.conv_out = nn.Conv2d(widths[-1], out_channels, 3, padding=1)   # zero-initialised in real checkpoints

    def forward(self, cond):
        x = F.silu(self.conv_in(cond))
        for b in self.blocks:
            x = F.silu(b(x))
        return self.conv_out(x)


class ControlNetOracle(nn.Module):
    widths = (320, 640, 1280, 1280)

    def __init__(self):
        super().__init__()
        w = self.widths
        self.conv_in = nn.Conv2d(4, w[0], 3, padding=1)
        self.time_embedding = TimestepEmbeddingNoCond(w[0], w[0] * 4)
        self.controlnet_cond_embedding = ControlNetConditioningEmbedding(w[0])
        downs, cin = [], w[0]
        for i, c in enumerate(w):
            downs.append(DownBlock(cin, c, attn=(i < 3), add_down=(i < 3)))
            cin = c
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(w[-1])
        chans = [w[0]]
        for i, c in enumerate(w):
            chans += [c, c] + ([c] if i < 3 else [])
        self.controlnet_down_blocks = nn.ModuleList([nn.Conv2d(c, c, 1) for c in chans])
        self.controlnet_mid_block = nn.Conv2d(w[-1], w[-1], 1)

    def forward(self, sample, timesteps, encoder_hidden_states, controlnet_cond, conditioning_scale=1.0, guess_mode=True):
        emb = self.time_embedding(timestep_sinusoid(timesteps, self.widths[0]))
        x = self.conv_in(sample) + self.controlnet_cond_embedding(controlnet_cond)
        feats = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            feats += outs
        x = self.mid_block(x, emb, encoder_hidden_states)
        down = [conv(f) for conv, f in zip(self.controlnet_down_blocks, feats)]
        mid = self.controlnet_mid_block(x)
        if guess_mode:
            scales = torch.logspace(-1, 0, len(down) + 1, device=sample.device) * conditioning_scale   # 0.1 .. 1.0
            down = [d * s for d, s in zip(down, scales)]
            mid = mid * scales[-1]
        else:
            down = [d * conditioning_scale for d in down]
            mid = mid * conditioning_scale
        return down, mid


def sobel_edges(img_rgb_u8, low_threshold=0.11, high_threshold=0.8):
    """canny_gpu.py:27-44 on a CPU device: returns the PIL 'L' image the reference hands to the pipeline."""
    image_gray = Image.fromarray(np.asarray(img_rgb_u8)).convert("L")
    t = torch.from_numpy(np.asarray(image_gray, dtype=np.uint8).copy()).float().div(255)[None, None]   # ToTensor
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]).view(1, 1, 3, 3)
    ky = torch.tensor([[-1.0, -2.0, -1.0], [0.0, 0.0, 0.0], [1.0, 2.0, 1.0]]).view(1, 1, 3, 3)
    ex = F.conv2d(t, kx, padding=1)
    ey = F.conv2d(t, ky, padding=1)
    edge = torch.sqrt(ex ** 2 + ey ** 2)
    edge = edge / edge.max()
    edge[edge >= high_threshold] = 1.0
    edge[edge <= low_threshold] = 0.0
    u8 = edge[0, 0].mul(255).byte().numpy()          # ToPILImage: float -> mul(255).byte() (truncation)
    return Image.fromarray(u8, mode="L")


def control_image_tensor(edge_pil):
    """prepare_control_image (lcm_controlnet.py:218-248): convert('RGB') -> /255 -> NCHW, no normalisation."""
    a = np.asarray(edge_pil.convert("RGB"), dtype=np.uint8)
    return torch.from_numpy(a.astype(np.float32) / 255.0).permute(2, 0, 1)[None].contiguous()
