"""Oracle restatement of diffusers' AutoencoderTiny (TAESD, madebyollin/taesd) (TEST INFRASTRUCTURE ONLY).
Spec: SURVEY.md Appendix A.5 [diffusers-knowledge]. The reference swaps it in at diffusert/videopipeline.py:67-69
and calls it at diffusert/lcm/lcm_controlnet.py:298-300 (encode(...).latents) and :594-596 (decode).
State-dict keys: encoder.layers.N.{weight,bias} / encoder.layers.N.conv.{0,2,4}.{weight,bias}; same for decoder.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class TinyBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(c, c, 3, padding=1), nn.ReLU(), nn.Conv2d(c, c, 3, padding=1), nn.ReLU(),
                                  nn.Conv2d(c, c, 3, padding=1))
        self.skip = nn.Identity()
        self.fuse = nn.ReLU()

    def forward(self, x):
        return self.fuse(self.conv(x) + self.skip(x))


class TinyEncoder(nn.Module):
    def __init__(self, c=64, latent=4):
        super().__init__()
        layers = [nn.Conv2d(3, c, 3, padding=1), TinyBlock(c)]
        for _ in range(3):
            layers.append(nn.Conv2d(c, c, 3, padding=1, stride=2, bias=False))
            layers += [TinyBlock(c) for _ in range(3)]
        layers.append(nn.Conv2d(c, latent, 3, padding=1))
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        # AutoencoderTiny.encode feeds the [-1,1] image through (x+1)/2 first
        return self.layers((x + 1) / 2)


class TinyDecoder(nn.Module):
    def __init__(self, c=64, latent=4):
        super().__init__()
        layers = [nn.Conv2d(latent, c, 3, padding=1), nn.ReLU()]
        for n in (3, 3, 3):
            layers += [TinyBlock(c) for _ in range(n)]
            layers.append(nn.Upsample(scale_factor=2))
            layers.append(nn.Conv2d(c, c, 3, padding=1, bias=False))
        layers.append(TinyBlock(c))
        layers.append(nn.Conv2d(c, 3, 3, padding=1))
        self.layers = nn.Sequential(*layers)

    def forward(self, z):
        z = torch.tanh(z / 3) * 3
        return self.layers(z) * 2 - 1


class TAESD(nn.Module):
    scaling_factor = 1.0

    def __init__(self):
        super().__init__()
        self.encoder = TinyEncoder()
        self.decoder = TinyDecoder()

    def encode(self, x):
        return self.encoder(x)

    def decode(self, z):
        return self.decoder(z)
