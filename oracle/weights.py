"""Deterministic random-init weights for the oracle modules (TEST INFRASTRUCTURE ONLY).

There is no network for checkpoints, so BASELINE.json's configs use "random-init weights of that architecture":
seeded default torch.nn initialisation of the restated modules; the same tensors are handed to the CUDA engine.
GroupNorm/LayerNorm affine parameters are perturbed away from (1, 0) so that a kernel ignoring them fails parity.
"""
import torch

from .taesd import TAESD
from .unet import UNetLCM


def _perturb_norms(module, gen):
    for m in module.modules():
        if isinstance(m, (torch.nn.GroupNorm, torch.nn.LayerNorm)):
            with torch.no_grad():
                m.weight.add_(0.1 * torch.randn(m.weight.shape, generator=gen))
                m.bias.add_(0.1 * torch.randn(m.bias.shape, generator=gen))


def build_unet(seed=1234):
    prev = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = UNetLCM().eval()
    _perturb_norms(net, torch.Generator().manual_seed(seed + 1))
    torch.random.set_rng_state(prev)
    for p in net.parameters():
        p.requires_grad_(False)
    return net


def _calibrate_taesd(net, seed):
    """Rescales the two output convolutions so that a random-init TAESD behaves like a trained one in magnitude:
    latents of unit scale (so the input frame matters next to the injected noise) and a decoded image that spans
    [0, 1] (so the uint8 output is not a near-black frame and PSNR is a meaningful test). Deterministic."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        x = torch.rand((1, 3, 64, 64), generator=g) * 2 - 1
        enc_out = net.encoder.layers[-1]
        z = net.encode(x)
        s = 1.0 / float(z.std())
        enc_out.weight.mul_(s)
        enc_out.bias.sub_(float(z.mean())).mul_(s)
        zz = torch.randn((1, 4, 16, 16), generator=g) * 1.5
        dec_out = net.decoder.layers[-1]
        raw = (net.decode(zz) + 1) / 2  # layers(...) before the final *2-1
        k = 0.25 / float(raw.std())
        dec_out.weight.mul_(k)
        dec_out.bias.mul_(k).add_(0.5 - float(raw.mean()) * k)


def build_taesd(seed=4321):
    prev = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = TAESD().eval()
    _calibrate_taesd(net, seed + 1)
    torch.random.set_rng_state(prev)
    for p in net.parameters():
        p.requires_grad_(False)
    return net


def random_context(batch=1, seed=7):
    """Stand-in for CLIPTextModel.last_hidden_state (B,77,768): the pipeline accepts prompt_embeds
    (lcm_controlnet.py:394, :143); no tokenizer vocabulary is available offline."""
    return torch.randn((batch, 77, 768), generator=torch.Generator().manual_seed(seed))


def build_controlnet(seed=9876):
    """Seeded random init; the (in real checkpoints zero-initialised) output convolutions keep their random values so
    that the ControlNet actually perturbs the UNet in the parity tests."""
    from .controlnet import ControlNetOracle

    prev = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = ControlNetOracle().eval()
    _perturb_norms(net, torch.Generator().manual_seed(seed + 1))
    torch.random.set_rng_state(prev)
    for p in net.parameters():
        p.requires_grad_(False)
    return net


def build_vae_kl(seed=2222, probe_hw=64):
    """Seeded random AutoencoderKL, calibrated like the TAESD stand-in so the parity tests are not vacuous: latent means of
    O(1 / 0.18215) (so `sample * scaling_factor` has unit scale like real SD latents), posterior std ~ e^-3, and a decoder
    whose output spans roughly [-1, 1]."""
    from .autoencoder_kl import SCALING, AutoencoderKLOracle

    prev = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = AutoencoderKLOracle().eval()
    g = torch.Generator().manual_seed(seed + 1)
    _perturb_norms(net, g)
    with torch.no_grad():
        x = torch.rand((1, 3, probe_hw, probe_hw), generator=g) * 2 - 1
        m = net.encode_moments(x)
        mean_std = m[:, :4].std().item()
        net.quant_conv.weight[:4] *= (1.0 / SCALING) / max(mean_std, 1e-6)
        net.quant_conv.bias[:4] *= (1.0 / SCALING) / max(mean_std, 1e-6)
        net.quant_conv.weight[4:] *= 0.5 / max(m[:, 4:].std().item(), 1e-6)
        net.quant_conv.bias[4:] = -6.0
        z = net.encode(x, torch.randn((1, 4, probe_hw // 8, probe_hw // 8), generator=g))
        y = net.decode(z)
        s = 0.5 / max(y.std().item(), 1e-6)
        net.decoder.conv_out.weight *= s
        net.decoder.conv_out.bias.copy_((net.decoder.conv_out.bias - y.mean()) * s)
    torch.random.set_rng_state(prev)
    for p in net.parameters():
        p.requires_grad_(False)
    return net


class KLAdapter:
    """Gives AutoencoderKLOracle the (encode, decode, scaling_factor) surface oracle/pipeline.py uses: encode() is
    `latent_dist.sample()` with the supplied noise (lcm_controlnet.py:298-313)."""

    def __init__(self, net, vae_noise):
        from .autoencoder_kl import SCALING
        self.net, self.noise, self.scaling_factor = net, vae_noise, SCALING

    def encode(self, x):
        return self.net.encode(x, self.noise.to(x.device)) / self.scaling_factor

    def decode(self, z):
        return self.net.decoder(self.net.post_quant_conv(z))
