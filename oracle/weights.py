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


def build_taesd(seed=4321):
    prev = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = TAESD().eval()
    torch.random.set_rng_state(prev)
    for p in net.parameters():
        p.requires_grad_(False)
    return net


def random_context(batch=1, seed=7):
    """Stand-in for CLIPTextModel.last_hidden_state (B,77,768): the pipeline accepts prompt_embeds
    (lcm_controlnet.py:394, :143); no tokenizer vocabulary is available offline."""
    return torch.randn((batch, 77, 768), generator=torch.Generator().manual_seed(seed))
