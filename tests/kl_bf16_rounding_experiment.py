"""CPU experiment behind DESIGN.md 4e: how far can bf16 storage take the AutoencoderKL encoder from the fp32 oracle?
Emulates the CUDA data path on the oracle (bf16 weights, bf16 rounding of every convolution / GroupNorm / resnet output) at
256x256 with the test weights and prints the max-norm relative error of the sampled latents. Result (this container):
all roundings 2.45e-2 (the CUDA path measures 2.6e-2), bf16 WEIGHTS ALONE 0.97e-2, activations alone 1.9e-2 -- the
1e-2 bar of the TAESD path is not reachable with bf16 operands for this 22-resnet encoder; it needs fp16 operands (the
reference dtype). Test infrastructure only (imports oracle/).  python tests/kl_bf16_rounding_experiment.py"""
import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import imageproc
from oracle.weights import build_vae_kl
import oracle.autoencoder_kl as kl
import torch.nn as nn, torch.nn.functional as F
torch.set_num_threads(16)
H=W=256
net=build_vae_kl()
y,u,v=imageproc.synthetic_frame(H,W,seed=3,shift=5)
rgb=imageproc.yuv420_to_rgb(y,u,v)[None]
img=torch.from_numpy(rgb).float().permute(0,3,1,2)/127.5-1.0
noise=torch.randn((1,4,H//8,W//8),generator=torch.Generator().manual_seed(99))
ref=net.encode(img,noise)
def rel(a,b): return float((a-b).abs().max()/b.abs().max())
def bf(x): return x.to(torch.bfloat16).float()
import copy
def run(round_act=True, round_w=True, round_res=True, round_gn=True, attn_p_bf16=True):
    n2=copy.deepcopy(net)
    hooks=[]
    for name,m in n2.named_modules():
        if isinstance(m,(nn.Conv2d,nn.Linear)):
            if round_w and not name.endswith('encoder.conv_in') and 'quant_conv' not in name:
                m.weight.data=bf(m.weight.data)
            if round_act and 'quant_conv' not in name and not name.endswith('encoder.conv_out'):
                hooks.append(m.register_forward_hook(lambda mod,i,o: bf(o)))
        if isinstance(m,nn.GroupNorm) and round_gn:
            hooks.append(m.register_forward_hook(lambda mod,i,o: bf(o)))
        if isinstance(m,kl.Resnet) and round_res:
            hooks.append(m.register_forward_hook(lambda mod,i,o: bf(o)))
    out=n2.encode(img,noise)
    return rel(out,ref)
print('all bf16 roundings', run())
print('weights only', run(round_act=False,round_res=False,round_gn=False))
print('acts only (no weights)', run(round_w=False))
print('weights+conv acts, no gn/res rounding', run(round_res=False,round_gn=False))
m=net.encode_moments(img)
print('mean absmax',float(m[:,:4].abs().max()),'logvar range',float(m[:,4:].min()),float(m[:,4:].max()), 'latent absmax', float(ref.abs().max()))
