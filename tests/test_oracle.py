"""CPU tests of the oracle (the checker): golden vectors produced by the reference's own code
(tests/golden/make_reference_golden.py), structural identities and the colour spec."""
import pytest
import numpy as np
import torch

from oracle import imageproc, pipeline
from oracle.scheduler import LCMSchedulerOracle, w_embedding
from oracle.taesd import TAESD
from oracle.weights import random_context


# ------------------------------------------------------------------ golden: the reference scheduler itself
def test_alphas_cumprod_match_reference(golden):
    s = LCMSchedulerOracle()
    assert np.array_equal(s.alphas_cumprod.numpy(), golden["alphas_cumprod"])


def test_timestep_tables_match_reference(golden):
    s = LCMSchedulerOracle()
    for k, (strength, steps) in enumerate(golden["table_cfg"]):
        assert s.set_timesteps(float(strength), int(steps)).tolist() == golden[f"table_{k}"].tolist()
    # SURVEY.md Appendix C spot values
    assert s.set_timesteps(0.5, 4).tolist() == [499, 379, 259, 139]
    assert s.set_timesteps(0.05, 4).tolist() == [39, 19]          # fewer timesteps than steps
    assert s.set_timesteps(0.4, 20).tolist()[:2] == [399, 379]


def test_step_matches_reference_bit_exact(golden):
    s = LCMSchedulerOracle()
    s.set_timesteps(0.5, 4)
    sample = torch.from_numpy(golden["step_sample"])
    mo = torch.from_numpy(golden["step_model_out"])
    for i in range(4):
        torch.manual_seed(1000 + i)
        prev, den = s.step(mo, i, sample)       # draws torch.randn from the global CPU RNG like the reference
        assert np.array_equal(prev.numpy(), golden[f"step_prev_{i}"])
        assert np.array_equal(den.numpy(), golden[f"step_den_{i}"])
    s.set_timesteps(0.5, 1)                      # single-step: no noise, prev == denoised
    prev, den = s.step(mo, 0, sample)
    assert np.array_equal(prev.numpy(), golden["single_prev"]) and np.array_equal(den.numpy(), golden["single_den"])


def test_add_noise_and_w_embedding_match_reference(golden):
    s = LCMSchedulerOracle()
    s.set_timesteps(0.5, 4)
    out = s.add_noise(torch.from_numpy(golden["step_sample"]), torch.from_numpy(golden["add_noise_noise"]),
                      s.timesteps[:1].repeat(2))
    assert np.array_equal(out.numpy(), golden["add_noise_out"])
    w = w_embedding(torch.tensor(7.5).repeat(2), 256)
    assert np.array_equal(w.numpy(), golden["w_embedding"])
    np.testing.assert_allclose(w[0, :3].numpy(), [-0.85123587, 0.84213120, 0.01187626], atol=2e-6)


def test_scheduler_constants_appendix_c():
    s = LCMSchedulerOracle()
    s.set_timesteps(0.5, 4)
    sc = s.step_scalars(0)
    assert abs(float(sc["sqrt_alpha"]) - 0.52694350) < 1e-7 and abs(float(sc["sqrt_beta"]) - 0.84990031) < 1e-7
    assert s.step_scalars(3)["prev_t"] == 139            # last step: prev_t = t
    assert abs(float(s.step_scalars(3)["c_out"]) - 0.99999994) < 1e-7


# ------------------------------------------------------------------ golden: the reference __call__ over the oracle modules
def test_pipeline_sequencing_matches_reference_call(golden, oracle_models):
    unet, vae = oracle_models
    rgb = golden["pipe_rgb_in"]
    ctx = random_context(1, seed=int(golden["pipe_ctx_seed"][0]))
    out = pipeline.lcm_img2img(unet, vae, rgb[None], ctx, steps=4, strength=0.5)
    assert out["timesteps"] == golden["pipe_timesteps"].tolist()
    for i in range(4):
        # identical modules and op order; only host thread count may reorder fp32 reductions
        np.testing.assert_allclose(out["latents_in"][i].numpy(), golden[f"pipe_latents_in_{i}"], rtol=0, atol=2e-4)
        np.testing.assert_allclose(out["eps"][i].numpy(), golden[f"pipe_eps_{i}"], rtol=0, atol=2e-4)
    diff = np.abs(out["rgb"][0].astype(int) - golden["pipe_rgb_out"].astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.01


def test_frame_noise_is_constant_per_frame():
    a = pipeline.frame_noise(1, 8, 8, 4)
    b = pipeline.frame_noise(1, 8, 8, 4)
    assert torch.equal(a[0], b[0]) and all(torch.equal(x, y) for x, y in zip(a[1], b[1]))
    assert len(pipeline.frame_noise(1, 8, 8, 1)[1]) == 0     # single timestep: no step noise drawn


# ------------------------------------------------------------------ structure (Appendix A.6 / A.7)
def test_parameter_counts(oracle_models):
    unet, vae = oracle_models
    assert sum(p.numel() for p in unet.parameters()) == 859_602_884
    assert sum(p.numel() for p in vae.encoder.parameters()) == 1_222_532
    assert sum(p.numel() for p in vae.decoder.parameters()) == 1_222_531


def test_state_dict_key_names(oracle_models):
    unet, vae = oracle_models
    k = set(unet.state_dict().keys())
    assert len(k) == 687
    for name in ["conv_in.weight", "time_embedding.cond_proj.weight", "time_embedding.linear_2.bias",
                 "down_blocks.0.attentions.1.transformer_blocks.0.attn2.to_k.weight",
                 "down_blocks.2.downsamplers.0.conv.bias", "mid_block.attentions.0.proj_out.weight",
                 "up_blocks.1.resnets.2.conv_shortcut.weight", "up_blocks.3.attentions.2.transformer_blocks.0.ff.net.0.proj.bias",
                 "up_blocks.2.upsamplers.0.conv.weight", "up_blocks.0.resnets.0.time_emb_proj.weight", "conv_norm_out.bias"]:
        assert name in k, name
    assert "down_blocks.3.attentions.0.norm.weight" not in k and "up_blocks.0.attentions.0.norm.weight" not in k
    assert "time_embedding.cond_proj.bias" not in k
    assert unet.state_dict()["up_blocks.1.resnets.2.conv1.weight"].shape == (1280, 1920, 3, 3)
    assert unet.state_dict()["up_blocks.3.resnets.0.conv1.weight"].shape == (320, 960, 3, 3)
    tk = set(TAESD().state_dict().keys())
    assert "encoder.layers.2.weight" in tk and "encoder.layers.2.bias" not in tk      # strided convs have no bias
    assert "decoder.layers.6.weight" in tk and "decoder.layers.6.bias" not in tk
    assert "encoder.layers.1.conv.4.bias" in tk and "decoder.layers.18.bias" in tk


def test_unet_odd_latent_size_runs(oracle_models):
    unet, _ = oracle_models
    x = torch.randn(1, 4, 6, 10)                       # not a multiple of 8: exercises upsample-to-skip-size
    out = unet(x, torch.tensor([499]), w_embedding(torch.tensor([7.5])), random_context(1))
    assert out.shape == x.shape and torch.isfinite(out).all()


# ------------------------------------------------------------------ image I/O spec (Appendix D)
def test_preprocess_postprocess_spec():
    u8 = np.arange(256, dtype=np.uint8).reshape(1, 16, 16, 1).repeat(3, axis=3)
    x = imageproc.preprocess(u8)
    assert x.shape == (1, 3, 16, 16) and x.dtype == torch.float32
    assert float(x.min()) == -1.0 and float(x.max()) == 1.0
    assert np.array_equal(imageproc.postprocess(x), u8)               # exact round trip for every code value
    t = torch.tensor([-3.0, 9.0, 0.0, 1.0]).view(1, 1, 1, 4).repeat(1, 3, 1, 1)
    assert imageproc.postprocess(t)[0, 0, :, 0].tolist() == [0, 255, 128, 255]   # clamp; 127.5 rounds half-to-even


def test_yuv_rgb_known_answers():
    def one(y, u, v):
        return imageproc.yuv420_to_rgb(np.full((2, 2), y, np.uint8), np.full((1, 1), u, np.uint8),
                                       np.full((1, 1), v, np.uint8))[0, 0].tolist()
    assert one(16, 128, 128) == [0, 0, 0]                 # limited-range black
    assert one(235, 128, 128) == [255, 255, 255]          # limited-range white
    # BT.601 primaries (hand-evaluated from the integer matrices in oracle/imageproc.py)
    assert one(82, 90, 240) == [255, 1, 0]                # (298*66 + 409*112 + 128) >> 8 = 255 ; G = 1 ; B < 0 -> 0
    assert one(144, 54, 34) == [0, 254, 0]
    assert one(41, 240, 110) == [0, 0, 255]
    y, u, v = imageproc.rgb_to_yuv420(np.array([[[255, 0, 0]] * 2] * 2, np.uint8))
    assert (int(y[0, 0]), int(u[0, 0]), int(v[0, 0])) == (82, 90, 240)
    y, u, v = imageproc.rgb_to_yuv420(np.array([[[0, 255, 0]] * 2] * 2, np.uint8))
    assert (int(y[0, 0]), int(u[0, 0]), int(v[0, 0])) == (144, 54, 34)
    y, u, v = imageproc.rgb_to_yuv420(np.zeros((2, 2, 3), np.uint8))
    assert (int(y[0, 0]), int(u[0, 0]), int(v[0, 0])) == (16, 128, 128)


def test_yuv_rgb_round_trip_property():
    y, u, v = imageproc.synthetic_frame(64, 96, seed=5)
    rgb = imageproc.yuv420_to_rgb(y, u, v)
    y2, u2, v2 = imageproc.rgb_to_yuv420(rgb)
    assert np.abs(y2.astype(int) - y.astype(int)).max() <= 2       # integer matrices invert to within rounding
    assert np.abs(u2.astype(int) - u.astype(int)).max() <= 2 and np.abs(v2.astype(int) - v.astype(int)).max() <= 2
    assert y.min() >= 16 and y.max() <= 235


# ------------------------------------------------------------------ ControlNet row
def test_controlnet_structure_and_reference_golden(golden, oracle_models):
    from oracle.controlnet import sobel_edges
    from oracle.weights import build_controlnet

    cn = build_controlnet()
    assert sum(p.numel() for p in cn.parameters()) == 361_279_120          # Appendix A.6
    keys = set(cn.state_dict().keys())
    for k in ["controlnet_cond_embedding.conv_in.weight", "controlnet_cond_embedding.blocks.5.bias",
              "controlnet_cond_embedding.conv_out.weight", "controlnet_down_blocks.11.weight", "controlnet_mid_block.bias",
              "down_blocks.2.attentions.1.transformer_blocks.0.attn2.to_v.weight", "mid_block.resnets.1.conv2.bias"]:
        assert k in keys, k
    assert "time_embedding.cond_proj.weight" not in keys and not any(k.startswith("up_blocks") for k in keys)
    # Sobel: bit-exact against the output of the reference's own SobelOperator (lcm/canny_gpu.py)
    assert np.array_equal(np.array(sobel_edges(golden["sobel_rgb_in"])), golden["sobel_out"])
    assert np.array_equal(np.array(sobel_edges(golden["pipe_rgb_in"])), golden["sobel_out_64"])
    # the reference __call__ with the ControlNet plugged in: conditioning scale passed through, guess mode on every step
    assert golden["cn_call_scales"].tolist() == [0.7] * 4 and golden["cn_guess_mode"].all()
    unet, vae = oracle_models
    out = pipeline.lcm_img2img(unet, vae, golden["pipe_rgb_in"][None], random_context(1, seed=int(golden["pipe_ctx_seed"][0])),
                               steps=4, strength=0.5, controlnet=cn, controlnet_scale=float(golden["cn_scale"][0]))
    np.testing.assert_allclose(out["control"].numpy(), golden["cn_control"], rtol=0, atol=0)
    for i in range(4):
        np.testing.assert_allclose(out["latents_in"][i].numpy(), golden[f"cn_latents_in_{i}"], rtol=0, atol=2e-4)
        np.testing.assert_allclose(out["eps"][i].numpy(), golden[f"cn_eps_{i}"], rtol=0, atol=2e-4)
    assert np.abs(out["rgb"][0].astype(int) - golden["cn_rgb_out"].astype(int)).max() <= 1


def test_clip_text_oracle_matches_transformers_clip_text_model():
    """Pins oracle/clip.py against the real third-party implementation the reference calls (transformers CLIPTextModel,
    lcm_controlnet.py:175-179), on the same seeded full-size weights: the text-encoder parity is NOT unpinned."""
    transformers = pytest.importorskip("transformers")
    from oracle.clip import ClipTextOracle
    from videosd_b200 import weights

    sd = weights.random_clip_state_dict(5)
    assert sum(v.numel() for v in sd.values()) == 123_060_480
    cfg = transformers.CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                                      num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu",
                                      layer_norm_eps=1e-5, projection_dim=768)
    hf = transformers.CLIPTextModel(cfg).eval()
    hf_keys = {k for k in hf.state_dict() if "position_ids" not in k}
    assert hf_keys == set(sd)                                           # same state-dict keys as the checkpoint format
    hf.load_state_dict(sd, strict=False)
    mine = ClipTextOracle()
    mine.load_state_dict(sd, strict=True)
    ids = torch.randint(0, 49408, (2, 77), generator=torch.Generator().manual_seed(1))
    ids[:, 0] = 49406
    ids[1, 9:] = 49407                                                  # a short prompt padded with <|endoftext|>
    with torch.no_grad():
        ref = hf(ids)[0]
    got = mine(ids)
    assert (ref - got).abs().max().item() < 1e-4
    # causal: positions before a changed token are unaffected
    ids2 = ids.clone()
    ids2[0, 40] = 123
    got2 = mine(ids2)
    assert torch.equal(got2[0, :40], got[0, :40]) and not torch.equal(got2[0, 40:], got[0, 40:])


def test_autoencoder_kl_oracle_structure_and_sampling():
    """AutoencoderKL restatement (parity unpinned: diffusers is not installable): shapes, the asymmetric stride-2 padding,
    `sample = mean + exp(0.5 * clamp(logvar)) * noise`, scaling factor, and the calibrated random init."""
    from oracle.autoencoder_kl import SCALING, AutoencoderKLOracle
    from oracle.weights import KLAdapter, build_vae_kl

    net = build_vae_kl()
    x = torch.rand((1, 3, 64, 48), generator=torch.Generator().manual_seed(0)) * 2 - 1
    m = net.encode_moments(x)
    assert m.shape == (1, 8, 8, 6)
    noise = torch.randn((1, 4, 8, 6), generator=torch.Generator().manual_seed(1))
    z = net.encode(x, noise)
    want = (m[:, :4] + torch.exp(0.5 * m[:, 4:].clamp(-30, 20)) * noise) * SCALING
    assert torch.allclose(z, want) and 0.3 < z.std().item() < 3.0
    assert torch.allclose(net.encode(x, torch.zeros_like(noise)), m[:, :4] * SCALING)
    img = net.decode(z)
    assert img.shape == (1, 3, 64, 48) and img.abs().max().item() < 4.0
    ad = KLAdapter(net, noise)
    assert torch.allclose(ad.encode(x) * ad.scaling_factor, z) and torch.allclose(ad.decode(z / ad.scaling_factor), img)
    d = AutoencoderKLOracle().encoder.down_blocks[0].downsamplers[0]
    y = d(torch.ones(1, 128, 8, 8))
    assert y.shape == (1, 128, 4, 4)          # pad right / bottom only, stride 2
