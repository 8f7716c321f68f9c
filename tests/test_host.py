"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the product's scheduler /
weight tables agree with the oracle, and the drop-in boundary keeps the reference's argument behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "videosd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vsd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from videosd_b200._lib import LIB_PATH, lib

    assert os.path.exists(LIB_PATH), "build the library first: python -m videosd_b200.build"
    L = lib()
    syms = _declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/videosd.h but not exported"
    L.vsd_abi_version.restype = ctypes.c_int
    assert L.vsd_abi_version() == 1


def test_no_gpu_means_loud_failure_not_fallback():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from videosd_b200._lib import VsdError
    from videosd_b200.engine import Engine

    with pytest.raises(VsdError):
        Engine(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "videosd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_product_schedule_equals_oracle():
    from oracle.scheduler import LCMSchedulerOracle, w_embedding
    from videosd_b200 import scheduler

    s, o = scheduler.LCMSchedule(), LCMSchedulerOracle()
    for strength, n in [(0.5, 4), (0.6, 4), (0.05, 4), (0.4, 20), (1.0, 4), (0.5, 1), (0.1, 12)]:
        ts = s.timesteps(strength, n)
        assert ts == o.set_timesteps(strength, n).tolist()
        sc = s.step_scalars(ts)
        for i in range(len(ts)):
            d = o.step_scalars(i)
            ref = np.array([float(d[k]) for k in ("sqrt_alpha", "sqrt_beta", "c_skip", "c_out", "sqrt_alpha_prev",
                                                  "sqrt_beta_prev")], dtype=np.float32)
            assert np.array_equal(sc[i], ref)
        a, b = s.add_noise_coeffs(ts[0])
        assert a == float(o.alphas_cumprod[ts[0]] ** 0.5) and b == float((1 - o.alphas_cumprod[ts[0]]) ** 0.5)
    assert np.array_equal(scheduler.guidance_embedding(7.5), w_embedding(torch.tensor([7.5]), 256)[0].numpy())
    with pytest.raises(ValueError):
        s.timesteps(0.5, 1001)


def test_product_noise_equals_oracle_rng_order():
    from oracle import pipeline
    from videosd_b200.scheduler import reference_cpu_noise

    a, b = reference_cpu_noise(2, 8, 12, 4), pipeline.frame_noise(2, 8, 12, 4)
    assert torch.equal(a[0], b[0]) and all(torch.equal(x, y) for x, y in zip(a[1], b[1]))


def test_weight_tables_equal_oracle_modules():
    from oracle.taesd import TAESD
    from oracle.unet import UNetLCM
    from videosd_b200 import weights

    u = {k: tuple(v.shape) for k, v in UNetLCM().state_dict().items()}
    assert u == {k: tuple(v) for k, v in weights.unet_param_shapes().items()}
    t = {k: tuple(v.shape) for k, v in TAESD().state_dict().items()}
    assert t == {k: tuple(v) for k, v in weights.taesd_param_shapes().items()}
    from oracle.controlnet import ControlNetOracle
    c = {k: tuple(v.shape) for k, v in ControlNetOracle().state_dict().items()}
    assert c == {k: tuple(v) for k, v in weights.controlnet_param_shapes().items()}
    sd = weights.random_state_dict(weights.taesd_param_shapes(), 1)
    sd2 = weights.random_state_dict(weights.taesd_param_shapes(), 1)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)


def test_boundary_ctor_contract():
    from videosd_b200.videopipeline import VideoSDPipeline

    with pytest.raises(KeyError):           # videopipeline.py:22-26: missing model / controlnet -> KeyError re-raised
        VideoSDPipeline(controlnet="x")
    with pytest.raises(KeyError):
        VideoSDPipeline(model="x")


def test_boundary_crop_resize_matches_reference_rule():
    from PIL import Image

    from videosd_b200.videopipeline import VideoSDPipeline

    img = Image.fromarray((np.random.RandomState(0).rand(480, 640, 3) * 255).astype(np.uint8))
    out = VideoSDPipeline._fit(img, 512, 512)
    assert out.size == (512, 512)
    # identity when the frame already has the working size (all BASELINE configs)
    same = VideoSDPipeline._fit(out, 512, 512)
    assert np.array_equal(np.asarray(same), np.asarray(out))
    wide = VideoSDPipeline._fit(img, 640, 360)
    assert wide.size == (640, 360)


def test_lanczos_tables_reproduce_pillow_bit_exactly():
    """The GPU resize applies host-computed windows / 22-bit coefficients; the same arithmetic in numpy must equal PIL."""
    from PIL import Image

    from videosd_b200 import resample
    from videosd_b200.videopipeline import VideoSDPipeline

    for (iw, ih, w, h) in [(640, 480, 512, 512), (1280, 720, 640, 360), (320, 240, 512, 512), (641, 479, 256, 256),
                           (640, 512, 512, 512), (512, 512, 512, 512)]:
        src = np.random.RandomState(iw).randint(0, 256, (ih, iw, 3)).astype(np.uint8)
        src[::5] = 255
        src[2::9] = 0
        ref = np.asarray(VideoSDPipeline._fit(Image.fromarray(src), w, h))
        got = resample.resize_reference_numpy(src, w, h)
        assert np.array_equal(got, ref), (iw, ih, w, h)
    plan = resample.resize_plan(512, 512, 512, 512)
    assert plan["identity"] and plan["crop"] == (0, 0, 512, 512)
    b, k, ks = resample.lanczos_coeffs(480, 512)
    assert ks == 7 and b.shape == (512, 2) and (k.sum(axis=1) - (1 << 22)).__abs__().max() <= 8   # rows sum to ~1.0


def test_clip_tokenizer_matches_transformers_on_a_synthetic_vocabulary(tmp_path):
    """The real vocab.json / merges.txt cannot be downloaded here; the BPE algorithm itself is checked against
    transformers.CLIPTokenizer on a small vocabulary (byte alphabet + a few merges)."""
    import json

    transformers = pytest.importorskip("transformers")
    from videosd_b200 import tokenizer as T

    chars = sorted(set(T._bytes_to_unicode().values()))
    vocab = {}
    for c in chars:
        vocab[c] = len(vocab)
    for c in chars:
        vocab[c + "</w>"] = len(vocab)
    vocab["</w>"] = len(vocab)
    merges = [("p", "i"), ("pi", "x"), ("a", "r</w>"), ("pix", "ar</w>"), ("c", "g</w>"), ("t", "h"), ("th", "e</w>"), ("c", "a"),
              ("ca", "t</w>"), ("o", "n</w>"), ("'", "s</w>"), ("m", "a"), ("ma", "t</w>")]
    for a, b in merges:
        vocab[a + b] = len(vocab)
    vocab["<|startoftext|>"] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    (tmp_path / "merges.txt").write_text("#version: 0.2\n" + "\n".join(a + " " + b for a, b in merges) + "\n")
    hf = transformers.CLIPTokenizer(str(tmp_path / "vocab.json"), str(tmp_path / "merges.txt"))
    mine = T.load(str(tmp_path))
    assert isinstance(mine, T.ClipTokenizer)
    for text in ["pixar, cg", "The cat's on   the 12 mat!!", "a photo of a caf\u00e9 \u2014 na\u00efve", "x" * 200, "",
                 "hello_world __ it's 3.14%"]:
        assert mine(text) == hf(text, padding="max_length", max_length=77, truncation=True).input_ids, text
    h = T.load(None, allow_hash=True)                       # stand-in without a vocabulary: framing and determinism
    ids = h("pixar, cg")
    assert len(ids) == 77 and ids[0] == 49406 and ids[4:] == [49407] * 73 and ids == h("Pixar,  CG")


def test_autoencoder_kl_weight_table_matches_oracle_module():
    from oracle.autoencoder_kl import AutoencoderKLOracle
    from videosd_b200 import weights

    net = AutoencoderKLOracle()
    want = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert weights.autoencoder_kl_param_shapes() == want
    assert sum(p.numel() for p in net.parameters()) == 83_653_863


def test_clip_weight_table_matches_oracle_module():
    from oracle.clip import ClipTextOracle
    from videosd_b200 import weights

    want = {k: tuple(v.shape) for k, v in ClipTextOracle().state_dict().items()}
    assert weights.clip_param_shapes() == want


def test_lanczos_tables_vs_pillow_random_geometries():
    """Property test: for random source / working sizes the host tables + fixed-point arithmetic the CUDA kernels apply
    reproduce Pillow's crop + LANCZOS resize bit for bit (the reference's own CPU path, videopipeline.py:92-107)."""
    from PIL import Image

    from videosd_b200 import resample
    from videosd_b200.videopipeline import VideoSDPipeline

    rs = np.random.RandomState(1234)
    for _ in range(40):
        iw, ih = int(rs.randint(9, 400)), int(rs.randint(9, 400))
        w, h = 8 * int(rs.randint(2, 33)), 8 * int(rs.randint(2, 33))
        src = rs.randint(0, 256, (ih, iw, 3)).astype(np.uint8)
        if rs.rand() < 0.5:
            src[:: int(rs.randint(2, 9))] = 255 * int(rs.randint(0, 2))      # hard edges: ringing must clamp like Pillow
        ref = np.asarray(VideoSDPipeline._fit(Image.fromarray(src), w, h))
        got = resample.resize_reference_numpy(src, w, h)
        assert got.shape == ref.shape and np.array_equal(got, ref), (iw, ih, w, h)
        plan = resample.resize_plan(iw, ih, w, h)
        x0, y0, cw, ch = plan["crop"]
        assert 0 <= x0 and 0 <= y0 and x0 + cw <= iw and y0 + ch <= ih
        hb, hk, hks = plan["h"]
        assert hb.shape == (w, 2) and hk.shape == (w, hks) and (hb[:, 0] + hb[:, 1] <= cw).all()


def test_session_router_pins_and_batches():
    from videosd_b200.parallel import SessionRouter, shard_streams

    r = SessionRouter(num_gpus=2, max_batch=4)
    owners = [r.assign(f"s{i}") for i in range(8)]
    assert owners == [0, 1, 0, 1, 0, 1, 0, 1] and r.assign("s0") == 0
    b = r.batches([f"s{i}" for i in range(8)] + ["s8", "s9"])
    assert all(len(x) <= 4 for g in b.values() for x in g)
    assert sorted(s for g in b.values() for x in g for s in x) == sorted(f"s{i}" for i in range(10))
    r.release("s0")
    assert r.assign("new") == 0
    all_streams = sorted(s for k in range(4) for s in shard_streams(32, 4, k))
    assert all_streams == list(range(32))


def test_safetensors_directory_loader(tmp_path):
    """weights.load_safetensors_dir: every *.safetensors file of a diffusers component directory, as fp32 tensors."""
    from safetensors.torch import save_file

    from videosd_b200 import weights

    d = tmp_path / "vae"
    d.mkdir()
    sd = weights.random_state_dict(weights.taesd_param_shapes(), 3)
    keys = sorted(sd)
    save_file({k: sd[k].to(torch.float16) for k in keys[:40]}, str(d / "a.safetensors"))
    save_file({k: sd[k] for k in keys[40:]}, str(d / "b.safetensors"))
    got = weights.load_safetensors_dir(str(d))
    assert sorted(got) == keys and all(v.dtype == torch.float32 for v in got.values())
    assert torch.equal(got[keys[-1]], sd[keys[-1]]) and torch.equal(got[keys[0]], sd[keys[0]].to(torch.float16).float())
    with pytest.raises(FileNotFoundError):
        weights.load_safetensors_dir(str(tmp_path))


def test_tuning_table_resolution_prefers_exact_then_nearest_smaller_frames_in_flight():
    """Engine.configure loads the committed GEMM table of (size, batch, frames in flight); a count without its own table uses
    the nearest smaller one (never a live timing run, never a table of another batch / size)."""
    import os

    from videosd_b200.engine import TUNING_DIR, resolve_tuning_table

    def name(p):
        return None if p is None else os.path.basename(p)

    assert name(resolve_tuning_table(1, 512, 512, 6)) == "512x512x1_n6.txt"      # bench.py default
    assert name(resolve_tuning_table(1, 512, 512, 5)) == "512x512x1_n4.txt"
    assert name(resolve_tuning_table(1, 512, 512, 64)) == "512x512x1_n6.txt"
    assert name(resolve_tuning_table(4, 512, 512, 3)) == "512x512x4_n3.txt"      # sessions config
    assert name(resolve_tuning_table(4, 768, 768, 2)) == "768x768x4_n1.txt"
    assert resolve_tuning_table(7, 512, 512, 1) is None and resolve_tuning_table(1, 520, 512, 1) is None
    # every committed table parses: "<shape key> bn=.. splits=.. occ=.. kbs=.. halo=.. us=.."
    for f in sorted(os.listdir(TUNING_DIR)):
        for ln in open(os.path.join(TUNING_DIR, f)):
            parts = ln.split()
            assert len(parts) == 7 and parts[0].count("|") == 5 and [p.split("=")[0] for p in parts[1:]] == ["bn", "splits", "occ", "kbs", "halo", "us"], (f, ln)
