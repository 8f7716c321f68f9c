"""Host logic of the per-GPU frame dispatcher (videosd_b200/dispatcher.py) and of the drop-in class around it, with a
recording stand-in for the CUDA engine: lane election, batch merging, per-slot context / seed routing, error propagation,
the Ray path of VideoSDPipeline.remote, and the no-silent-fallback rules. No GPU, no libvideosd.so compute call."""
import sys
import threading
import time
import types

import numpy as np
import pytest
import torch

from videosd_b200 import dispatcher as D


class FakeEngine:
    """Stands in for videosd_b200.engine.Engine: output plane = input plane + context id of the slot + seed-derived noise id."""
    instances = []

    def __init__(self, device=0, parent=None):
        self.device, self.parent = device, parent
        self.ctx, self.noise, self.calls, self.closed = {}, None, [], False
        self._tune_for = 1
        FakeEngine.instances.append(self)

    def set_autotune(self, n): self._tune_for = n
    def configure(self, nb, h, w): self.batch, self.height, self.width = nb, h, w
    def set_schedule(self, strength, steps, g=7.5): return [499, 379, 259, 139][:steps]
    def tuning_load(self, text): return 0
    def tuning_misses(self): return 0
    def tuning_report(self): return ""
    def set_context(self, slot, emb): self.ctx[slot] = int(emb[0, 0])
    def set_noise(self, init, steps): self.noise = init.clone()
    def set_controlnet(self, on, scale): self.calls.append(("cn", scale))
    def set_vae(self, kind): self.calls.append(("vae", kind))
    def set_vae_noise(self, n): self.calls.append(("vae_noise", tuple(n.shape)))
    def set_resize(self, w, h): self._resize_key = (w, h, self.width, self.height, self.batch)
    def close(self): self.closed = True

    fail = None

    def infer_yuv420(self, y, u, v, oy, ou, ov):
        time.sleep(0.01)
        if FakeEngine.fail is not None:
            raise FakeEngine.fail
        self.calls.append(("yuv", y.shape[0]))
        for b in range(y.shape[0]):
            oy[b] = y[b] + self.ctx[b]
            ou[b] = u[b]
            ov[b] = v[b] + int(self.noise[b, 0, 0, 0] * 0)   # the slot's noise exists
    infer_yuv420_resized = infer_yuv420

    def infer_rgb(self, a, o):
        self.calls.append(("rgb", a.shape[0]))
        o.copy_(a)
    infer_rgb_resized = infer_rgb


@pytest.fixture()
def fake(monkeypatch):
    FakeEngine.instances.clear()
    FakeEngine.fail = None
    monkeypatch.setattr(D, "Engine", FakeEngine)
    return FakeEngine


def _req(i, ctx_id, seed=42, h=16, w=16):
    y = np.full((1, h, w), i, np.uint8)
    u = np.full((1, h // 2, w // 2), 100 + i, np.uint8)
    v = np.full((1, h // 2, w // 2), 200, np.uint8)
    emb = torch.full((77, 768), float(ctx_id))
    return D.FrameRequest("yuv", (y, u, v), 1, h, w, h, w, 0.5, 4, seed, ("prompt", f"c{ctx_id}"), emb, 1.0)


def test_concurrent_sessions_are_batched_and_routed_to_their_own_slots(fake):
    disp = D.FrameDispatcher(fake(0), frames_in_flight=1, max_batch=4, noise_mode="reference_cpu")
    results = {}

    def session(i):
        for f in range(6):
            ctx_id = (i + f // 2) % 5                      # prompt switches mid-stream
            oy, ou, ov = disp.run(_req(10 * i + f, ctx_id))
            results[(i, f)] = (int(oy[0, 0, 0]), int(ou[0, 0, 0]), ctx_id)

    ths = [threading.Thread(target=session, args=(i,)) for i in range(8)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert len(results) == 48
    for (i, f), (oy, ou, ctx_id) in results.items():       # every session got ITS frame conditioned on ITS prompt
        assert oy == 10 * i + f + ctx_id and ou == 100 + 10 * i + f
    assert disp.stats["frames"] == 48 and disp.stats["launches"] < 48 and disp.stats["merged"] > 0
    assert max(n for k, n in (c for e in fake.instances for c in e.calls) if k == "yuv") <= 4
    assert disp._free == disp.lanes and not disp._pending  # everything handed back


def test_lanes_run_concurrently_and_share_the_weight_owner(fake):
    root = fake(0)
    disp = D.FrameDispatcher(root, frames_in_flight=3, max_batch=1, noise_mode="reference_cpu")
    assert root._tune_for == 3
    t0 = time.time()
    ths = [threading.Thread(target=lambda i=i: disp.run(_req(i, 1))) for i in range(9)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert time.time() - t0 < 9 * 0.01 * 0.9                # 9 x 10 ms on 3 lanes, not serial
    engines = {id(st.engine) for lane in disp.lanes for st in lane.states.values()}
    assert len(engines) == 3 and id(root) in engines        # lane 0 runs on the weight owner, the others are its lanes
    assert all(e.parent is root for e in fake.instances if e is not root)
    disp.close()
    assert all(e.closed for e in fake.instances if e is not root) and not root.closed


def test_engine_error_reaches_every_waiter_and_the_lane_is_released(fake):
    disp = D.FrameDispatcher(fake(0), frames_in_flight=1, max_batch=4, noise_mode="reference_cpu")
    boom = RuntimeError("device-side wait timed out")
    errs, oks = [], []

    def go(i):
        try:
            oks.append(disp.run(_req(i, 0)))
        except RuntimeError as e:
            errs.append(e)

    go(0)                                                   # a healthy frame first
    fake.fail = boom                                        # then the engine reports a fault (VsdError in the product)
    ths = [threading.Thread(target=go, args=(i,)) for i in range(1, 5)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert len(oks) == 1 and len(errs) == 4 and all(e is boom for e in errs)   # leaders and merged followers alike
    assert disp._free == disp.lanes and not disp._pending   # never hang: the lane is handed back (server.py:110-111)
    fake.fail = None
    go(5)
    assert len(oks) == 2


def test_same_frame_and_options_do_not_depend_on_the_batch(fake):
    disp = D.FrameDispatcher(fake(0), 1, 4, noise_mode="reference_cpu")
    init1, steps1, _ = disp._noise(1, 8, 8, (7,), 4, 0)
    init3, steps3, _ = disp._noise(3, 8, 8, (7, 9, 7), 4, 0)
    assert torch.equal(init3[0:1], init1) and torch.equal(init3[2:3], init1) and len(steps1) == len(steps3) == 4
    assert all(torch.equal(s3[1:2], s1) for s1, s3 in zip(steps1, steps3))
    # the CPU stream: init noise is draw 0 of a fresh generator, step noise i is draw i + 1 (videopipeline.py:126, :331, :1033)
    g = torch.Generator()
    draws = [torch.randn((1, 4, 8, 8), generator=g) for _ in range(5)]
    assert torch.equal(init1, draws[0]) and all(torch.equal(a, b) for a, b in zip(steps1, draws[1:]))
    assert disp._noise(1, 8, 8, (7,), 1, 0)[1] == []        # single step: no step noise


def test_remote_uses_ray_when_ray_imports(monkeypatch):
    from videosd_b200 import videopipeline as V

    seen = {}

    class _Actor:
        def __init__(self, cls): self.cls = cls
        def remote(self, *a, **k):
            seen["ctor"] = (a, k)
            return "actor-handle"

    fake_ray = types.SimpleNamespace(remote=lambda **opts: (lambda cls: (seen.update(opts=opts, cls=cls), _Actor(cls))[1]))
    monkeypatch.setitem(sys.modules, "ray", fake_ray)
    monkeypatch.setattr(V, "_HAVE_RAY", True)
    h = V.VideoSDPipeline.remote(model="m", controlnet="c", gpus=4, compile=False, frames_in_flight=3)
    assert h == "actor-handle"
    assert seen["opts"] == {"num_gpus": 1, "num_cpus": 4, "max_concurrency": 3}     # videopipeline.py:11 + threaded actor
    assert issubclass(seen["cls"], V.VideoSDPipeline) and seen["ctor"][1]["model"] == "m"


def test_no_silent_stand_ins_for_real_checkpoints(tmp_path):
    from videosd_b200 import tokenizer
    from videosd_b200.videopipeline import VideoSDPipeline

    with pytest.raises(FileNotFoundError):
        tokenizer.load(str(tmp_path))                        # no vocab.json / merges.txt -> no HashTokenizer
    assert tokenizer.load(None, allow_hash=True)("a b")[0] == 49406
    p = VideoSDPipeline.__new__(VideoSDPipeline)             # host-side conditioning logic only
    p.prompt_encoder, p.use_text_encoder, p.random_init = None, False, False
    p._emb_cache, p._emb_lock = {}, threading.Lock()
    with pytest.raises(RuntimeError, match="never replaced"):
        p._embedding("pixar, cg", None)
    key, emb = p._embedding("ignored", torch.ones(1, 77, 768))
    assert key[0] == "emb" and emb.shape == (77, 768)
    p.random_init = True
    k1, e1 = p._embedding(["pixar, cg"], None)
    k2, e2 = p._embedding(["pixar, cg"], None)
    assert k1 == k2 and e1 is e2                             # cached per prompt: the text tower runs once per prompt change
