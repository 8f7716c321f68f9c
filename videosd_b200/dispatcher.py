"""Per-GPU frame dispatcher: the product-side scheduler between the reference's frame-transform interface and the lanes.

The reference schedules frames in diffusert/server.py:132-137 (the first GPU whose `generating` flag is clear gets the
frame) onto one Ray actor per GPU (:318-321) that runs one frame at a time. Here one GPU serves several frames at once:

* **lanes** -- `frames_in_flight` engines on one copy of the weights (vsd_create_lane), each with its own stream, arena and
  CUDA graph; a request runs on the first idle lane;
* **batches** -- requests of different sessions that are waiting while every lane is busy are merged, up to `max_batch`
  frames with compatible geometry / schedule, into ONE launch (frame batch NB > 1: the deep UNet layers stream their
  weights once per batch); every slot keeps its own prompt context (vsd_set_context re-projects only that slot's
  cross-attention K / V) and its own seed;
* no extra threads: the calling threads (the worker pool behind `.remote()`, Ray's actor threads, or a direct caller) elect
  a leader per batch, the others wait for their result -- a session never sees another session's frame, and the result of
  a (frame, options) pair does not depend on what it was batched with (per-slot noise is the B = 1 draw of its seed).

Everything per-frame runs in libvideosd.so; this module only moves host buffers and keys caches.
"""
import collections
import hashlib
import threading

import numpy as np
import torch

from .engine import Engine


class FrameRequest:
    """One infer call: `nb` frames with the same options. kind 'rgb': data = u8 (nb, in_h, in_w, 3); 'yuv': data =
    (y (nb, in_h, in_w), u, v (nb, in_h/2, in_w/2))."""

    __slots__ = ("kind", "data", "nb", "in_h", "in_w", "height", "width", "strength", "steps", "seed", "prompt_key",
                 "embedding", "cn_scale", "result", "error", "done")

    def __init__(self, kind, data, nb, in_h, in_w, height, width, strength, steps, seed, prompt_key, embedding, cn_scale):
        self.kind, self.data, self.nb, self.in_h, self.in_w = kind, data, nb, in_h, in_w
        self.height, self.width, self.strength, self.steps, self.seed = height, width, strength, steps, seed
        self.prompt_key, self.embedding, self.cn_scale = prompt_key, embedding, cn_scale
        self.result = self.error = None
        self.done = False

    def compat(self):
        return (self.kind, self.in_h, self.in_w, self.height, self.width, self.strength, self.steps, self.cn_scale)


def embedding_key(emb):
    """Cache key of an explicit (77, 768) prompt embedding: a digest of its bytes (exact, not an identity check)."""
    a = np.ascontiguousarray(torch.as_tensor(emb).detach().to(torch.float32).cpu().numpy())
    return ("emb", hashlib.blake2b(a.tobytes(), digest_size=16).hexdigest())


class _EngineState:
    """What one configured engine currently holds, so that only what changed between two launches is re-sent."""

    def __init__(self, engine):
        self.engine = engine
        self.noise_key = None
        self.sched = self.timesteps = None
        self.slot_ctx = {}
        self.cn_scale = None
        self.vae_set = False
        self.pin = {}      # (kind, in_h, in_w) -> pinned staging tensors (inputs, outputs)


class _Lane:
    def __init__(self, index, root):
        self.index = index
        self.root = root               # the engine owning the weights
        self.states = {}               # (nb, height, width) -> _EngineState

    def state(self, nb, height, width, keep=4):
        """The engine of this lane configured for (nb, height, width); the `keep` most recently created ones stay alive
        (the UI changes the working size, batches come in several sizes). Lane 0 runs on the weight owner itself."""
        key = (nb, height, width)
        st = self.states.get(key)
        if st is None:
            while len(self.states) >= keep:
                dead = self.states.pop(next(iter(self.states)))
                if dead.engine is not self.root:
                    dead.engine.close()
            root_free = self.index == 0 and all(s.engine is not self.root for s in self.states.values())
            eng = self.root if root_free else Engine(self.root.device, parent=self.root)
            eng.configure(nb, height, width)
            st = self.states[key] = _EngineState(eng)
        return st

    def close(self):
        for st in self.states.values():
            if st.engine is not self.root:
                st.engine.close()
        self.states.clear()


class FrameDispatcher:
    def __init__(self, root_engine, frames_in_flight=1, max_batch=1, noise_mode="reference_cuda", use_controlnet=False,
                 vae_kind="taesd"):
        self.root = root_engine
        self.max_batch = max(1, int(max_batch))
        self.noise_mode = noise_mode
        self.use_controlnet = bool(use_controlnet)
        self.vae_kind = vae_kind
        n = max(1, int(frames_in_flight))
        root_engine.set_autotune(n)                 # GEMM configurations (and the committed table) for n frames in flight
        self.lanes = [_Lane(i, root_engine) for i in range(n)]
        self._free = list(reversed(self.lanes))
        self._pending = collections.deque()
        self._cv = threading.Condition()
        # Plans are built one at a time, and every engine starts from the GEMM choices of the ones built before it: with a
        # shape the committed tuning table does not cover, lanes that timed it independently could pick different
        # configurations and the same frame would differ in the last bit from lane to lane.
        self._plan_lock = threading.Lock()
        self._tuning_text = ""
        self.last_state = None
        self.stats = collections.Counter()          # launches, frames, merged (frames that rode in someone else's launch)

    # ------------------------------------------------------------------------------------------------ scheduling
    def run(self, req):
        """Blocks until `req` is done (executed by this thread as a batch leader, or by another leader)."""
        with self._cv:
            self._pending.append(req)
        while True:
            with self._cv:
                while not req.done and not (self._free and req in self._pending):
                    self._cv.wait()
                if req.done:
                    break
                lane = self._free.pop()
                batch = self._take_batch(req)
            try:
                self._execute(lane, batch)
            except BaseException as e:  # noqa: BLE001  -- every waiter must be released; the caller re-raises
                for r in batch:
                    r.error = e
            finally:
                with self._cv:
                    for r in batch:
                        r.done = True
                    self._free.append(lane)
                    self._cv.notify_all()
        if req.error is not None:
            raise req.error
        return req.result

    def _take_batch(self, leader):
        """Called with the lock held: the leader plus the oldest compatible single-frame requests, <= max_batch frames."""
        self._pending.remove(leader)
        batch = [leader]
        if self.max_batch > 1 and leader.nb == 1:
            key = leader.compat()
            for r in list(self._pending):
                if len(batch) >= self.max_batch:
                    break
                if r.nb == 1 and r.compat() == key:
                    self._pending.remove(r)
                    batch.append(r)
        return batch

    # ------------------------------------------------------------------------------------------------ one launch
    def _noise(self, nb, h8, w8, seeds, n_ts, device):
        """Per-slot noise = what the reference draws for a B = 1 call with that slot's seed: the result of a
        (frame, options) pair must not depend on what it is batched with. Returns (init, [step noises], vae noise | None)."""
        kl = self.vae_kind == "kl"
        shape = (1, 4, h8, w8)
        gc = torch.Generator()      # a fresh generator = the state videopipeline.py:126 re-arms the CPU global RNG with
        if self.noise_mode == "reference_cpu":
            # CPU device: latent_dist.sample() (AutoencoderKL only) and the init noise come from the CPU global RNG too
            vn1 = torch.randn(shape, generator=gc) if kl else None
            init1 = torch.randn(shape, generator=gc)
            init = init1.expand(nb, -1, -1, -1).contiguous()
            vae_noise = vn1.expand(nb, -1, -1, -1).contiguous() if kl else None
        else:
            # CUDA device: torch.manual_seed(seed) seeds the device Philox that draws the init noise in the model dtype
            # (lcm_controlnet.py:331; the generator argument is not forwarded, :503-513); AutoencoderKL's sample() draws
            # from the same generator first (:298-331). The CPU generator is untouched until scheduler.step (:1033).
            by_seed = {}
            for sd in seeds:
                if sd not in by_seed:
                    g = torch.Generator(device=f"cuda:{device}").manual_seed(int(sd))
                    draw = lambda: torch.randn(shape, generator=g, device=f"cuda:{device}", dtype=torch.float16).float().cpu()  # noqa: E731
                    vn = draw() if kl else None
                    by_seed[sd] = (draw(), vn)
            init = torch.cat([by_seed[sd][0] for sd in seeds], 0)
            vae_noise = torch.cat([by_seed[sd][1] for sd in seeds], 0) if kl else None
        steps = [torch.randn(shape, generator=gc).expand(nb, -1, -1, -1).contiguous() for _ in range(n_ts)] if n_ts > 1 else []
        return init, steps, vae_noise

    def _execute(self, lane, batch):
        lead = batch[0]
        nb = sum(r.nb for r in batch)
        H, W = lead.height, lead.width
        st = lane.state(nb, H, W)
        eng = st.engine
        if self.vae_kind == "kl" and not st.vae_set:
            eng.set_vae("kl")
            st.vae_set = True
        if self.use_controlnet and st.cn_scale != lead.cn_scale:
            eng.set_controlnet(True, lead.cn_scale)             # lcm_controlnet.py:553-556 (keep = 1.0)
            st.cn_scale = lead.cn_scale
        if st.sched != (lead.strength, lead.steps) or getattr(eng, "_sched_key", None) is None:
            st.timesteps = self._set_schedule(eng, lead.strength, lead.steps)     # (re)builds the launch plan
            st.sched = (lead.strength, lead.steps)
        ts = st.timesteps
        seeds = tuple(s for r in batch for s in [r.seed] * r.nb)
        nkey = (seeds, len(ts), self.noise_mode)
        if nkey != st.noise_key:
            init, steps, vae_noise = self._noise(nb, H // 8, W // 8, seeds, len(ts), eng.device)
            if vae_noise is not None:
                eng.set_vae_noise(vae_noise)
            eng.set_noise(init, steps)
            st.noise_key = nkey
        slot = 0
        for r in batch:
            for _ in range(r.nb):
                if st.slot_ctx.get(slot) != r.prompt_key:
                    eng.set_context(slot, r.embedding)
                    st.slot_ctx[slot] = r.prompt_key
                    self.stats["context_switches"] += 1
                slot += 1
        resized = (lead.in_h, lead.in_w) != (H, W)
        if resized and getattr(eng, "_resize_key", None) != (lead.in_w, lead.in_h, W, H, nb):
            eng.set_resize(lead.in_w, lead.in_h)
        pin = self._staging(st, lead.kind, nb, lead.in_h, lead.in_w, H, W)
        # gather the batch into the pinned staging planes (a single multi-frame request is copied the same way)
        row = 0
        for r in batch:
            planes = (r.data,) if r.kind == "rgb" else r.data
            for dst, src in zip(pin[0], planes):
                dst[row:row + r.nb].numpy()[...] = src.numpy() if isinstance(src, torch.Tensor) else src
            row += r.nb
        if lead.kind == "rgb":
            (eng.infer_rgb_resized if resized else eng.infer_rgb)(pin[0][0], pin[1][0])
        else:
            (eng.infer_yuv420_resized if resized else eng.infer_yuv420)(*pin[0], *pin[1])
        row = 0
        for r in batch:
            outs = tuple(t[row:row + r.nb].numpy().copy() for t in pin[1])
            r.result = outs[0] if r.kind == "rgb" else outs
            row += r.nb
        self.last_state = st                         # tests read the debug taps of the engine that ran the last launch
        self.stats["launches"] += 1
        self.stats["frames"] += nb
        self.stats["merged"] += nb - lead.nb

    def _set_schedule(self, eng, strength, steps):
        with self._plan_lock:
            if self._tuning_text:
                eng.tuning_load(self._tuning_text)
            before = eng.tuning_misses()
            ts = eng.set_schedule(strength, steps, 7.5)    # the UI's guidance_scale is dropped by the reference (F8)
            if eng.tuning_misses() != before or not self._tuning_text:
                self._tuning_text = eng.tuning_report()
        return ts

    @staticmethod
    def _staging(st, kind, nb, in_h, in_w, H, W):
        key = (kind, in_h, in_w)
        pin = st.pin.get(key)
        if pin is None:
            def buf(*shape):      # page-locked so the H2D / D2H copies inside vsd_infer_* are asynchronous DMA
                t = torch.empty(shape, dtype=torch.uint8)
                return t.pin_memory() if torch.cuda.is_available() else t
            if kind == "rgb":
                pin = ((buf(nb, in_h, in_w, 3),), (buf(nb, H, W, 3),))
            else:
                pin = ((buf(nb, in_h, in_w), buf(nb, in_h // 2, in_w // 2), buf(nb, in_h // 2, in_w // 2)),
                       (buf(nb, H, W), buf(nb, H // 2, W // 2), buf(nb, H // 2, W // 2)))
            st.pin[key] = pin
        return pin

    def close(self):
        for lane in self.lanes:
            lane.close()
