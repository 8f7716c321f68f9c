"""Python host wrapper over the frame-level C ABI (include/videosd.h, "engine" section).

PyTorch/numpy only own host buffers here; every per-frame computation happens inside libvideosd.so. One Engine per
GPU, calls serialised by the caller (like one Ray actor per GPU in the reference, diffusert/videopipeline.py:11).
"""
import ctypes
import os

import numpy as np
import torch

from . import scheduler as _sched
from ._lib import VsdError, check, lib

c_int = ctypes.c_int

# Committed GEMM tuning tables (videosd_b200/tuning/<H>x<W>x<B>_n<frames in flight>.txt, written by
# tools/make_tuning_tables.py on a B200): loaded by default so every process runs the same kernel configurations
# (same (frame, options) -> same output across processes and ranks) and configure() does not stall on device timing.
# VSD_TUNING_DIR overrides the directory; VSD_TUNING_TABLES=0 disables loading (used when the tables are regenerated).
TUNING_DIR = os.environ.get("VSD_TUNING_DIR") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "tuning")


def tuning_table_path(batch, height, width, frames_in_flight):
    return os.path.join(TUNING_DIR, f"{height}x{width}x{batch}_n{max(1, int(frames_in_flight))}.txt")


def resolve_tuning_table(batch, height, width, frames_in_flight):
    """Path of the committed table to load: the one for exactly this many frames in flight, else the nearest smaller count
    (the same file in every process, so the kernels still do not depend on device timing; live tuning of ~80 shapes would
    stall the stream for 20 s). None when the (size, batch) has no table at all."""
    for n in range(max(1, int(frames_in_flight)), 0, -1):
        cand = tuning_table_path(batch, height, width, n)
        if os.path.exists(cand):
            return cand
    return None


def _fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _u8ptr(t):
    if isinstance(t, torch.Tensor):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


class Engine:
    def __init__(self, device=0, parent=None):
        """parent: another Engine on the same GPU whose (already loaded) weights this lane shares."""
        L = lib()
        L.vsd_create.restype = ctypes.c_void_p
        L.vsd_create_lane.restype = ctypes.c_void_p
        L.vsd_stream.restype = ctypes.c_void_p
        L.vsd_launches_per_frame.restype = ctypes.c_long
        L.vsd_arena_peak_bytes.restype = ctypes.c_long
        self._L = L
        self.device = device
        self._parent = parent   # keeps the weight owner alive
        self._ctx = L.vsd_create_lane(parent._ctx) if parent is not None else L.vsd_create(c_int(device))
        if not self._ctx:
            raise VsdError("vsd_create failed: " + (L.vsd_last_error() or b"?").decode())
        self._ctx = ctypes.c_void_p(self._ctx)
        if parent is not None:
            parent._sched_key = None   # vsd_create_lane drops a plan the parent built while it was alone on the GPU
        self._schedule = _sched.LCMSchedule()
        self.batch = self.height = self.width = None
        self.timesteps = None
        self._sched_key = None
        self._tune_for = parent._tune_for if parent is not None else 1
        self._tables_loaded = set()

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.vsd_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ---- weights ---------------------------------------------------------------------------------------
    def load_state_dict(self, prefix, state_dict):
        """prefix: 'unet' or 'vae'; state_dict: name -> torch tensor (any float dtype) in PyTorch layout."""
        for name, t in state_dict.items():
            a = t.detach().to(torch.float32).contiguous().cpu().numpy()
            shape = (ctypes.c_int64 * a.ndim)(*a.shape)
            check(self._L.vsd_load_weight(self._ctx, f"{prefix}.{name}".encode(), _fptr(a), shape, c_int(a.ndim)),
                  f"vsd_load_weight({prefix}.{name})")

    # ---- configuration ---------------------------------------------------------------------------------
    def configure(self, batch, height, width):
        check(self._L.vsd_configure(self._ctx, c_int(batch), c_int(height), c_int(width)), "vsd_configure")
        self.batch, self.height, self.width = batch, height, width
        self._sched_key = None
        self._resize_key = None      # vsd_configure drops the resize tables / staging buffers
        self.load_tuning_table()

    def load_tuning_table(self):
        """Loads the committed table for the current (size, batch, frames in flight); shapes it does not cover are
        timed on the device when the plan is built (tuning_misses() counts them). Returns the entries loaded."""
        if self._tune_for <= 0 or os.environ.get("VSD_TUNING_TABLES") == "0" or self.batch is None:
            return 0
        path = resolve_tuning_table(self.batch, self.height, self.width, self._tune_for)
        if path is None or path in self._tables_loaded:
            return 0
        self._tables_loaded.add(path)
        with open(path) as f:
            return self.tuning_load(f.read())

    def tuning_misses(self):
        self._L.vsd_tuning_misses.restype = ctypes.c_long
        return int(self._L.vsd_tuning_misses(self._ctx))

    def set_schedule(self, strength, steps, guidance_scale=7.5):
        ts = self._schedule.timesteps(strength, steps)
        if not ts:
            raise ValueError(f"strength {strength} yields an empty LCM timestep table")
        key = (tuple(ts), float(guidance_scale))
        if key == self._sched_key:
            return ts
        sc = np.ascontiguousarray(self._schedule.step_scalars(ts))
        a, b = self._schedule.add_noise_coeffs(ts[0])
        w = np.ascontiguousarray(_sched.guidance_embedding(guidance_scale))
        tarr = (ctypes.c_int * len(ts))(*ts)
        check(self._L.vsd_set_schedule(self._ctx, c_int(len(ts)), tarr, _fptr(sc), ctypes.c_float(a), ctypes.c_float(b),
                                       _fptr(w), c_int(1 if len(ts) > 1 else 0)), "vsd_set_schedule")
        self.timesteps = ts
        self._sched_key = key
        return ts

    def set_controlnet(self, enabled, scale=1.0):
        """Enable the canny ControlNet branch with conditioning scale `scale` (guess mode, as the reference calls it)."""
        sc = np.ascontiguousarray((torch.logspace(-1, 0, 13) * float(scale)).numpy().astype(np.float32))
        check(self._L.vsd_set_controlnet(self._ctx, c_int(1 if enabled else 0), _fptr(sc)), "vsd_set_controlnet")
        if bool(enabled) != getattr(self, "_cn_enabled", False):
            self._cn_enabled = bool(enabled)
            self._sched_key = None      # the launch plan must be rebuilt

    def set_context(self, slot, context):
        """context: (77, 768) float tensor/array (CLIP last_hidden_state for the prompt)."""
        a = np.ascontiguousarray(torch.as_tensor(context).detach().to(torch.float32).cpu().numpy())
        if a.shape != (77, 768):
            raise ValueError(f"context must be (77, 768), got {a.shape}")
        check(self._L.vsd_set_context(self._ctx, c_int(slot), _fptr(a)), "vsd_set_context")

    def set_vae(self, kind):
        """'taesd' (AutoencoderTiny, the default: what the reference loads) or 'kl' (AutoencoderKL, weights under 'vae_kl')."""
        k = {"taesd": 0, "tiny": 0, 0: 0, "kl": 1, 1: 1}[kind]
        check(self._L.vsd_set_vae(self._ctx, c_int(k)), "vsd_set_vae")
        self._sched_key = None

    def set_vae_noise(self, noise_nchw):
        """Noise of AutoencoderKL's `latent_dist.sample()`: (B, 4, h/8, w/8)."""
        a = np.ascontiguousarray(torch.as_tensor(noise_nchw).detach().to(torch.float32).permute(0, 2, 3, 1).contiguous().numpy())
        if a.shape != (self.batch, self.height // 8, self.width // 8, 4):
            raise ValueError(f"vae noise must be (B, 4, h/8, w/8), got NHWC {a.shape}")
        check(self._L.vsd_set_vae_noise(self._ctx, _fptr(a)), "vsd_set_vae_noise")

    def encode_prompt(self, token_ids):
        """token_ids: 77 ints (CLIP tokenizer output padded to max_length) -> (77, 768) fp32 tensor, the text encoder's
        last_hidden_state (lcm_controlnet.py:175-179). Needs the 'text_encoder' weights."""
        ids = np.ascontiguousarray(np.asarray(token_ids, dtype=np.int32).reshape(-1))
        if ids.shape != (77,):
            raise ValueError(f"expected 77 token ids, got {ids.shape}")
        out = np.empty((77, 768), dtype=np.float32)
        check(self._L.vsd_encode_prompt(self._ctx, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _fptr(out)),
              "vsd_encode_prompt")
        return torch.from_numpy(out)

    def set_noise(self, init_noise_nchw, step_noise_nchw):
        """init: (B,4,h,w); steps: list of (B,4,h,w) (empty for single-step)."""
        init = np.ascontiguousarray(init_noise_nchw.permute(0, 2, 3, 1).contiguous().float().numpy())
        if len(step_noise_nchw):
            st = torch.stack(list(step_noise_nchw), 0).permute(0, 1, 3, 4, 2).contiguous().float().numpy()
            st = np.ascontiguousarray(st)
            check(self._L.vsd_set_noise(self._ctx, _fptr(init), _fptr(st)), "vsd_set_noise")
        else:
            check(self._L.vsd_set_noise(self._ctx, _fptr(init), None), "vsd_set_noise")

    def set_reference_noise(self):
        init, steps = _sched.reference_cpu_noise(self.batch, self.height // 8, self.width // 8, len(self.timesteps))
        self.set_noise(init, steps)

    # ---- frames ----------------------------------------------------------------------------------------
    def infer_yuv420(self, y, u, v, out_y, out_u, out_v):
        """u8 host buffers (numpy arrays or CPU torch tensors, ideally pinned); synchronous."""
        check(self._L.vsd_infer_yuv420(self._ctx, _u8ptr(y), _u8ptr(u), _u8ptr(v), _u8ptr(out_y), _u8ptr(out_u),
                                       _u8ptr(out_v)), "vsd_infer_yuv420")

    def infer_rgb(self, rgb_in, rgb_out):
        check(self._L.vsd_infer_rgb(self._ctx, _u8ptr(rgb_in), _u8ptr(rgb_out)), "vsd_infer_rgb")

    def set_resize(self, in_w, in_h):
        """Prepare the GPU center-crop + Lanczos resize from in_w x in_h frames to the configured working size."""
        from . import resample
        plan = resample.resize_plan(in_w, in_h, self.width, self.height)
        x0, y0, cw, ch = plan["crop"]
        (hb, hk, hks), (vb, vk, vks) = plan["h"], plan["v"]
        ip = lambda a: np.ascontiguousarray(a, dtype=np.int32).ctypes.data_as(ctypes.POINTER(ctypes.c_int))  # noqa: E731
        check(self._L.vsd_set_resize(self._ctx, c_int(in_w), c_int(in_h), c_int(x0), c_int(y0), c_int(cw), c_int(ch), ip(hb),
                                     ip(hk), c_int(hks), ip(vb), ip(vk), c_int(vks)), "vsd_set_resize")
        self._resize_key = (in_w, in_h, self.width, self.height, self.batch)

    def infer_rgb_resized(self, rgb_src, rgb_out):
        """rgb_src: u8 (batch, in_h, in_w, 3) host; rgb_out: u8 (batch, height, width, 3) host."""
        check(self._L.vsd_infer_rgb_resized(self._ctx, _u8ptr(rgb_src), _u8ptr(rgb_out)), "vsd_infer_rgb_resized")

    def infer_yuv420_resized(self, y, u, v, out_y, out_u, out_v):
        """y/u/v: u8 planes at the set_resize() source geometry; out_*: u8 planes at the working size."""
        check(self._L.vsd_infer_yuv420_resized(self._ctx, _u8ptr(y), _u8ptr(u), _u8ptr(v), _u8ptr(out_y), _u8ptr(out_u),
                                               _u8ptr(out_v)), "vsd_infer_yuv420_resized")

    def debug_read_rgb_in(self):
        a = np.empty((self.batch, self.height, self.width, 3), dtype=np.uint8)
        check(self._L.vsd_debug_read_rgb_in(self._ctx, _u8ptr(a)), "vsd_debug_read_rgb_in")
        return a

    def upload_yuv420(self, y, u, v):
        check(self._L.vsd_upload_yuv420(self._ctx, _u8ptr(y), _u8ptr(u), _u8ptr(v)), "vsd_upload_yuv420")

    def run_yuv420(self):
        check(self._L.vsd_run_yuv420(self._ctx), "vsd_run_yuv420")

    def download_yuv420(self, y, u, v):
        check(self._L.vsd_download_yuv420(self._ctx, _u8ptr(y), _u8ptr(u), _u8ptr(v)), "vsd_download_yuv420")

    def sync(self):
        check(self._L.vsd_sync(self._ctx), "vsd_sync")

    @property
    def stream(self):
        return self._L.vsd_stream(self._ctx)

    def launches_per_frame(self, yuv=True):
        return int(self._L.vsd_launches_per_frame(self._ctx, c_int(1 if yuv else 0)))

    def set_autotune(self, frames_in_flight):
        """0 / False: shape heuristics only. n >= 1: time candidate GEMM configurations on the device and pick for n frames in
        flight on this GPU (1 = lowest latency; more = configurations that leave SMs to the other frames)."""
        check(self._L.vsd_set_autotune(self._ctx, c_int(int(frames_in_flight))), "vsd_set_autotune")
        self._tune_for = int(frames_in_flight)
        self.load_tuning_table()

    def tuning_report(self):
        buf = ctypes.create_string_buffer(1 << 18)
        self._L.vsd_tuning_report(self._ctx, buf, ctypes.c_long(len(buf)))
        return buf.value.decode()

    def tuning_load(self, text):
        return int(self._L.vsd_tuning_load(self._ctx, text.encode()))

    def arena_peak_bytes(self):
        return int(self._L.vsd_arena_peak_bytes(self._ctx))

    # ---- debug taps (parity tests) ------------------------------------------------------------------------
    def debug_read(self, what, index=0, channels=4, spatial="latent"):
        h, w = (self.height // 8, self.width // 8) if spatial == "latent" else (self.height, self.width)
        a = np.empty((self.batch, h, w, channels), dtype=np.float32)
        check(self._L.vsd_debug_read(self._ctx, what.encode(), c_int(index), _fptr(a), ctypes.c_long(a.size)),
              f"vsd_debug_read({what})")
        return torch.from_numpy(a).permute(0, 3, 1, 2).contiguous()  # NCHW like the oracle

    def debug_unet(self, latents_nchw, step):
        a = np.ascontiguousarray(latents_nchw.permute(0, 2, 3, 1).contiguous().float().numpy())
        o = np.empty_like(a)
        check(self._L.vsd_debug_unet(self._ctx, _fptr(a), c_int(step), _fptr(o)), "vsd_debug_unet")
        return torch.from_numpy(o).permute(0, 3, 1, 2).contiguous()

    def debug_profile_sections(self, depth=2, reps=10):
        buf = ctypes.create_string_buffer(1 << 20)
        check(self._L.vsd_debug_profile_sections(self._ctx, c_int(depth), c_int(reps), buf, ctypes.c_long(len(buf))),
              "vsd_debug_profile_sections")
        return [(t, int(n), float(us)) for t, n, us in (ln.split() for ln in buf.value.decode().splitlines() if ln.strip())]

    def debug_run_eager(self, yuv=True):
        check(self._L.vsd_debug_run_eager(self._ctx, c_int(1 if yuv else 0)), "vsd_debug_run_eager")


class LanePool:
    """N lanes on one GPU sharing one copy of the weights; frames are handed to lanes round-robin so up to N frames
    are in flight (CUDA kernels of independent frames overlap and fill the SMs that a single batch-1 frame leaves
    idle). Configuration calls are broadcast to every lane."""

    def __init__(self, device=0, lanes=2, tune_for=None):
        self.lanes = [Engine(device)]
        self._n = lanes
        self._next = 0
        # the GEMM autotuner picks configurations for `tune_for` frames in flight (default: the number of lanes)
        self.lanes[0].set_autotune(lanes if tune_for is None else tune_for)

    def load_state_dict(self, prefix, sd):
        self.lanes[0].load_state_dict(prefix, sd)

    def finalize(self):
        """Call after all weights are loaded: creates the additional lanes."""
        while len(self.lanes) < self._n:
            self.lanes.append(Engine(self.lanes[0].device, parent=self.lanes[0]))

    def configure(self, batch, height, width):
        self.finalize()
        for e in self.lanes:
            e.configure(batch, height, width)

    def set_schedule(self, strength, steps, guidance_scale=7.5):
        ts = None
        for i, e in enumerate(self.lanes):
            if i:
                e.tuning_load(self.lanes[0].tuning_report())   # tune once
            ts = e.set_schedule(strength, steps, guidance_scale)
        return ts

    def set_controlnet(self, enabled, scale=1.0):
        for e in self.lanes:
            e.set_controlnet(enabled, scale)

    def set_vae(self, kind):
        for e in self.lanes:
            e.set_vae(kind)

    def set_vae_noise(self, noise_nchw):
        for e in self.lanes:
            e.set_vae_noise(noise_nchw)

    def set_context(self, slot, context):
        for e in self.lanes:
            e.set_context(slot, context)

    def set_reference_noise(self):
        for e in self.lanes:
            e.set_reference_noise()

    def set_noise(self, init, steps):
        for e in self.lanes:
            e.set_noise(init, steps)

    def next_lane(self):
        e = self.lanes[self._next]
        self._next = (self._next + 1) % len(self.lanes)
        return e

    def sync(self):
        for e in self.lanes:
            e.sync()

    def close(self):
        for e in reversed(self.lanes):
            e.close()
