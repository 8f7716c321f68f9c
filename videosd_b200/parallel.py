"""Multi-GPU plumbing for the stream/frame-parallel deployment: the hot path has no exchange step,
so ranks only agree on who owns which stream and on the slowest rank's time. torch.distributed is used for the
barriers and the max-reduction only (NCCL on GPUs, gloo in the CPU tests); no collective touches frame data.

The reference does the same job with one Ray actor per GPU and first-idle-GPU dispatch (server.py:132-137,
:320-321); here sessions are pinned to a GPU and co-located sessions are batched.
"""
from dataclasses import dataclass, field
from typing import Dict, List


def shard_streams(num_streams, world_size, rank):
    """Round-robin pinning of stream ids to ranks; every stream has exactly one owner."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return [s for s in range(num_streams) if s % world_size == rank]


def max_over_ranks(value, device=None):
    """Max of a python float across the default process group (identity when not initialised)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def aggregate_fps(frames_per_rank, seconds, device=None):
    """Whole-job frames/s: all ranks' frames over the slowest rank's time."""
    import torch
    import torch.distributed as dist

    slowest = max_over_ranks(seconds, device)
    if dist.is_available() and dist.is_initialized():
        t = torch.tensor([float(frames_per_rank)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total = float(t[0])
    else:
        total = float(frames_per_rank)
    return total / slowest


@dataclass
class SessionRouter:
    """Pins WebRTC sessions to GPUs (least-loaded first) and forms per-GPU frame batches of <= max_batch."""
    num_gpus: int
    max_batch: int = 4
    _owner: Dict[str, int] = field(default_factory=dict)
    _load: List[int] = field(default_factory=list)

    def __post_init__(self):
        self._load = [0] * self.num_gpus

    def assign(self, session_id):
        if session_id in self._owner:
            return self._owner[session_id]
        gpu = min(range(self.num_gpus), key=lambda g: (self._load[g], g))
        self._owner[session_id] = gpu
        self._load[gpu] += 1
        return gpu

    def release(self, session_id):
        gpu = self._owner.pop(session_id, None)
        if gpu is not None:
            self._load[gpu] -= 1

    def batches(self, pending_session_ids):
        """pending: sessions that have a fresh frame. -> {gpu: [[session ids of one batch], ...]}"""
        per_gpu: Dict[int, List[str]] = {}
        for s in pending_session_ids:
            per_gpu.setdefault(self.assign(s), []).append(s)
        return {g: [ids[i:i + self.max_batch] for i in range(0, len(ids), self.max_batch)] for g, ids in per_gpu.items()}
