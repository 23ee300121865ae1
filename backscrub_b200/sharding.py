"""Stream sharding across GPUs.

Frames of different streams are independent; frames inside a stream only share the 3-frame IIR
state.  So the unit of distribution is the *stream*: stream s runs on rank s mod world, each
rank owns its contexts / weights / background / CUDA graphs, and there is no data-path
collective.  The only communication is the end-of-run reduction of (frames done, elapsed).
"""
from __future__ import annotations


def streams_for_rank(n_streams: int, rank: int, world: int) -> list[int]:
    """Global stream ids served by `rank` (round-robin)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return [s for s in range(n_streams) if s % world == rank]


def reduce_throughput(frames_local: int, seconds_local: float, dist=None, device=None) -> tuple[int, float, float]:
    """Whole-job (frames, seconds, frames/s): SUM of frames, MAX of elapsed over ranks."""
    frames, secs = frames_local, seconds_local
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        f = torch.tensor([float(frames_local)], dtype=torch.float64, device=device)
        t = torch.tensor([float(seconds_local)], dtype=torch.float64, device=device)
        dist.all_reduce(f, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        frames, secs = int(round(f.item())), float(t.item())
    return frames, secs, (frames / secs if secs > 0 else 0.0)
