"""Multi-GPU plumbing of the EMAGE path.

Clips are independent (no cross-sample op in eval mode), so the path shards by a contiguous split of the clip
batch: one process per GPU, the checkpoint broadcast from rank 0 once at load (NCCL over NVLink on the GPU
box, gloo in the CPU tests), and NO collective inside the step.  Results stay on their rank unless the caller
asks for a gather (outside any timed region).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [start, end) of `n_items` owned by `rank`; sizes differ by at most one, earlier ranks
    take the remainder; empty ranges when world > n_items."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_checkpoint(*modules, src: int = 0) -> int:
    """Make every parameter and buffer of `modules` equal to rank `src`'s (one broadcast per tensor).
    Returns the number of bytes broadcast.  Drops any packed engine so it is rebuilt from the new weights."""
    nbytes = 0
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src)
            nbytes += t.numel() * t.element_size()
        for sub in m.modules():
            if hasattr(sub, "_engine"):
                sub._engine = None
    return nbytes


def generate_sharded(model, motion_vq, audio_all, rank=None, world=None, **kw):
    """Run pipeline.generate on this rank's contiguous slice of the clip batch.  audio_all: (clips, n) on any
    device (every rank passes the same tensor, or at least its own slice region).  Returns
    (start, end, latent, pred); latent/pred are None for an empty slice."""
    from .pipeline import generate
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    start, end = shard_range(audio_all.shape[0], rank, world)
    if end == start:
        return start, end, None, None
    dev = next(model.parameters()).device
    lat, pred = generate(model, motion_vq, audio_all[start:end].to(dev), **kw)
    return start, end, lat, pred


def gather_clips(local: torch.Tensor | None, n_total: int, dst: int = 0):
    """Gather per-rank (clips_r, ...) tensors into (n_total, ...) on rank `dst` (None elsewhere).  Variable
    shard sizes are handled by padding to the largest shard.  Not part of the timed step."""
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    biggest = max(e - s for s, e in sizes)
    shape = torch.zeros(8, dtype=torch.long)
    if local is not None:
        shape[0] = local.dim()
        shape[1:1 + local.dim()] = torch.tensor(local.shape)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape)
    ref = next(s for s in shapes if s[0] > 0)
    tail = tuple(int(v) for v in ref[2:1 + int(ref[0])])
    dev = local.device if local is not None else torch.device("cpu")
    dtype = local.dtype if local is not None else torch.float32
    pad = torch.zeros((biggest,) + tail, dtype=dtype, device=dev)
    if local is not None:
        pad[:local.shape[0]] = local
    bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:e - s] for b, (s, e) in zip(bufs, sizes)], dim=0)
