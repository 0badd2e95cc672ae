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


def _comm_device() -> torch.device:
    """Device collectives must use: the current CUDA device under NCCL, the CPU under gloo."""
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_checkpoint(*modules, src: int = 0) -> int:
    """Make every parameter and buffer of `modules` equal to rank `src`'s: the tensors are packed into ONE flat
    arena per dtype (a single NCCL broadcast of the weight arena over NVLink at load, 0.6 GB in
    fp32), broadcast, and copied back.  Returns the number of bytes broadcast.  Drops any packed engine so it is rebuilt
    from the new weights."""
    tensors, seen = [], set()
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            if id(t) not in seen:
                seen.add(id(t))
                tensors.append(t.data)
    nbytes = 0
    dev = _comm_device()
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, group in sorted(by_dtype.items(), key=lambda kv: str(kv[0])):     # same order on every rank
        arena = torch.cat([t.reshape(-1).to(dev) for t in group]) if group else None
        dist.broadcast(arena, src=src)
        off = 0
        for t in group:
            n = t.numel()
            t.copy_(arena[off:off + n].view(t.shape))
            off += n
        nbytes += arena.numel() * arena.element_size()
    for m in modules:
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
    dev = _comm_device()                       # NCCL: every buffer on this rank's GPU, ranks with an empty shard included
    dtypes = [torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int64, torch.int32, torch.uint8, torch.bool]
    meta = torch.zeros(9, dtype=torch.long)
    if local is not None:
        meta[0] = local.dim()
        meta[1:1 + local.dim()] = torch.tensor(local.shape)
        meta[8] = dtypes.index(local.dtype)
    meta = meta.to(dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    ref = next(m for m in metas if int(m[0]) > 0).cpu()
    tail = tuple(int(v) for v in ref[2:1 + int(ref[0])])
    dtype = dtypes[int(ref[8])]                # ranks with nothing to send learn shape and dtype from the others
    pad = torch.zeros((biggest,) + tail, dtype=dtype, device=dev)
    if local is not None:
        pad[:local.shape[0]] = local.to(dev)
    bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:e - s] for b, (s, e) in zip(bufs, sizes)], dim=0)
