"""Multi-GPU plumbing: channels are independent, so they shard across ranks; the only data-path collective is the
broadcast of the shared wide-band IQ batch from the ingest rank (the reference runs one
demod thread per channel on every buffer, multifm/receiver.c:78-98, 195-244).  One process per GPU,
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(nr_channels: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous channel range [lo, hi) of `rank`: sizes differ by at most one, lower ranks take the extra ones."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, extra = divmod(nr_channels, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_iq(dist, buf, src: int = 0):
    """Every rank needs the whole IQ batch.  `buf` is a torch int16 tensor of 2*n interleaved I,Q values, filled on
    `src` and overwritten elsewhere.  Returns buf."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(buf.view(_torch().uint8), src=src)       # neither NCCL nor gloo moves int16: send the bytes
    return buf


def _torch():
    import torch
    return torch


def gather_counts(dist, torch, local_outputs: int, device="cpu") -> int:
    """Whole-job output count (sum over ranks) -- what bench.py divides by the max-over-ranks time."""
    t = torch.tensor([float(local_outputs)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
