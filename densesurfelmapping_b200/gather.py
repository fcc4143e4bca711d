"""Multi-GPU plumbing for batches of independent frames.

The path shards trivially: frame b of a batch goes to rank b mod G, every rank runs the whole
per-frame pipeline on its own frames with NO data-path collective.  The only exchange is one
gather of the per-rank surfel deltas (new surfels + updated local pools) onto rank 0 at the end
of a batch.  torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests); the
tensors are zero-copy views over the C-ABI library's device buffers.
"""
import numpy as np
import torch
import torch.distributed as dist

from .elements import SURFEL_DTYPE

SURFEL_WORDS = 11  # 44 bytes


def shard_frames(n_frames: int, rank: int, world: int):
    """Frame indices owned by `rank`: frame b -> rank b mod world (SURVEY.md §8e partitioning)."""
    return list(range(rank, n_frames, world))


def gather_deltas(new_surfels: torch.Tensor, new_counts: torch.Tensor, pool: torch.Tensor,
                  pool_count: int, dst: int = 0, bufs=None):
    """One gather of this rank's deltas onto `dst`.

    new_surfels: float32 [B*S*11] view of the library's new-surfel buffer ([B][S] records),
    new_counts:  int32 [B], pool: float32 [cap*11] view of the (updated in place) local pool,
    pool_count: valid surfels in `pool`.  Fixed-size gathers (equal length on every rank, as
    dist.gather requires) plus the counts needed to unpack them.  `bufs` (optional, dst only)
    preallocated receive lists so a benchmark does not allocate inside the timed region.
    Returns on dst: (list of new_surfels tensors, list of counts tensors, list of pool tensors,
    list of pool counts); elsewhere None.
    """
    world = dist.get_world_size()
    rank = dist.get_rank()
    meta = torch.cat([new_counts.to(torch.int32), torch.tensor([pool_count], dtype=torch.int32, device=new_counts.device)])
    if rank == dst:
        if bufs is None:
            bufs = ([torch.empty_like(new_surfels) for _ in range(world)],
                    [torch.empty_like(meta) for _ in range(world)],
                    [torch.empty_like(pool) for _ in range(world)])
        dist.gather(meta, bufs[1], dst=dst)
        dist.gather(new_surfels, bufs[0], dst=dst)
        dist.gather(pool, bufs[2], dst=dst)
        return bufs[0], [m[:-1] for m in bufs[1]], bufs[2], [m[-1] for m in bufs[1]]
    dist.gather(meta, None, dst=dst)
    dist.gather(new_surfels, None, dst=dst)
    dist.gather(pool, None, dst=dst)
    return None


def unpack_new(new_flat: torch.Tensor, counts: torch.Tensor, seeds_per_frame: int):
    """[B*S*11] float32 + [B] counts -> list of numpy SURFEL_DTYPE arrays (host side, rank 0)."""
    a = new_flat.detach().cpu().numpy().view(np.uint8).reshape(-1, seeds_per_frame, 44)
    c = counts.detach().cpu().numpy()
    return [np.ascontiguousarray(a[b, :int(c[b])]).view(SURFEL_DTYPE).reshape(-1) for b in range(len(c))]
