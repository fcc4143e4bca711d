"""Multi-GPU plumbing for batches of independent frames.

The path shards trivially: frame b of a batch goes to rank b mod G, every rank runs the whole per-frame pipeline on
its own frames with NO data-path collective.  The only exchange is ONE gather of the per-rank surfel deltas (valid
new surfels + updated local pools) onto a root rank at the end of a batch.  On GPUs that gather lives in the C ABI
(`dsm_comm_init` / `dsm_gather_deltas`, csrc/dsm_comm.cu: a pack kernel that writes the valid records straight into
the root's memory over NVLink; NCCL send/recv where no peer mapping exists).  This module holds what the host side needs around it:

* `shard_frames`                   the partitioning;
* `pack_payload` / `unpack_payload` the wire format of one rank's payload, restated in numpy (the GPU tests demand
                                    that the device packer produces exactly these bytes);
* `gather_payloads`                the same variable-length gather over a torch.distributed process group (gloo in
                                    the CPU tests: counts first, then the payloads) -- what a host-memory caller does.
"""
import numpy as np

from .elements import SURFEL_DTYPE

MAGIC = 0x444D5344  # 'DSMD'


def shard_frames(n_frames: int, rank: int, world: int):
    """Frame indices owned by `rank`: frame b -> rank b mod world (SURVEY.md section 8e partitioning)."""
    return list(range(rank, n_frames, world))


def header_bytes(n_frames: int) -> int:
    return ((4 + n_frames + n_frames + 1) * 4 + 15) // 16 * 16


def pack_payload(new_per_frame, pool, pool_offsets) -> np.ndarray:
    """One rank's delta payload as a uint8 array (layout: include/dsm.h, "multi-GPU").
    new_per_frame: list of SURFEL_DTYPE arrays (valid new surfels of every frame, seed-index order);
    pool: SURFEL_DTYPE array (updated local surfels, batch order); pool_offsets: int32 [n_frames + 1]."""
    nb = len(new_per_frame)
    n_new = [len(a) for a in new_per_frame]
    hdr = np.zeros(header_bytes(nb) // 4, np.int32)
    hdr[0], hdr[1], hdr[2], hdr[3] = MAGIC, nb, sum(n_new), len(pool)
    hdr[4:4 + nb] = n_new
    hdr[4 + nb:4 + nb + nb + 1] = np.asarray(pool_offsets, np.int32)
    parts = [hdr.view(np.uint8)]
    parts += [np.ascontiguousarray(a, dtype=SURFEL_DTYPE).view(np.uint8).reshape(-1) for a in new_per_frame if len(a)]
    if len(pool):
        parts.append(np.ascontiguousarray(pool, dtype=SURFEL_DTYPE).view(np.uint8).reshape(-1))
    return np.concatenate(parts)


def unpack_payload(buf: np.ndarray):
    """Inverse of pack_payload: (list of new-surfel arrays per frame, pool array, pool offsets)."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    hdr = buf[:16].view(np.int32)
    if hdr[0] != MAGIC:
        raise ValueError("not a DSMD payload")
    nb, n_new_total, n_pool = int(hdr[1]), int(hdr[2]), int(hdr[3])
    hb = header_bytes(nb)
    full = buf[:hb].view(np.int32)
    n_new = full[4:4 + nb]
    ofs = full[4 + nb:4 + nb + nb + 1].copy()
    body = buf[hb:hb + (n_new_total + n_pool) * 44].view(SURFEL_DTYPE)
    news, at = [], 0
    for c in n_new:
        news.append(body[at:at + int(c)].copy())
        at += int(c)
    return news, body[n_new_total:n_new_total + n_pool].copy(), ofs


def gather_payloads(payload: np.ndarray, dst: int = 0):
    """Variable-length gather of one uint8 payload per rank onto `dst` over the default torch.distributed group:
    the byte counts first (all_gather), then the payloads (point-to-point to dst).  Returns the list of payloads on dst,
    None elsewhere.  Mirrors dsm_gather_deltas step by step (counts, then exact-size messages)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([payload.size], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, mine)
    t = torch.from_numpy(np.ascontiguousarray(payload))
    if rank == dst:
        out = []
        for r in range(world):
            if r == rank:
                out.append(payload.copy())
            else:
                b = torch.empty(int(counts[r].item()), dtype=torch.uint8)
                dist.recv(b, src=r)
                out.append(b.numpy())
        return out
    dist.send(t, dst=dst)
    return None
