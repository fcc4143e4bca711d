"""CPU test of the N>1 plumbing: world_size-2 gloo run of the shard + gather-of-deltas logic
(densesurfelmapping_b200/gather.py) that bench.py uses over NCCL."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from densesurfelmapping_b200 import gather
from densesurfelmapping_b200.elements import SURFEL_DTYPE

S, B = 12, 3


def _fake_rank_data(rank):
    rng = np.random.RandomState(100 + rank)
    counts = rng.randint(0, S + 1, size=B).astype(np.int32)
    new = np.zeros((B, S), SURFEL_DTYPE)
    for b in range(B):
        for f in ("px", "py", "pz", "weight"):
            new[f][b, :counts[b]] = rng.rand(counts[b]).astype(np.float32)
        new["update_times"][b, :counts[b]] = 1
        new["last_update"][b, :counts[b]] = rank * 10 + b
    npool = 5 + rank
    pool = np.zeros(B * S, SURFEL_DTYPE)
    pool["px"][:npool] = rng.rand(npool).astype(np.float32)
    pool["update_times"][:npool] = rng.randint(0, 4, npool)
    return counts, new, pool, npool


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts, new, pool, npool = _fake_rank_data(rank)
    t_new = torch.from_numpy(new.view(np.uint8).reshape(-1).view(np.float32).copy())
    t_pool = torch.from_numpy(pool.view(np.uint8).reshape(-1).view(np.float32).copy())
    res = gather.gather_deltas(t_new, torch.from_numpy(counts), t_pool, npool, dst=0)
    if rank == 0:
        news, cnts, pools, pcnts = res
        ok = True
        for r in range(world):
            c, n, p, npl = _fake_rank_data(r)
            got = gather.unpack_new(news[r], cnts[r], S)
            ok &= all(got[b].tobytes() == n[b, :c[b]].tobytes() for b in range(B))
            ok &= int(pcnts[r]) == npl
            ok &= pools[r].numpy().view(np.uint8)[:npl * 44].tobytes() == p[:npl].tobytes()
        q.put(ok)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_frames():
    assert gather.shard_frames(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((gather.shard_frames(256, r, 8) for r in range(8)), [])) == list(range(256))


def test_gather_deltas_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) is True
