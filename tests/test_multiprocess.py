"""CPU tests of the N>1 plumbing: the wire format of the surfel-delta gather (restated in numpy, gather.py; the device
packer of csrc/dsm_comm.cu must produce exactly these bytes, tests/test_gpu_comm.py) and a world_size-2 gloo run of
the shard + variable-length gather logic."""
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from densesurfelmapping_b200 import gather
from densesurfelmapping_b200.elements import SURFEL_DTYPE

S, B = 12, 3


def _fake_rank_data(rank):
    rng = np.random.RandomState(100 + rank)
    counts = rng.randint(0, S + 1, size=B).astype(np.int32)
    news = []
    for b in range(B):
        a = np.zeros(counts[b], SURFEL_DTYPE)
        for f in ("px", "py", "pz", "weight"):
            a[f] = rng.rand(counts[b]).astype(np.float32)
        a["update_times"] = 1
        a["last_update"] = rank * 10 + b
        news.append(a)
    per = rng.randint(0, 6, size=B)
    ofs = np.concatenate([[0], np.cumsum(per)]).astype(np.int32)
    pool = np.zeros(int(ofs[-1]), SURFEL_DTYPE)
    pool["px"] = rng.rand(len(pool)).astype(np.float32)
    pool["update_times"] = rng.randint(0, 4, len(pool))
    return news, pool, ofs


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    news, pool, ofs = _fake_rank_data(rank)
    res = gather.gather_payloads(gather.pack_payload(news, pool, ofs), dst=0)
    if rank == 0:
        ok = len(res) == world
        for r in range(world):
            n, p, o = _fake_rank_data(r)
            gn, gp, go = gather.unpack_payload(res[r])
            ok &= len(gn) == B and all(gn[b].tobytes() == n[b].tobytes() for b in range(B))
            ok &= gp.tobytes() == p.tobytes() and (go == o).all()
        q.put(bool(ok))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_frames():
    assert gather.shard_frames(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((gather.shard_frames(256, r, 8) for r in range(8)), [])) == list(range(256))


def test_payload_round_trip_and_layout():
    news, pool, ofs = _fake_rank_data(3)
    buf = gather.pack_payload(news, pool, ofs)
    hdr = buf[:16].view(np.int32)
    assert hdr[0] == 0x444D5344 and hdr[1] == B and hdr[2] == sum(len(a) for a in news) and hdr[3] == len(pool)
    assert buf.size == gather.header_bytes(B) + (hdr[2] + hdr[3]) * 44 and gather.header_bytes(B) % 16 == 0
    gn, gp, go = gather.unpack_payload(buf)
    assert all(a.tobytes() == b.tobytes() for a, b in zip(gn, news)) and gp.tobytes() == pool.tobytes() and (go == ofs).all()
    empty = gather.pack_payload([np.zeros(0, SURFEL_DTYPE)] * 2, np.zeros(0, SURFEL_DTYPE), [0, 0, 0])
    gn, gp, go = gather.unpack_payload(empty)
    assert [len(a) for a in gn] == [0, 0] and len(gp) == 0


def test_gather_payloads_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) is True
