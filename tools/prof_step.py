"""Minimal driver for ncu: uploads one batch of KITTI-shaped frames and runs N resident steps (one launch sequence per
step; the production default runs two half-batches concurrently, which a profiler serialises anyway)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from densesurfelmapping_b200 import capi, synth
from densesurfelmapping_b200.elements import SURFEL_DTYPE

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cam = synth.KITTI
S = (cam.width // 8) * (cam.height // 8)
prev, cur = bench.make_batch(cam, B, 0)
ctx = capi.Context(cam, max_batch=B, max_local_surfels=B * S + 64)
ctx.set_concurrency(1)  # the whole batch as one launch sequence: 17 full-width launches per step, as in bench.py's per-kernel table
_, pools = ctx.fuse_batch([0] * B, np.stack([f[0] for f in prev]), np.stack([f[1] for f in prev]), np.stack([f[2] for f in prev]),
                          np.zeros(0, SURFEL_DTYPE), np.zeros(B + 1, np.int32))
ofs = np.concatenate([[0], np.cumsum([len(p) for p in pools])]).astype(np.int32)
ctx.batch_upload(np.zeros(B, np.int32), np.stack([f[0] for f in cur]), np.stack([f[1] for f in cur]), np.stack([f[2] for f in cur]),
                 np.concatenate(pools), ofs)
for _ in range(steps):
    ctx.batch_restore_pool()
    ctx.batch_run()
ctx.sync()
print("done", B, steps)
