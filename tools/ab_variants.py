"""A/B harness for the experimental kernel variants (dsm_debug_set_variants).

For every variant mask given on the command line (default: 0 1 2 4 8 16 32 64 128 255) on one resident batch of B KITTI-shaped frames:
  * parity: labels, clustering state and surfels of every frame must be BYTE-IDENTICAL to mask 0 (the measured
    default path) -- the variants only change data movement;
  * timing: wall time of N graph-replayed steps (CUDA-event based profile off), then the per-kernel CUDA-event
    profile of 3 more steps.
Prints one JSON line per mask.  Needs a GPU:  python tools/ab_variants.py [B] [N] [mask ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from densesurfelmapping_b200 import capi, synth  # noqa: E402
from densesurfelmapping_b200.elements import SURFEL_DTYPE  # noqa: E402


def snapshot(ctx, B):
    local, new = ctx.batch_download()
    return dict(local=local.tobytes(), new=[n.tobytes() for n in new],
                labels=[ctx.labels(b).tobytes() for b in range(B)], seeds=[ctx.seeds(b).tobytes() for b in range(B)])


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    masks = [int(a, 0) for a in sys.argv[3:]] or [0, 1, 2, 4, 8, 16, 32, 64, 128, 255]
    if masks[0] != 0:
        masks = [0] + masks
    cam = synth.KITTI
    S = (cam.width // 8) * (cam.height // 8)
    prev, cur = bench.make_batch(cam, B, 0)
    ctx = capi.Context(cam, max_batch=B, max_local_surfels=B * S + 64)
    _, pools = ctx.fuse_batch([0] * B, np.stack([f[0] for f in prev]), np.stack([f[1] for f in prev]), np.stack([f[2] for f in prev]),
                              np.zeros(0, SURFEL_DTYPE), np.zeros(B + 1, np.int32))
    ofs = np.concatenate([[0], np.cumsum([len(p) for p in pools])]).astype(np.int32)
    ctx.batch_upload(np.zeros(B, np.int32), np.stack([f[0] for f in cur]), np.stack([f[1] for f in cur]), np.stack([f[2] for f in cur]),
                     np.concatenate(pools), ofs)
    names = capi.kernel_names()
    ref = None
    for m in masks:
        ctx.debug_set_variants(m)
        ctx.profile_enable(0)
        ctx.batch_restore_pool()
        ctx.batch_run()
        snap = snapshot(ctx, B)
        viol = ctx.invariant_violations()
        if ref is None:
            ref = snap
        same = {k: snap[k] == ref[k] for k in snap}
        for _ in range(3):  # warm-up (graph capture)
            ctx.batch_restore_pool()
            ctx.batch_run()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(N):
            ctx.batch_restore_pool()
            ctx.batch_run()
        ctx.sync()
        ms = (time.perf_counter() - t0) * 1e3 / N
        ctx.profile_enable((1 << capi.NUM_KERNELS) - 1)
        ctx.profile_reset()
        for _ in range(3):
            ctx.batch_restore_pool()
            ctx.batch_run()
        pms, pn = ctx.profile_read()
        per = {names[k]: round(float(pms[k]) * 1e3 / 3, 1) for k in range(capi.NUM_KERNELS) if pn[k]}
        print(json.dumps(dict(mask=m, frames=B, ms_per_step=round(ms, 4), frames_per_s=round(B / ms * 1e3, 1),
                              identical_to_default=same, invariant_violations=viol, us_per_step_by_kernel=per)), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
