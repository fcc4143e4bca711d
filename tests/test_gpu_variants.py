"""Experimental kernel variants (dsm_debug_set_variants): every variant must reproduce the default
path byte for byte.  The variants were written after the round's GPU budget was spent, so this module only runs
when DSM_TEST_VARIANTS=1 (first thing to run on hardware next round: `DSM_TEST_VARIANTS=1 pytest tests/test_gpu_variants.py`,
then `DSM_EXPERIMENTAL_VARIANTS=<mask> pytest tests -m gpu` for the whole suite under a variant)."""
import os

import numpy as np
import pytest

from densesurfelmapping_b200 import capi, synth
from densesurfelmapping_b200.elements import SURFEL_DTYPE

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("DSM_TEST_VARIANTS") != "1", reason="experimental variants: set DSM_TEST_VARIANTS=1")]

MASKS = [1, 2, 4, 8, 16, 32, 64, 128, 255]


def run_all_masks(cam, frames, pool):
    """frames: list of (gray, depth, pose); every frame is fused into the same pool, one context per camera."""
    ctx = capi.Context(cam, max_batch=1, max_local_surfels=max(len(pool), 1) + 64)
    ref = None
    for m in [0] + MASKS:
        ctx.debug_set_variants(m)
        outs = []
        for gray, depth, pose in frames:
            lo, no = ctx.fuse_frame(1, gray, depth, pose, pool.copy())
            outs.append((ctx.labels().tobytes(), ctx.seeds().tobytes(), lo.tobytes(), no.tobytes()))
            assert ctx.invariant_violations() == 0
        if ref is None:
            ref = outs
        for i, (a, b) in enumerate(zip(outs, ref)):
            for what, x, y in zip(("labels", "seeds", "local", "new"), a, b):
                assert x == y, f"variant mask {m}: {what} of frame {i} differ from the default path"
    ctx.close()


@pytest.mark.parametrize("cam", [synth.KITTI, synth.KITTI00, synth.VGA, synth.HD], ids=lambda c: f"{c.width}x{c.height}")
def test_variants_on_benchmark_shapes(cam):
    frames, pool = [], np.zeros(0, SURFEL_DTYPE)
    for t in range(2):
        pose = synth.pose_stream(t)
        gray, depth = synth.make_frame(cam, 900 + t, pose)
        frames.append((gray, depth, pose))
    ctx = capi.Context(cam, max_batch=1, max_local_surfels=64)
    _, pool = ctx.fuse_frame(0, *frames[0][:2], frames[0][2], pool)
    ctx.close()
    run_all_masks(cam, frames, pool)


@pytest.mark.parametrize("shape", [(24, 24), (36, 28), (64, 48), (68, 44), (132, 100), (260, 36)])
def test_variants_on_small_and_odd_shapes(shape):
    """strip / tile edge cases: one block per row, W % 8 remainders, windows clipped on every side"""
    W, H = shape
    cam = synth.Camera(W, H, 60.0, 60.0, (W - 1) / 2, (H - 1) / 2, 0.5, 30.0)
    yy, xx = np.mgrid[0:H, 0:W]
    frames = []
    for seed in range(6):
        rng = np.random.RandomState(seed)
        gray = rng.randint(0, 256, (H, W)).astype(np.uint8) if seed % 2 else ((xx * 3 + yy * 2) % 256).astype(np.uint8)
        depth = (rng.uniform(0.5, 20.0, (H, W)) if seed % 3 else 2.0 + 0.01 * xx + 0.02 * yy + rng.normal(0, 0.002, (H, W))).astype(np.float32)
        if seed >= 3:
            depth[rng.rand(H, W) < 0.3] = 0
        frames.append((gray, depth, synth.identity_pose()))
    ctx = capi.Context(cam, max_batch=1, max_local_surfels=64)
    _, pool = ctx.fuse_frame(0, *frames[1][:2], frames[1][2], np.zeros(0, SURFEL_DTYPE))
    ctx.close()
    run_all_masks(cam, frames, pool)


def test_variants_on_a_batch():
    """the batch path (frame index from the grid): 6 KITTI frames, all masks, byte-identical deltas"""
    import bench
    cam, B = synth.KITTI, 6
    S = (cam.width // 8) * (cam.height // 8)
    prev, cur = bench.make_batch(cam, B, 0)
    ctx = capi.Context(cam, max_batch=B, max_local_surfels=B * S + 64)
    _, pools = ctx.fuse_batch([0] * B, np.stack([f[0] for f in prev]), np.stack([f[1] for f in prev]), np.stack([f[2] for f in prev]),
                              np.zeros(0, SURFEL_DTYPE), np.zeros(B + 1, np.int32))
    ofs = np.concatenate([[0], np.cumsum([len(p) for p in pools])]).astype(np.int32)
    args = (np.zeros(B, np.int32), np.stack([f[0] for f in cur]), np.stack([f[1] for f in cur]), np.stack([f[2] for f in cur]),
            np.concatenate(pools), ofs)
    ref = None
    for m in [0] + MASKS:
        ctx.debug_set_variants(m)
        lo, no = ctx.fuse_batch(*args)
        got = (lo.tobytes(), [n.tobytes() for n in no], [ctx.labels(b).tobytes() for b in range(B)])
        ref = ref or got
        assert got == ref, f"variant mask {m} differs from the default path"
    ctx.close()
