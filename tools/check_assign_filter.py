"""CPU check of the error bound behind the filtered assign pass (k_assign2, csrc/dsm_tile.cu).

For seeded synthetic frames and the seeds the oracle holds after 0, 1 and 2 update passes, every (pixel, valid
candidate) cost is evaluated twice with numpy:
  exact   the reference's mixed float/double expression (fusion_functions.cpp:364-387), as calc_cost does;
  fast    the fp32 / FMA expression of the kernel's fast path (1/mean_depth as hi + lo).
Asserts |fast - exact| <= 2^-20 (fast + exact) + 1e-12 (the bound the kernel's test 2^-18 (m1 + m2) + 1e-10 relies on with
4x slack), that every pixel the filter calls certain has the reference's argmin, and prints the share of pixels
that fall back to the exact path.  Needs no GPU.   python tools/check_assign_filter.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402
from densesurfelmapping_b200 import synth  # noqa: E402

f32, f64 = np.float32, np.float64


def fma32(a, b, c):
    return (a.astype(f64) * b.astype(f64) + c.astype(f64)).astype(f32)


def check(cam, gray, depth, seeds, tag):
    H, W = gray.shape
    spw, sph = W // 8, H // 8
    yy, xx = np.mgrid[0:H, 0:W]
    pi = gray.astype(f32)
    with np.errstate(divide="ignore"):
        pinv = np.where(depth.astype(f64) > 0.01, (1.0 / depth.astype(f64)).astype(f32), f32(0))
    sx, sy, sI, smd = (seeds[k].astype(f32) for k in ("x", "y", "mean_intensity", "mean_depth"))
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / smd.astype(f64)
    hi = np.where(smd > 0, inv.astype(f32), f32(0))
    lo = np.where(smd > 0, (inv - hi.astype(f64)).astype(f32), f32(0))
    hi = np.where((smd > 0) & (smd < 2.0 ** -10), f32(np.inf), hi)
    bx, by = xx // 8, yy // 8
    exact_d, exact_n, fast_d, fast_n, valid, has, sid = [], [], [], [], [], [], []
    for ci in (-1, 0, 1):          # dx outer, dy inner (:413-414)
        for cj in (-1, 0, 1):
            cx, cy = bx + ci, by + cj
            ok = (np.abs(cx * 8 + 4 - xx) < 8) & (np.abs(cy * 8 + 4 - yy) < 8) & (cx >= 0) & (cx < spw) & (cy >= 0) & (cy < sph)
            s = np.clip(cy, 0, sph - 1) * spw + np.clip(cx, 0, spw - 1)
            ax = sx[s] - xx.astype(f32)
            ay = sy[s] - yy.astype(f32)
            # exact (reference): float dist, /16, float idf^2, / 100.0 and the add in double, rounded to float
            dist = (ax * ax + ay * ay).astype(f32)
            n = dist / f32(16)
            idf = sI[s] - pi
            nd = (n.astype(f64) + (idf * idf).astype(f32).astype(f64) / 100.0).astype(f32)
            h = (smd[s] > 0) & (pinv > 0)
            idd = (inv[s] - pinv.astype(f64)).astype(f32)
            with np.errstate(invalid="ignore", over="ignore"):
                wd = (nd.astype(f64) + (idd * idd).astype(f32).astype(f64) * 400.0).astype(f32)
            wd = np.where(h, wd, nd)
            # fast (kernel): fp32 with FMA
            nf = fma32(ax, ax, (ay * ay).astype(f32)) * f32(0.0625)
            cn = fma32((idf * idf).astype(f32), np.full_like(idf, f32(0.01)), nf)
            with np.errstate(invalid="ignore", over="ignore"):
                t = ((hi[s] - pinv).astype(f32) + lo[s]).astype(f32)
                cd = fma32((t * t).astype(f32), np.full_like(t, f32(400.0)), cn)
            exact_d.append(wd), exact_n.append(nd), fast_d.append(cd), fast_n.append(cn), valid.append(ok), has.append(h), sid.append(s)
    exact_d, exact_n, fast_d, fast_n, valid, has, sid = (np.stack(a) for a in (exact_d, exact_n, fast_d, fast_n, valid, has, sid))
    allhas = np.all(has | ~valid, axis=0)
    exact = np.where(allhas[None], exact_d, exact_n)
    fast = np.where(allhas[None], fast_d, fast_n)
    # reference argmin: first strict minimum below 1e6 in visiting order
    ex = np.where(valid, exact, f32(np.inf))
    ex = np.where(ex < f32(1e6), ex, f32(np.inf))
    ref_arg = np.argmin(ex, axis=0)          # first occurrence of the minimum
    ref_win = np.where(np.isfinite(ex.min(axis=0)), np.take_along_axis(sid, ref_arg[None], 0)[0], -1)
    # error bound on every valid candidate whose exact cost is finite and < 1e7
    m = valid & np.isfinite(exact) & (exact < 1e7) & np.isfinite(fast)
    err = np.abs(fast.astype(f64) - exact.astype(f64))
    bound = 2.0 ** -20 * (fast.astype(f64) + exact.astype(f64)) + 1e-12
    worst = float((err[m] / bound[m]).max())
    assert worst <= 1.0, f"{tag}: bound violated, err/bound = {worst}"
    # the kernel's filter
    fa = np.where(valid, fast, f32(1e30))
    fa = np.where(np.isnan(fa), f32(1e30), fa)
    order = np.sort(fa, axis=0)
    m1, m2 = order[0], order[1]
    with np.errstate(over="ignore", invalid="ignore"):
        certain = ((m2 - m1) > f32(2.0 ** -18) * (m2 + m1) + f32(1e-10)) & (m1 < f32(9e5))
    fast_win = np.take_along_axis(sid, np.argmin(fa, axis=0)[None], 0)[0]
    bad = certain & (fast_win != ref_win)
    assert not bad.any(), f"{tag}: {int(bad.sum())} certain pixels with the wrong winner, first {np.argwhere(bad)[:4].tolist()}"
    print(f"{tag}: max err/bound {worst:.3f}; exact-path pixels {100.0 * (1 - certain.mean()):.4f} %  ({int((~certain).sum())} of {certain.size})")


def main():
    cases = [(synth.KITTI, 0, False), (synth.KITTI, 3, False), (synth.VGA, 1, False), (synth.HD, 2, False)]
    for cam, fid, _ in cases:
        gray, depth = synth.make_frame(cam, fid, synth.pose_stream(fid))
        ro = pyoracle.Restatement(cam)
        for iters, upd in ((1, False), (1, True), (2, True), (3, True)):
            _, seeds = ro.debug_iters(gray, depth, iters, upd)
            check(cam, gray, depth, seeds, f"{cam.width}x{cam.height} frame {fid} seeds after {iters - (0 if upd else 1)} update(s)")
    # flat scene: exact ties everywhere -> many exact-path pixels, still never a wrong certain winner
    cam = synth.KITTI
    gray = np.full((cam.height, cam.width), 128, np.uint8)
    depth = np.full((cam.height, cam.width), 5.0, np.float32)
    ro = pyoracle.Restatement(cam)
    for iters, upd in ((1, False), (1, True), (2, True)):
        _, seeds = ro.debug_iters(gray, depth, iters, upd)
        check(cam, gray, depth, seeds, f"constant image, seeds after {iters - (0 if upd else 1)} update(s)")


if __name__ == "__main__":
    main()
