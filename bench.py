#!/usr/bin/env python
"""bench.py — frames/s of the DenseSurfelMapping per-frame hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (superpixels -> normals/plane fit -> fuse -> initialise,
kernels K0..K6) over one batch of FRAMES_PER_GPU independent synthetic KITTI-shaped 1226x370
depth+gray frames per GPU, each frame with its own pose and its own ~6 k-surfel local pool
(BASELINE configs[2] at N=1, configs[3] at N=8; weak scaling: per-GPU work is fixed).

  value  = frames/s, whole job, inputs already resident in HBM (kernels + pool restore only)
  e2e    = frames/s through the reference-facing C ABI call dsm_fuse_batch with pinned HOST buffers
           (H2D of gray/depth/poses/pool + kernels + D2H of pool and new surfels inside the timed region)
  roofline = dominant kernel: algorithmic bytes per launch / its CUDA-event duration inside the timed region
  cpu_baseline = the reference's own fusion_functions.cpp (oracle/_ref, 10 std::threads per frame,
           one instance per 10 host threads) on the same frames, rank 0 at N=1 only

`--impl reference` times only that CPU arm.  Multi-GPU: launched by torchrun, one rank per GPU; at
the end of every step the per-GPU surfel deltas (new surfels + updated pool) are gathered on rank 0
over NCCL.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_GPU = int(os.environ.get("DSM_BENCH_FRAMES", "32"))  # 32 is the graded configuration
METRIC = "frames/sec (KITTI 1226x370 depth+gray)"


def make_batch(cam, nframes, rank):
    """nframes independent frames (own pose each) + their predecessor views (pose one step back)."""
    from densesurfelmapping_b200 import synth
    cache = f"/tmp/dsm_bench_batch_{cam.width}x{cam.height}_{nframes}_{rank}.npz"
    if os.path.exists(cache):  # both arms of a round use the same frames; rendering them is untimed setup
        try:
            z = np.load(cache)
            return ([(z["g0"][i], z["d0"][i], z["p0"][i]) for i in range(nframes)],
                    [(z["g1"][i], z["d1"][i], z["p1"][i]) for i in range(nframes)])
        except Exception:
            pass
    prev, cur = [], []
    for i in range(nframes):
        fid = rank * 1000 + i
        t = (fid * 7) % 50 + 1
        p0, p1 = synth.pose_stream(t - 1), synth.pose_stream(t)
        g0, d0 = synth.make_frame(cam, 2 * fid, p0)
        g1, d1 = synth.make_frame(cam, 2 * fid + 1, p1)
        prev.append((g0, d0, p0))
        cur.append((g1, d1, p1))
    try:
        np.savez(cache, g0=np.stack([f[0] for f in prev]), d0=np.stack([f[1] for f in prev]), p0=np.stack([f[2] for f in prev]),
                 g1=np.stack([f[0] for f in cur]), d1=np.stack([f[1] for f in cur]), p1=np.stack([f[2] for f in cur]))
    except Exception:
        pass
    return prev, cur


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons.  nvidia-smi takes a few hundred ms to start, so it
    is launched before the warm-up and its samples are filtered to the timed region by timestamp."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self, t_begin, t_end):
        """t_begin/t_end: time.time() around the timed region."""
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        rows = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        inside = [r for r in rows if t_begin - 0.02 <= r[0] <= t_end + 0.02]
        window = "timed region"
        if len(inside) < 3:  # region shorter than the sampling period: use every sample taken under load (warm-up + timed)
            inside, window = rows, "warm-up + timed region (timed region shorter than 3 samples)"
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm = [r[1] for r in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max([r[2] for r in inside]) if inside else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU reference arm (the ONLY place bench.py executes anything under oracle/)
# --------------------------------------------------------------------------------------------
def cpu_reference_fps(cam, frames, pools, refs, budget_s, label):
    """frames/s of the reference's own CPU fuse_initialize_map on `frames` (list of (gray, depth, pose)),
    run as T independent instances x 10 std::threads each so that all host threads are used."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    kind = "reference" if os.path.exists(os.path.join(pyoracle.REFDIR, "libdsm_ref_mt.so")) else "port"
    ncpu = os.cpu_count() or 1
    Tmax = max(1, ncpu // 10) if kind == "reference" else max(1, ncpu)
    Tmax = min(Tmax, len(frames))
    mk = (lambda: pyoracle.RefMT(cam)) if kind == "reference" else (lambda: pyoracle.Restatement(cam))
    insts = [mk() for _ in range(Tmax)]

    def trial(T, n):
        def w(k):
            for i in range(k, n, T):
                g, d, p = frames[i]
                insts[k].fuse(refs[i], g, d, p, pools[i])
        th = [threading.Thread(target=w, args=(k,)) for k in range(T)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        return n / (time.perf_counter() - t0)

    # give the CPU its best configuration: every candidate instance count T (each instance forks 10 std::threads per
    # phase) runs the same frames twice -- at least 2 frames per instance and never fewer than 8 frames, so that a
    # trial is not dominated by start-up -- and the best pass counts; all rates are reported
    cands = sorted({1, 2, 3, 4, 6, 8, max(1, Tmax // 2), Tmax})
    cands = [c for c in cands if c <= Tmax]
    trial(1, 1)
    rates = {}
    for c in cands:
        nfr = min(len(frames), max(8, 2 * c))
        rates[c] = max(trial(c, nfr), trial(c, nfr))
    T = max(rates, key=rates.get)
    # warm-up + calibration on one frame
    g, d, p = frames[0]
    insts[0].fuse(refs[0], g, d, p, pools[0])
    t0 = time.perf_counter()
    insts[0].fuse(refs[0], g, d, p, pools[0])
    t_frame = time.perf_counter() - t0
    n = int(max(T, min(len(frames), budget_s * rates[T])))
    n = min(n, len(frames))

    def worker(k):
        for i in range(k, n, T):
            g, d, p = frames[i]
            insts[k].fuse(refs[i], g, d, p, pools[i])

    def run_once():
        th = [threading.Thread(target=worker, args=(k,)) for k in range(T)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        return time.perf_counter() - t0

    return {"run_once": run_once, "n": n, "T": T, "kind": kind, "t_frame_ms": t_frame * 1e3,
            "calibration_frames_per_s": {str(k): round(v, 1) for k, v in rates.items()},
            "cores": T * 10 if kind == "reference" else T,
            "sample": f"{n} of the step's {len(frames)} frames ({label}), {T} concurrent FusionFunctions instance(s)"
                      + (" x 10 std::threads (THREAD_NUM)" if kind == "reference" else " x 1 thread (restatement)")}


def run_reference_arm(args, rank, world):
    from densesurfelmapping_b200 import synth
    if rank != 0:
        return
    cam = synth.KITTI
    nfr = FRAMES_PER_GPU
    prev, cur = make_batch(cam, nfr, 0)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    from densesurfelmapping_b200.elements import SURFEL_DTYPE
    seeder = pyoracle.RefMT(cam) if pyoracle.have_reference() else pyoracle.Restatement(cam)
    pools = []
    for (g, d, p) in prev[:nfr]:
        _, new = seeder.fuse(0, g, d, p, np.zeros(0, SURFEL_DTYPE))
        pools.append(new)
    refs = [0] * nfr
    total_steps = args.steps + args.warmup
    arm = cpu_reference_fps(cam, cur, pools, refs, budget_s=max(0.2, 150.0 / max(total_steps, 1)), label="bounded so the whole run ends in minutes")
    for _ in range(args.warmup):
        arm["run_once"]()
    t = 0.0
    for _ in range(args.steps):
        t += arm["run_once"]()
    fps = arm["n"] * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"batch of independent synthetic KITTI-shaped 1226x370 frames, ~6k-surfel pool each; CPU step = {arm['n']} frames",
                   "frames_per_step": arm["n"], "host_cpus": os.cpu_count()},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": arm["cores"], "kind": arm["kind"], "sample": arm["sample"],
                         "instances_tried_frames_per_s": arm["calibration_frames_per_s"]},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class _DevView:
    """__cuda_array_interface__ view over a raw device pointer of the C-ABI library."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def kernel_alg_bytes(name, P, S, npool, nnew):
    """Compulsory (algorithmic) HBM bytes of ONE launch of a kernel over ONE frame: every input it
    needs read once, every output written once (DESIGN.md 'Kernels').  P pixels, S seeds; 0.9 P = listed member depths,
    0.8 P = plane-fit inliers (measured shares of the benchmark frames)."""
    return {
        "seed_init": 5 * S + 36 * S,
        "slic_assign_first": (1 + 4) * P + (4 + 4 + 1) * P + 24 * S,  # gray + depth in, inverse depth + labels + label codes out, seeds
        "slic_assign": (1 + 4 + 1) * P + 28 * S,                   # gray + inverse depth + label codes in (labels / codes rewritten where they change), seeds + flags
        "slic_gather": 6 * P + 4 * S + 20 * S + 4 * 0.9 * P,       # label codes + depth + gray in, integer sums + ordered depth lists out
        "slic_newton": 4 * 0.9 * P + 24 * S + 52 * S,              # lists + sums in, seed state out
        "plane_gather": 5 * P + 16 * S + 12 * 0.8 * P + (32 + 192) * S,  # label codes + depth in, centred points + per-seed sums out
        "plane_solve": (32 + 192 + 16) * S + 48 * S,
        "surfel_fuse": 88 * npool + 48 * S,
        "surfel_init": 52 * S + 44 * nnew,
    }.get(name, 0)


# kernels that are passes of one reference phase are judged together (update_pixels_kernel runs 3 times per frame)
FAMILY = {"slic_assign_first": "slic_assign", "slic_assign": "slic_assign"}


def parity_block(ctx, cam, cur, pools, offsets, frames):
    """Untimed: the state the context holds after the last timed step (labels, updated pools, new surfels of the batch)
    against the reference's own serial build (or the restatement pinned to it) on a sample of the benchmark frames."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    orc = pyoracle.RefSerial(cam) if pyoracle.have_reference() else pyoracle.Restatement(cam)
    local, news = ctx.batch_download()
    out = {"oracle": "reference fusion_functions.cpp, serial build (oracle/_ref/libdsm_ref_serial.so)" if pyoracle.have_reference()
           else "plain-C restatement (oracle/dsm_oracle.c)", "frames_checked": [int(b) for b in frames], "label_mismatches": 0,
           "count_mismatches": 0, "max_pos_err": 0.0, "max_normal_err": 0.0, "max_size_err": 0.0, "max_weight_err": 0.0, "int_field_mismatches": 0}
    for b in frames:
        g, d, p = cur[b]
        lo, no = orc.fuse(0, g, d, p, pools[b])
        out["label_mismatches"] += int((ctx.labels(b) != orc.labels()).sum())
        for got, want in ((local[offsets[b]:offsets[b + 1]], lo), (news[b], no)):
            if len(got) != len(want):
                out["count_mismatches"] += 1
                continue
            e = pyoracle.surfel_errors(got, want)
            out["max_pos_err"] = max(out["max_pos_err"], float(e["pos"]))
            out["max_normal_err"] = max(out["max_normal_err"], float(e["nrm"]))
            out["max_size_err"] = max(out["max_size_err"], float(e["size"]))
            out["max_weight_err"] = max(out["max_weight_err"], float(e["weight"]))
            out["int_field_mismatches"] += int(e["int_mismatch"])
    out["ok"] = bool(out["label_mismatches"] == 0 and out["count_mismatches"] == 0 and out["int_field_mismatches"] == 0 and
                     max(out["max_pos_err"], out["max_normal_err"], out["max_size_err"], out["max_weight_err"]) <= 1e-4)
    out["bar"] = "labels bit-exact; positions / normals / radii / weights within 1e-4 (norm-based), integer fields exact"
    return out


def run_extras(cam, local_rank, stream):
    """Secondary measurements reported next to the headline (not part of `value`):
    - stream: BASELINE configs[1], a sequential 1226x370 stream on the GPU-resident pool
      (dsm_fuse_frame_resident: per frame H2D of the image pair + all kernels + device-side compaction);
    - pool_transform: the loop-closure re-deformation kernel, a pure streaming kernel."""
    import torch
    from densesurfelmapping_b200 import capi, synth
    from densesurfelmapping_b200.elements import SURFEL_DTYPE
    out = {}
    T = 24
    cache = f"/tmp/dsm_bench_stream_{cam.width}x{cam.height}_{T}.npz"
    try:
        z = np.load(cache)
        G, D, Pz = z["g"], z["d"], z["p"]
    except Exception:
        fr = [synth.make_frame(cam, 5000 + t, synth.pose_stream(t)) for t in range(T)]
        G, D = np.stack([f[0] for f in fr]), np.stack([f[1] for f in fr])
        Pz = np.stack([synth.pose_stream(t) for t in range(T)])
        try:
            np.savez(cache, g=G, d=D, p=Pz)
        except Exception:
            pass
    tg, td = torch.from_numpy(G).pin_memory(), torch.from_numpy(D).pin_memory()
    hg, hd = tg.numpy(), td.numpy()
    ctx = capi.Context(cam, max_batch=2, max_local_surfels=4_000_000, device=local_rank, cuda_stream=stream.cuda_stream)
    ctx.pool_upload(np.zeros(0, SURFEL_DTYPE))
    warm = 6
    for t in range(warm):
        ctx.fuse_frame_resident(t // 4, hg[t], hd[t], Pz[t])
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps, nt = 3, 0
    e0.record()
    for rep in range(reps):  # the 18-frame drive is replayed; the pool keeps evolving on the device
        for t in range(warm, T):
            ctx.fuse_frame_resident((rep * T + t) // 4, hg[t], hd[t], Pz[t])
            nt += 1
    e1.record()
    ctx.sync()
    ms = e0.elapsed_time(e1)
    out["stream"] = {"workload": f"sequential synthetic 1226x370 stream (BASELINE configs[1]), GPU-resident pool, {nt} timed frames",
                     "frames_per_s": nt / (ms * 1e-3), "ms_per_frame": ms / nt, "final_pool_surfels": ctx.pool_size(),
                     "h2d_bytes_per_frame": int(cam.width * cam.height * 5 + 140), "api": "dsm_fuse_frame_resident (C ABI, pinned host frames)"}
    # where a single frame's time goes (plain launches with event pairs; latency-bound at batch 1)
    ctx.profile_enable((1 << capi.NUM_KERNELS) - 1)
    ctx.profile_reset()
    for t in range(warm, warm + 6):
        ctx.fuse_frame_resident(200 + t // 4, hg[t], hd[t], Pz[t])
    pms, pn = ctx.profile_read()
    ctx.profile_enable(0)
    kn = capi.kernel_names()
    out["stream"]["kernel_us_per_frame"] = {kn[i]: round(float(pms[i]) / 6 * 1e3, 1) for i in range(len(kn)) if pn[i]}
    # the same stream through dsm_fuse_stream_resident, n frames per call: the pose- and pool-independent stages of the n
    # frames run as one batch, only fuse / initialise / compaction run frame by frame (results identical to the
    # frame-by-frame stream, tests/test_gpu_resident.py); trades n-1 frames of latency for throughput
    out["stream_chunked"] = {}
    for chunk in (8, 16):
        c2 = capi.Context(cam, max_batch=2 * chunk, max_local_surfels=4_000_000, device=local_rank, cuda_stream=stream.cuda_stream)
        c2.pool_upload(np.zeros(0, SURFEL_DTYPE))
        usable = (T // chunk) * chunk

        def drive(rep):
            for t in range(0, usable, chunk):
                c2.fuse_stream_resident([(rep * T + t + i) // 4 for i in range(chunk)], hg[t:t + chunk], hd[t:t + chunk], Pz[t:t + chunk])
        drive(0)
        c2.sync()
        e0.record()
        for rep in range(1, 4):
            drive(rep)
        e1.record()
        c2.sync()
        msc = e0.elapsed_time(e1)
        out["stream_chunked"][str(chunk)] = {"frames_per_call": chunk, "frames_per_s": 3 * usable / (msc * 1e-3), "ms_per_frame": msc / (3 * usable),
                                             "final_pool_surfels": c2.pool_size(), "api": "dsm_fuse_stream_resident (C ABI, pinned host frames)"}
        c2.close()
    # the literal drop-in call (surfel_map.cpp:1066-1073 -> dsm_fuse_frame): host pointers, the caller's local surfels
    # cross PCIe both ways, the call synchronises -- what the reference node pays per frame when INTEGRATION.md's
    # three-line patch is applied; next to the reference's own single-instance CPU time for the same call
    pool_host = ctx.pool_download()
    c3 = capi.Context(cam, max_batch=1, max_local_surfels=max(len(pool_host), 1) + 64, device=local_rank)
    newbuf = np.zeros(c3.S, SURFEL_DTYPE)
    nn = ctypes.c_int(0)
    lat = []
    for t in range(warm, T):
        loc = pool_host.copy()
        t0 = time.perf_counter()
        rc = c3.lib.dsm_fuse_frame(c3.h, 300 + t // 4, hg[t].ctypes.data, hg[t].strides[0], hd[t].ctypes.data, hd[t].strides[0], Pz[t].ctypes.data,
                                   loc.ctypes.data if len(loc) else None, len(loc), newbuf.ctypes.data, c3.S, ctypes.byref(nn))
        lat.append(time.perf_counter() - t0)
        assert rc == 0
    c3.close()
    out["drop_in_call"] = {"api": "dsm_fuse_frame (host pointers, pool upload + download, synchronous): FusionFunctions::fuse_initialize_map as SurfelMap::fuse_map calls it",
                           "local_surfels": int(len(pool_host)), "ms_per_frame_median": float(np.median(lat[2:]) * 1e3), "ms_per_frame_max": float(max(lat[2:]) * 1e3)}
    # the reference's own SurfelMap node logic, compiled in place with the product's adapter under it
    # (oracle/_ref/libdsm_refmap_b200.so, INTEGRATION.md's three-line patch): frames/s of the whole node callback chain
    if os.environ.get("DSM_BENCH_NODE", "1") == "1":
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle"))
        import pyoracle
        if pyoracle.have_refmap(b200=True):
            node = pyoracle.RefMap(cam, drift_free_poses=10, b200=True)
            path, t0 = [], None
            for t in range(T):
                if t == warm:
                    t0 = time.perf_counter()
                p7 = pyoracle.pose_to_ros7(Pz[t])
                node.frame(10.0 + 0.1 * t, hg[t], hd[t], p7, t % 4 == 0, max(len(path) - 1, 0), path7=np.array(path).reshape(-1, 7))
                if t % 4 == 0:
                    path.append(p7)
            dt = time.perf_counter() - t0
            out["reference_node_over_product"] = {"frames_per_s": (T - warm) / dt, "ms_per_frame": dt / (T - warm) * 1e3,
                                                  "local_surfels": len(node.local()),
                                                  "what": "unmodified surfel_map.cpp callbacks (pose feed, image, depth) with dsm::FusionFunctions "
                                                          "in place of FusionFunctions; host-pointer dsm_fuse_frame per frame, wall clock (includes the node's own CPU work)"}
            node.close()
    # loop-closure transform on a large pool
    n = 4_000_000
    rng = np.random.RandomState(7)
    big = np.zeros(n, SURFEL_DTYPE)
    for f in ("px", "py", "pz", "nx", "ny", "nz"):
        big[f] = rng.standard_normal(n).astype(np.float32)
    big["update_times"] = 1
    ctx.pool_upload(big)
    a = np.deg2rad(2.0)
    Wm = np.eye(4)
    Wm[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
    Wm[:3, 3] = [0.5, 0.0, 0.25]
    w = np.ascontiguousarray(Wm.T.astype(np.float32).reshape(16))
    for _ in range(3):
        ctx.pool_transform(w)
    ctx.sync()
    e0.record()
    for _ in range(10):
        ctx.pool_transform(w)
    e1.record()
    ctx.sync()
    ms = e0.elapsed_time(e1) / 10
    out["pool_transform"] = {"surfels": n, "ms": ms, "algorithmic_GBps": 48 * n / (ms * 1e-3) / 1e9, "moved_GBps": 88 * n / (ms * 1e-3) / 1e9,
                             "note": "48 B/surfel compulsory (p,n read+write); the 44-byte ABI records are moved whole (88 B/surfel)"}
    ctx.close()
    out["hd_loop_closure_stream"] = run_cfg5(local_rank, stream)
    return out


def run_cfg5(local_rank, stream):
    """BASELINE configs[4]: a 1280x720 VINS-style stream (the reference's RGBD constant set, fusion_functions.h:17-21) on the
    GPU-resident map with loop-closure pose updates: keyframes leaving the drift-free window move to the device-resident
    inactive store (SurfelMap::move_add_surfels, surfel_map.cpp:1479-1497); every 25 frames a loop correction re-deforms
    the active pool (warp_active_surfels_cpu_kernel :750-789) and every inactive pose's surfels (warp_inactive_surfels_cpu_kernel
    :681-748).  Reports frames/s of the whole loop and the cost of one re-deformation event."""
    import torch
    from densesurfelmapping_b200 import capi, synth
    from densesurfelmapping_b200.elements import SURFEL_DTYPE
    cam = synth.Camera(1280, 720, 720.0, 720.0, 639.5, 359.5, 0.3, 8.0)
    T = 30
    cache = f"/tmp/dsm_bench_hd_{T}.npz"
    try:
        z = np.load(cache)
        G, D, Pz = z["g"], z["d"], z["p"]
    except Exception:
        fr = [synth.make_frame(cam, 7000 + t, synth.pose_stream(t, step_m=0.2)) for t in range(T)]
        G = np.stack([f[0] for f in fr])
        D = (np.stack([f[1] for f in fr]) * np.float32(0.2)).astype(np.float32)  # indoor range
        Pz = np.stack([synth.pose_stream(t, step_m=0.2) for t in range(T)])
        try:
            np.savez(cache, g=G, d=D, p=Pz)
        except Exception:
            pass
    tg, td = torch.from_numpy(G).pin_memory(), torch.from_numpy(D).pin_memory()
    hg, hd = tg.numpy(), td.numpy()
    ctx = capi.Context(cam, max_batch=2, max_local_surfels=3_000_000, device=local_rank, cuda_stream=stream.cuda_stream)
    ctx.set_constants(capi.CONSTANTS_RGBD)
    ctx.pool_upload(np.zeros(0, SURFEL_DTYPE))
    ctx.inactive_reserve(6_000_000)
    a = np.deg2rad(2.0)
    Wm = np.eye(4)
    Wm[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
    Wm[:3, 3] = [0.05, 0.0, 0.02]
    w = np.ascontiguousarray(Wm.T.astype(np.float32).reshape(16))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    frames, events_ms, retired = 0, [], []

    def one(fid, t):
        kf = fid // 4  # every 4th frame is a keyframe; the reference index is the newest keyframe
        ctx.fuse_frame_resident(kf, hg[t], hd[t], Pz[t])
        if fid % 4 == 3 and kf >= 10:  # the keyframe that leaves the 10-pose drift-free window
            ctx.inactive_retire(kf - 10)
            retired.append(kf - 10)
        if fid % 25 == 24:  # loop closure: active pool + every inactive pose
            ea.record()
            ctx.pool_transform(w)
            for k in retired:
                ctx.inactive_transform(k, w)
            eb.record()
            ctx.sync()
            events_ms.append(ea.elapsed_time(eb))
    for fid in range(8):
        one(fid, fid % T)
    ctx.sync()
    e0.record()
    for fid in range(8, 8 + 3 * T):
        one(fid, fid % T)
        frames += 1
    e1.record()
    ctx.sync()
    ms = e0.elapsed_time(e1)
    ninact, nseg = ctx.inactive_size()
    res = {"workload": f"1280x720 stream, RGBD constant set, {frames} timed frames, keyframe every 4 frames, drift-free window 10, loop closure every 25 frames",
           "frames_per_s": frames / (ms * 1e-3), "ms_per_frame": ms / frames, "active_surfels": ctx.pool_size(), "inactive_surfels": int(ninact),
           "inactive_poses": int(nseg), "loop_closures": len(events_ms), "ms_per_loop_closure": float(np.mean(events_ms)) if events_ms else None,
           "api": "dsm_fuse_frame_resident + dsm_inactive_retire + dsm_pool_transform + dsm_inactive_transform (C ABI)"}
    ctx.close()
    return res


def gpu_numa_cpus(local_rank):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None when the box does not say.  Used only while the pinned
    host buffers of the e2e leg are allocated: pinned pages land on the allocating thread's node, and at N=8 eight
    55 GB/s copy streams should not all read from one socket's DRAM."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None, None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        return (node, cpus) if cpus else (None, None)
    except Exception:
        return None, None


def run_gpu_arm(args, rank, world, local_rank):
    if args.contexts is None:
        args.contexts = 3 if world > 1 else 2  # N>1: one more batch in flight covers the gather's pack kernels (measured at N=2: 0.901 vs 0.920 ms)
    if world > 1 and args.contexts < 2:
        args.contexts = 2
    import torch
    import torch.distributed as dist
    from densesurfelmapping_b200 import capi, synth
    from densesurfelmapping_b200.elements import SURFEL_DTYPE

    torch.cuda.set_device(local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a few hundred ms to come up: start it long before the timed region
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cam = synth.KITTI
    B = FRAMES_PER_GPU
    P, S = cam.width * cam.height, (cam.width // 8) * (cam.height // 8)

    prev, cur = make_batch(cam, B, rank)
    # a dedicated (non-default) torch stream: the library enqueues everything on it, so torch.cuda.Event
    # timing, torch.distributed collectives and the kernels all share one ordered stream
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = capi.Context(cam, max_batch=B, max_local_surfels=B * S + 64, device=local_rank, cuda_stream=stream.cuda_stream)
    empty = np.zeros(0, SURFEL_DTYPE)
    zofs = np.zeros(B + 1, np.int32)
    # pools: the predecessor views run through initialise on the GPU path itself (untimed setup)
    _, pools = ctx.fuse_batch([0] * B, np.stack([f[0] for f in prev]), np.stack([f[1] for f in prev]),
                              np.stack([f[2] for f in prev]), empty, zofs)
    offsets = np.concatenate([[0], np.cumsum([len(p) for p in pools])]).astype(np.int32)
    npool = int(offsets[-1])
    # pinned host buffers for the e2e path, allocated on the GPU's NUMA node
    numa_node, numa_cpus = gpu_numa_cpus(local_rank)
    aff0 = os.sched_getaffinity(0)
    if numa_cpus:
        os.sched_setaffinity(0, numa_cpus)

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    t_gray, h_gray = pinned(np.stack([f[0] for f in cur]))
    t_depth, h_depth = pinned(np.stack([f[1] for f in cur]))
    t_pose, h_pose = pinned(np.stack([f[2] for f in cur]).astype(np.float32))
    pool_np = np.concatenate(pools) if npool else empty
    t_pool, h_pool = pinned(pool_np.view(np.uint8))
    # dsm_fuse_batch updates `local` in place (like the reference), so every e2e step gets its own
    # pre-filled pinned copy of the pool: no host-side reset inside the timed region
    ring = [pinned(pool_np.view(np.uint8) if npool else np.zeros(44, np.uint8)) for _ in range(min(args.steps + 2, 48))]
    out_bufs = [(pinned(np.zeros(B * S * 44, np.uint8))[1], pinned(np.zeros(B, np.int32))[1]) for _ in range(2)]
    h_new, h_cnt = out_bufs[0]
    if numa_cpus:
        os.sched_setaffinity(0, aff0)  # only the allocations above were bound (the CPU baseline leg uses every core)
    refs = np.zeros(B, np.int32)
    L = ctx.lib

    # e2e: two contexts used alternately through the public async call, so the H2D copies of step k+1 overlap
    # the kernels of step k.  Every step's inputs come from pinned host memory and every step's results
    # (updated pools, new surfels, counts) land in pinned host memory, all inside the timed region.
    ctx2 = capi.Context(cam, max_batch=B, max_local_surfels=B * S + 64, device=local_rank)
    e2e_ctx = [ctx, ctx2]
    e2e_count = [0]

    def e2e_step():
        k = e2e_count[0]
        e2e_count[0] += 1
        c = e2e_ctx[k & 1]
        hn, hc = out_bufs[k & 1]
        h_pool_io = ring[k % len(ring)][1]
        rc = L.dsm_batch_wait(c.h)  # the batch issued on this context two steps ago
        assert rc == 0, L.dsm_last_error(c.h)
        rc = L.dsm_fuse_batch_async(c.h, B, refs.ctypes.data, h_gray.ctypes.data, h_depth.ctypes.data, h_pose.ctypes.data,
                                    h_pool_io.ctypes.data, offsets.ctypes.data, hn.ctypes.data, hc.ctypes.data)
        assert rc == 0, L.dsm_last_error(c.h)

    def e2e_drain():
        for c in e2e_ctx:
            assert L.dsm_batch_wait(c.h) == 0

    # ---- resident mode: upload once.  Two contexts used alternately (at every N): a step's kernels are enqueued without
    # waiting for the previous step, so consecutive batches overlap on the GPU the way a caller with a queue of batches
    # runs them (the e2e leg below does the same through the host-buffer call).  `value` is the steady-state throughput.
    ctx.batch_upload(refs, h_gray, h_depth, h_pose, pool_np, offsets)
    ctx2.batch_upload(refs, h_gray, h_depth, h_pose, pool_np, offsets)
    res_ctx = [ctx, ctx2] if args.contexts >= 2 else [ctx]
    ctx3 = None
    if args.contexts == 3:
        ctx3 = capi.Context(cam, max_batch=B, max_local_surfels=B * S + 64, device=local_rank)
        ctx3.batch_upload(refs, h_gray, h_depth, h_pose, pool_np, offsets)
        res_ctx = [ctx, ctx2, ctx3]
    if world > 1:
        # Multi-GPU step = this rank's kernels + ONE gather of its surfel deltas onto rank 0 through the C ABI
        # (dsm_gather_deltas: a pack kernel that writes the valid records straight into the rank's slot of the root's
        # receive buffer over NVLink peer memory; nothing waits on the host).  The gather of step k-1 is issued right after
        # step k's kernels have been enqueued: it waits for its batch by event on a side stream while the other context's
        # batch keeps the GPU busy, exactly as the two contexts overlap at N=1.  Each context has its own communicator
        # (ids broadcast over torch.distributed).
        ids = [capi.comm_unique_id() for _ in res_ctx] if rank == 0 else [None for _ in res_ctx]
        dist.broadcast_object_list(ids, src=0)
        for c, uid in zip(res_ctx, ids):
            c.comm_init(uid, rank, world)
    step_no = [0]
    GATHER_LAG = max(len(res_ctx) - 1, 1)

    def step():
        k = step_no[0]
        step_no[0] += 1
        c = res_ctx[k % len(res_ctx)]
        c.batch_restore_pool()
        c.batch_run()
        if world > 1 and k >= GATHER_LAG:
            res_ctx[(k - GATHER_LAG) % len(res_ctx)].gather_deltas(0)

    def drain():
        if world > 1:
            for k in range(max(step_no[0] - GATHER_LAG, 0), step_no[0]):  # the steps whose gather has not been issued yet
                res_ctx[k % len(res_ctx)].gather_deltas(0)
            for c in res_ctx:
                c.gather_wait()
            step_no[0] = 0
        for c in res_ctx:
            c.sync()

    def barrier_sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up; per-kernel breakdown measured during the warm-up steps on ONE context, one step at a time, with the
    # batch as ONE launch sequence (concurrency 1): in the production schedule sub-batches and consecutive steps run
    # concurrently on several streams and stretch each other's event-pair durations, which would not be per-kernel
    # times any more
    ctx.set_concurrency(1)
    ctx.profile_enable(((1 << capi.NUM_KERNELS) - 1) & ~(1 << capi.kernel_names().index("repack")))
    ctx.profile_reset()
    nwarm = max(args.warmup, 3)
    for _ in range(nwarm):
        ctx.batch_restore_pool()
        ctx.batch_run()
        ctx.sync()
    ms, nl = ctx.profile_read()
    names = capi.kernel_names()
    warm_on_ctx = nwarm
    kernel_ms = {names[i]: float(ms[i]) / warm_on_ctx for i in range(len(names)) if nl[i]}  # ms per step
    per_launch_ms = {names[i]: float(ms[i] / nl[i]) for i in range(len(names)) if nl[i]}
    # dominant kernel = the reference phase with the largest share of the step; the passes of one phase count together
    fam_ms = {}
    for k, v in kernel_ms.items():
        fam_ms[FAMILY.get(k, k)] = fam_ms.get(FAMILY.get(k, k), 0.0) + v
    dom = max(fam_ms, key=fam_ms.get)
    dom_members = [k for k in kernel_ms if FAMILY.get(k, k) == dom]
    dom_mask = 0
    for k in dom_members:
        dom_mask |= 1 << names.index(k)
    launches_per_step = int((sum(nl) + nl[names.index("slic_newton")]) // warm_on_ctx)  # the Newton stage is two launches (k_newton2 + k_newton_hard) under one event pair
    nnew_avg = float(np.mean([len(p) for p in ctx.batch_download()[1]]))
    dom_launches_per_step = {k: int(nl[names.index(k)]) // warm_on_ctx for k in dom_members}
    for c in res_ctx:
        c.profile_enable(0)
        c.set_concurrency(args.sub_batches)
    for _ in range(3 * len(res_ctx)):  # untimed: the schedules of the sub-batches are captured (CUDA graphs) on first use
        step()
    drain()
    ctx.profile_enable(dom_mask)  # inside the timed region only the dominant phase's kernels carry events
    ctx.profile_reset()

    # ---- timed region: value (inputs resident in HBM)
    barrier_sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    drain()
    e1.record()
    barrier_sync()
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    dms, dn = ctx.profile_read()
    dom_steps = max(1, int(min(dn[names.index(k)] // max(dom_launches_per_step[k], 1) for k in dom_members)))  # steps profiled on this context
    dom_ms = float(sum(dms[names.index(k)] for k in dom_members)) / dom_steps  # ms per step spent in the dominant phase
    ctx.profile_enable(0)
    parity = parity_block(ctx, cam, cur, pools, offsets, sorted({0, B // 3, (2 * B) // 3, B - 1})) if rank == 0 else None
    t = torch.tensor([ms_total], device=f"cuda:{local_rank}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())

    # ---- e2e: through the C-ABI with pinned host buffers, copies inside the timed region
    for _ in range(4):
        e2e_step()
    e2e_drain()
    # PCIe context for the e2e number: pinned H2D rate of one contiguous copy of the step's depth array
    dtmp = torch.empty(t_depth.numel(), dtype=torch.float32, device=f"cuda:{local_rank}")
    dtmp.copy_(t_depth.view(-1), non_blocking=True)
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    dtmp.copy_(t_depth.view(-1), non_blocking=True)
    eb.record()
    torch.cuda.synchronize()
    h2d_gbs = t_depth.numel() * 4 / (ea.elapsed_time(eb) * 1e-3) / 1e9
    del dtmp
    barrier_sync()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e2e_drain()  # host has seen both contexts finish; e1 is recorded after that
    e1.record()
    barrier_sync()
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], device=f"cuda:{local_rank}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d = B * P * 5 + B * 64 + npool * 44 + (B + 1) * 4 + B * 4
    d2h = npool * 44 + B * S * 44 + B * 4

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        alg = sum(kernel_alg_bytes(k, P, S, npool / B, nnew_avg) * dom_launches_per_step[k] for k in dom_members) * B  # bytes per step
        achieved = alg / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        path_bytes = (9 * P + 60 * S) * B + 88 * npool + 44 * nnew_avg * B
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:
            pass
        value = world * B * args.steps / (ms_total * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"batch={B} independent synthetic KITTI-shaped 1226x370 depth+gray frames per GPU, own pose and "
                                   f"~{npool // B}-surfel local pool each (BASELINE configs[2]/[3]); superpixel+normal+plane-fit+fuse+initialise per frame",
                       "frames_per_gpu_per_step": B, "sub_batches": args.sub_batches, "resident_contexts": len(res_ctx), "pool_surfels_per_frame": npool // B, "new_surfels_per_frame": nnew_avg,
                       "l2": f"per-step working set {(B * (13.6 * P + 200 * S) + 88 * npool) / 1e6:.0f} MB > 126 MB L2 (inputs larger than L2)",
                       "parallelism": f"frames sharded {B}/GPU, no data-path collective; one gather of the valid surfel deltas per step through the C ABI (dsm_gather_deltas: pack kernel writing into the root's slots over NVLink peer memory)" if world > 1 else "single GPU"},
            "e2e": {"value": world * B * args.steps / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps, "api": "dsm_fuse_batch_async + dsm_batch_wait on two alternating contexts (C ABI, pinned host buffers)",
                    "pinned_h2d_gbs": h2d_gbs, "pinned_numa_node": numa_node},
            "gpu_launches": launches_per_step * args.sub_batches * args.steps,  # every sub-batch runs the whole launch sequence on its frames
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "kernels": {k: dom_launches_per_step[k] for k in dom_members},
                         "achieved": alg / (fam_ms[dom] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (fam_ms[dom] * 1e-3) / 1e9 / peak,
                         "traffic": traffic, "traffic_source": "ncu --set full capture of tools/prof_step.py (same 32-frame step), profiles/traffic.json (bytes per step of the phase)",
                         "alg_bytes_per_step": alg, "kernel_ms_per_step": fam_ms[dom], "peak_source": peak_src,
                         "timing": "CUDA event pairs around every launch of the phase, on the launching stream, live in this run: the "
                                   f"{nwarm} profiled steps before the timed region, whole batch as ONE launch sequence, one step at a time "
                                   "(the kernel has the GPU to itself, like in the ncu launch list)",
                         "in_timed_region": {"kernel_ms_per_step": dom_ms, "achieved": achieved, "frac": achieved / peak,
                                             "timing": f"same event pairs inside the timed region, where {args.sub_batches} sub-batches x {len(res_ctx)} "
                                                       "contexts run concurrently: these durations include the time the kernel shares the SMs "
                                                       "with other streams' kernels (they add up to more than the step)"},
                         "path": {"alg_bytes_per_step": path_bytes, "achieved": path_bytes / (ms_total / args.steps * 1e-3) / 1e9,
                                  "frac": path_bytes / (ms_total / args.steps * 1e-3) / 1e9 / peak}},
            "kernel_ms_per_step": kernel_ms, "kernel_ms_per_launch": per_launch_ms,
            "kernel_timing": "event pairs per launch during the warm-up steps, whole batch as ONE launch sequence (dsm_set_concurrency(1)); "
                             "their sum exceeds ms_per_step, which runs the production schedule of concurrent sub-batches",
            "parity": parity,
        }
        if world == 1 and os.environ.get("DSM_BENCH_NO_EXTRAS") != "1":
            line["extras"] = run_extras(cam, local_rank, stream)
        if world == 1 and not args.no_cpu:
            pools_cpu = [pool_np[offsets[b]:offsets[b + 1]] for b in range(B)]
            arm = cpu_reference_fps(cam, cur, pools_cpu, [0] * B, budget_s=20.0, label="~10-30 s of CPU work")
            arm["run_once"]()
            tt = arm["run_once"]()
            tt = min(tt, arm["run_once"]())
            line["cpu_baseline"] = {"value": arm["n"] / tt, "unit": "frames/s", "cores": arm["cores"], "kind": arm["kind"],
                                    "sample": arm["sample"], "host_cpus": os.cpu_count(), "single_instance_ms_per_frame": arm["t_frame_ms"],
                                    "instances_tried_frames_per_s": arm["calibration_frames_per_s"]}
        print(json.dumps(line), flush=True)
    ctx2.close()
    if ctx3 is not None:
        ctx3.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--contexts", type=int, default=None, choices=(1, 2, 3), help="resident contexts used in rotation by the timed loop (default: 2 at N=1, 3 at N>1)")
    ap.add_argument("--sub-batches", type=int, default=2, help="concurrent sub-batches of dsm_batch_run (C ABI default 2)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
