"""The C++ drop-in: include/dsm_fusion_functions.hpp mirrors the reference's FusionFunctions class.
CPU: it compiles against cv::Mat / Eigen-like caller types and links to the C ABI; without a GPU it
fails loudly (exit code 3, no fallback).  GPU: its results match the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from densesurfelmapping_b200 import synth
from densesurfelmapping_b200.elements import SURFEL_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "densesurfelmapping_b200")
BIN = os.path.join(ROOT, "tests", "cpp", "adapter_main")


def build_adapter():
    src = os.path.join(ROOT, "tests", "cpp", "adapter_main.cpp")
    if os.path.exists(BIN) and os.path.getmtime(BIN) > max(os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "include", "dsm_fusion_functions.hpp"))):
        return
    cmd = ["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "shim"), src,
           "-o", BIN, "-L", PKG, "-ldsm_b200", f"-Wl,-rpath,{PKG}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def run_adapter(tmp, cam, ref, pose, gray, depth, local):
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<i", ref) + np.asarray(pose, np.float32).tobytes() + struct.pack("<i", len(local)))
        f.write(gray.tobytes() + depth.tobytes() + local.tobytes())
    args = [BIN, str(cam.width), str(cam.height)] + [repr(float(v)) for v in (cam.fx, cam.fy, cam.cx, cam.cy, cam.far, cam.near)] + [fin, fout]
    r = subprocess.run(args, capture_output=True, text=True)
    return r, fout


def test_adapter_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    build_adapter()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    cam = synth.VGA
    g, d = synth.make_frame(cam, 0)
    r, _ = run_adapter(str(tmp_path), cam, 0, synth.identity_pose(), g, d, np.zeros(0, SURFEL_DTYPE))
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr


@pytest.mark.gpu
def test_adapter_matches_oracle(tmp_path):
    from util import check_surfels, oracle_for
    build_adapter()
    cam = synth.VGA
    g, d = synth.make_frame(cam, 0)
    pose = synth.identity_pose()
    orc = oracle_for(cam)
    _, pool = orc.fuse(0, g, d, pose, np.zeros(0, SURFEL_DTYPE))
    want_local, want_new = orc.fuse(1, g, d, pose, pool)
    r, fout = run_adapter(str(tmp_path), cam, 1, pose, g, d, pool)
    assert r.returncode == 0, r.stderr
    raw = open(fout, "rb").read()
    nl = struct.unpack_from("<i", raw, 0)[0]
    local = np.frombuffer(raw, SURFEL_DTYPE, nl, 4)
    nn = struct.unpack_from("<i", raw, 4 + 44 * nl)[0]
    new = np.frombuffer(raw, SURFEL_DTYPE, nn, 8 + 44 * nl)
    check_surfels(local, want_local, "adapter local")
    check_surfels(new, want_new, "adapter new")
