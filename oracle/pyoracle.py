"""TEST INFRASTRUCTURE ONLY — ctypes access to the CPU oracles.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Never from the product path (the product is the CUDA library
behind include/dsm.h and fails loudly if it is missing).

Three libraries under oracle/_ref/ (built by oracle/Makefile):
  libdsm_ref_serial.so  the reference's own fusion_functions.cpp, thread bodies inline  -> PARITY ORACLE
  libdsm_ref_mt.so      the reference's own fusion_functions.cpp as shipped (10 threads) -> timed CPU baseline
  libdsm_oracle.so      this repo's plain-C restatement (oracle/dsm_oracle.c), pinned to ref_serial
"""
import ctypes
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")

import sys
sys.path.insert(0, os.path.dirname(HERE))
from densesurfelmapping_b200.elements import SEED_DTYPE, SURFEL_DTYPE, num_seeds  # noqa: E402


def build(verbose=False):
    """Build the restatement (always) and the reference variants (when /root/reference exists)."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def _lib(name):
    path = os.path.join(REFDIR, name)
    if not os.path.exists(path):
        build()
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return ctypes.CDLL(path)


_u8p = ctypes.POINTER(ctypes.c_uint8)
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


class _Base:
    """Common wrapper; both the reference driver and the restatement export the same entry points
    with prefix dsmref_ / dsmor_."""
    prefix = None
    libname = None

    def __init__(self, cam):
        self.cam = cam
        self.lib = _lib(self.libname)
        p = self.prefix
        L = self.lib
        getattr(L, p + "create").restype = ctypes.c_void_p
        getattr(L, p + "create").argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_float] * 6
        getattr(L, p + "destroy").argtypes = [ctypes.c_void_p]
        getattr(L, p + "fuse").restype = ctypes.c_int
        getattr(L, p + "fuse").argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        getattr(L, p + "superpixels").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        getattr(L, p + "get_labels").argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        getattr(L, p + "get_seeds").argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        getattr(L, p + "get_norm_map").argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        self.h = getattr(L, p + "create")(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.far, cam.near)
        self.S = num_seeds(cam.width, cam.height)

    def close(self):
        if self.h:
            getattr(self.lib, self.prefix + "destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fuse(self, ref_idx, gray, depth, pose, local):
        """fuse_initialize_map semantics: returns (local_updated, new_surfels)."""
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        pose = np.ascontiguousarray(pose, dtype=np.float32).reshape(16)
        local = np.array(local, dtype=SURFEL_DTYPE, copy=True)
        new = np.zeros(self.S, dtype=SURFEL_DTYPE)
        n = getattr(self.lib, self.prefix + "fuse")(
            self.h, int(ref_idx), gray.ctypes.data, depth.ctypes.data, pose.ctypes.data,
            local.ctypes.data if len(local) else None, len(local), new.ctypes.data, self.S)
        return local, new[:n].copy()

    def superpixels(self, gray, depth):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        getattr(self.lib, self.prefix + "superpixels")(self.h, gray.ctypes.data, depth.ctypes.data)
        return self.labels(), self.seeds()

    def labels(self):
        out = np.empty((self.cam.height, self.cam.width), dtype=np.int32)
        getattr(self.lib, self.prefix + "get_labels")(self.h, out.ctypes.data)
        return out

    def seeds(self):
        out = np.zeros(self.S, dtype=SEED_DTYPE)
        getattr(self.lib, self.prefix + "get_seeds")(self.h, out.ctypes.data)
        return out

    def norm_map(self):
        out = np.empty((self.cam.height, self.cam.width, 3), dtype=np.float32)
        getattr(self.lib, self.prefix + "get_norm_map")(self.h, out.ctypes.data)
        return out


class RefSerial(_Base):
    prefix = "dsmref_"
    libname = "libdsm_ref_serial.so"


class RefMT(_Base):
    prefix = "dsmref_"
    libname = "libdsm_ref_mt.so"


class Restatement(_Base):
    prefix = "dsmor_"
    libname = "libdsm_oracle.so"

    def debug_iters(self, gray, depth, iters, last_with_update=True):
        """seed init + `iters` assign passes (last update_seeds optional); returns (labels, seeds)."""
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        self.lib.dsmor_debug_iters.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        self.lib.dsmor_debug_iters(self.h, gray.ctypes.data, depth.ctypes.data, int(iters), int(bool(last_with_update)))
        return self.labels(), self.seeds()


def fuse_map_poststep(local, new):
    """SurfelMap::fuse_map post-step (surfel_map.cpp:1077-1109) via the restatement; returns the new pool."""
    lib = _lib("libdsm_oracle.so")
    lib.dsmor_fuse_map_poststep.restype = ctypes.c_int
    lib.dsmor_fuse_map_poststep.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    buf = np.zeros(len(local) + len(new) + 1, dtype=SURFEL_DTYPE)
    buf[:len(local)] = local
    new = np.ascontiguousarray(new, dtype=SURFEL_DTYPE)
    n = lib.dsmor_fuse_map_poststep(buf.ctypes.data, len(local), new.ctypes.data if len(new) else None, len(new))
    return buf[:n].copy()


def retire(local, kf):
    """move_add_surfels removal half (surfel_map.cpp:1479-1497) via the restatement: returns (local_after, retired)."""
    lib = _lib("libdsm_oracle.so")
    lib.dsmor_retire.restype = ctypes.c_int
    lib.dsmor_retire.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    loc = np.array(local, dtype=SURFEL_DTYPE, copy=True)
    out = np.zeros(len(loc) + 1, dtype=SURFEL_DTYPE)
    n = lib.dsmor_retire(loc.ctypes.data if len(loc) else None, len(loc), int(kf), out.ctypes.data)
    return loc, out[:n].copy()


def warp_active(surfels, W_colmajor):
    """warp_active_surfels_cpu_kernel (surfel_map.cpp:750-789) via the restatement."""
    lib = _lib("libdsm_oracle.so")
    lib.dsmor_warp_active.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    out = np.array(surfels, dtype=SURFEL_DTYPE, copy=True)
    w = np.ascontiguousarray(W_colmajor, dtype=np.float32).reshape(16)
    lib.dsmor_warp_active(out.ctypes.data if len(out) else None, len(out), w.ctypes.data)
    return out


def have_reference():
    return os.path.exists(os.path.join(REFDIR, "libdsm_ref_serial.so"))


# ---- comparison metrics (norm-based, not per-component relative) ----
def surfel_errors(a, b):
    """Returns dict of max errors between two equally long SURFEL_DTYPE arrays."""
    assert len(a) == len(b)
    if len(a) == 0:
        return dict(pos=0.0, nrm=0.0, size=0.0, weight=0.0, color=0.0, int_mismatch=0)
    pa = np.stack([a["px"], a["py"], a["pz"]], -1).astype(np.float64)
    pb = np.stack([b["px"], b["py"], b["pz"]], -1).astype(np.float64)
    na = np.stack([a["nx"], a["ny"], a["nz"]], -1).astype(np.float64)
    nb = np.stack([b["nx"], b["ny"], b["nz"]], -1).astype(np.float64)
    pos = np.linalg.norm(pa - pb, axis=1) / np.maximum(np.linalg.norm(pb, axis=1), 1e-12)
    nrm = np.linalg.norm(na - nb, axis=1)
    size = np.abs(a["size"].astype(np.float64) - b["size"]) / np.maximum(np.abs(b["size"]), 1e-12)
    wgt = np.abs(a["weight"].astype(np.float64) - b["weight"]) / np.maximum(np.abs(b["weight"]), 1e-12)
    col = np.abs(a["color"].astype(np.float64) - b["color"])
    im = int(np.sum(a["update_times"] != b["update_times"]) + np.sum(a["last_update"] != b["last_update"]))
    return dict(pos=float(pos.max()), nrm=float(nrm.max()), size=float(size.max()),
                weight=float(wgt.max()), color=float(col.max()), int_mismatch=im)
