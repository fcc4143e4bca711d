"""ctypes binding of the C ABI in include/dsm.h (the reference-facing boundary).

`FusionFunctions` mirrors the reference class of the same name (fusion_functions.h:23-95): the
same two public methods, same argument meaning.  It is a thin veneer: every call goes straight
into libdsm_b200.so; there is no Python or CPU implementation behind it, and loading fails
loudly when the library (or a B200) is missing.
"""
import ctypes
import os
import numpy as np

from .elements import POINT_DTYPE, SEED_DTYPE, SURFEL_DTYPE, num_seeds

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdsm_b200.so")

DSM_OK = 0
ERRORS = {-1: "DSM_E_INVALID", -2: "DSM_E_SHAPE", -3: "DSM_E_NODEVICE", -4: "DSM_E_CUDA",
          -5: "DSM_E_NOMEM", -6: "DSM_E_CAPACITY", -7: "DSM_E_STATE", -8: "DSM_E_NCCL", -9: "DSM_E_IO"}
NUM_KERNELS = 10

# every symbol include/dsm.h declares (tests/test_abi.py checks the library exports all of them)
EXPORTS = [
    "dsm_version", "dsm_strerror", "dsm_last_error", "dsm_create", "dsm_destroy", "dsm_num_seeds",
    "dsm_fuse_frame", "dsm_batch_upload", "dsm_batch_run", "dsm_batch_download", "dsm_sync",
    "dsm_fuse_batch", "dsm_fuse_batch_async", "dsm_batch_wait", "dsm_batch_restore_pool", "dsm_pool_upload", "dsm_fuse_frame_resident",
    "dsm_pool_transform", "dsm_pool_retire", "dsm_pool_append", "dsm_pool_size", "dsm_pool_download", "dsm_get_labels", "dsm_get_seeds",
    "dsm_debug_stop_after", "dsm_debug_invariant_violations", "dsm_profile_enable", "dsm_profile_reset", "dsm_profile_read", "dsm_kernel_name", "dsm_device_buffer",
    "dsm_pool_export_cloud", "dsm_pool_export_surfels", "dsm_write_pcd", "dsm_write_ply_mesh", "dsm_mesh_vertices", "dsm_fuse_stream_resident",
    "dsm_inactive_reserve", "dsm_inactive_retire", "dsm_inactive_reactivate", "dsm_inactive_transform", "dsm_inactive_export_cloud",
    "dsm_inactive_download", "dsm_inactive_size",
    "dsm_comm_unique_id", "dsm_comm_init", "dsm_comm_destroy", "dsm_gather_deltas", "dsm_gather_wait", "dsm_gathered_device",
    "dsm_gathered_rank_bytes", "dsm_gathered_download", "dsm_set_constants", "dsm_set_concurrency",
]


class DsmParams(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32),
                ("fx", ctypes.c_float), ("fy", ctypes.c_float), ("cx", ctypes.c_float), ("cy", ctypes.c_float),
                ("fuse_far", ctypes.c_float), ("fuse_near", ctypes.c_float),
                ("max_batch", ctypes.c_int32), ("max_local_surfels", ctypes.c_int32)]


class DsmConstants(ctypes.Structure):
    _fields_ = [("huber_range", ctypes.c_double), ("baseline", ctypes.c_double), ("disparity_error", ctypes.c_double),
                ("min_tolerate_diff", ctypes.c_double)]


CONSTANTS_DRIVE = (0.4, 0.5, 4.0, 0.1)    # fusion_functions.h:13-16
CONSTANTS_RGBD = (0.05, 0.08, 1.0, 0.05)  # fusion_functions.h:18-21


class DsmError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__(f"{ERRORS.get(code, code)}: {detail}")


_lib = None


def load_library():
    """Loads libdsm_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not built: run `python -m densesurfelmapping_b200.build` "
                                "(or __graft_entry__.build()); this package has no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    L.dsm_version.restype = ci
    L.dsm_strerror.restype = ctypes.c_char_p
    L.dsm_strerror.argtypes = [ci]
    L.dsm_last_error.restype = ctypes.c_char_p
    L.dsm_last_error.argtypes = [vp]
    L.dsm_kernel_name.restype = ctypes.c_char_p
    L.dsm_kernel_name.argtypes = [ci]
    L.dsm_create.argtypes = [ctypes.POINTER(DsmParams), ci, vp, ctypes.POINTER(vp)]
    L.dsm_destroy.argtypes = [vp]
    L.dsm_destroy.restype = None
    L.dsm_num_seeds.argtypes = [vp]
    L.dsm_fuse_frame.argtypes = [vp, ci, vp, cs, vp, cs, vp, vp, ci, vp, ci, ctypes.POINTER(ci)]
    L.dsm_batch_upload.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp]
    L.dsm_batch_run.argtypes = [vp]
    L.dsm_batch_download.argtypes = [vp, vp, vp, vp]
    L.dsm_sync.argtypes = [vp]
    L.dsm_fuse_batch.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dsm_fuse_batch_async.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dsm_batch_wait.argtypes = [vp]
    L.dsm_batch_restore_pool.argtypes = [vp]
    L.dsm_pool_upload.argtypes = [vp, vp, ci]
    L.dsm_fuse_frame_resident.argtypes = [vp, ci, vp, cs, vp, cs, vp, ctypes.POINTER(ci)]
    L.dsm_pool_transform.argtypes = [vp, vp]
    L.dsm_fuse_stream_resident.argtypes = [vp, ci, vp, vp, vp, vp, vp]
    L.dsm_inactive_reserve.argtypes = [vp, ci]
    L.dsm_inactive_retire.argtypes = [vp, ci, ctypes.POINTER(ci)]
    L.dsm_inactive_reactivate.argtypes = [vp, ci, ctypes.POINTER(ci)]
    L.dsm_inactive_transform.argtypes = [vp, ci, vp]
    L.dsm_inactive_export_cloud.argtypes = [vp, vp, ci, ctypes.POINTER(ci)]
    L.dsm_inactive_download.argtypes = [vp, ci, vp, ci, ctypes.POINTER(ci)]
    L.dsm_inactive_size.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci)]
    L.dsm_pool_retire.argtypes = [vp, ci, vp, ci, ctypes.POINTER(ci)]
    L.dsm_pool_append.argtypes = [vp, vp, ci]
    L.dsm_pool_size.argtypes = [vp, ctypes.POINTER(ci)]
    L.dsm_pool_download.argtypes = [vp, vp, ci, ctypes.POINTER(ci)]
    L.dsm_get_labels.argtypes = [vp, ci, vp]
    L.dsm_get_seeds.argtypes = [vp, ci, vp]
    L.dsm_debug_stop_after.argtypes = [vp, ci]
    L.dsm_debug_invariant_violations.argtypes = [vp, ctypes.POINTER(ci)]
    L.dsm_profile_enable.argtypes = [vp, ctypes.c_uint32]
    L.dsm_profile_reset.argtypes = [vp]
    L.dsm_profile_read.argtypes = [vp, vp, vp]
    L.dsm_device_buffer.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(cs)]
    L.dsm_pool_export_cloud.argtypes = [vp, ci, vp, ci, ctypes.POINTER(ci)]
    L.dsm_pool_export_surfels.argtypes = [vp, ci, vp, ci, ctypes.POINTER(ci)]
    L.dsm_write_pcd.argtypes = [ctypes.c_char_p, vp, cs, ci]
    L.dsm_write_ply_mesh.argtypes = [ctypes.c_char_p, vp, cs]
    L.dsm_mesh_vertices.argtypes = [vp, cs, vp]
    L.dsm_set_constants.argtypes = [vp, ctypes.POINTER(DsmConstants)]
    L.dsm_set_concurrency.argtypes = [vp, ctypes.c_int]
    L.dsm_comm_unique_id.argtypes = [vp]
    L.dsm_comm_init.argtypes = [vp, vp, ci, ci]
    L.dsm_comm_destroy.argtypes = [vp]
    L.dsm_gather_deltas.argtypes = [vp, ci]
    L.dsm_gather_wait.argtypes = [vp]
    L.dsm_gathered_device.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(cs)]
    L.dsm_gathered_rank_bytes.argtypes = [vp, ci, ctypes.POINTER(cs)]
    L.dsm_gathered_download.argtypes = [vp, ci, vp, cs]
    _lib = L
    return L


def comm_unique_id() -> bytes:
    """128-byte NCCL unique id (rank 0 creates it and hands it to the other ranks out of band)."""
    L = load_library()
    buf = ctypes.create_string_buffer(128)
    rc = L.dsm_comm_unique_id(buf)
    if rc != DSM_OK:
        raise DsmError(rc, L.dsm_strerror(rc).decode())
    return buf.raw


def kernel_names():
    L = load_library()
    return [L.dsm_kernel_name(i).decode() for i in range(NUM_KERNELS)]


def _ptr(a):
    return None if a is None else a.ctypes.data


def pose16(pose):
    """Eigen::Matrix4f memory order (column-major float32[16]) of a pose.  A flat 16-vector is taken as already
    column-major (what synth.pose_stream returns); a (4, 4) array is taken as the MATRIX T_world<-cam the way numpy
    writes it (row i, column j = pose[i, j]) and is transposed into column-major memory -- it is not reinterpreted."""
    a = np.asarray(pose, dtype=np.float32)
    if a.shape == (4, 4):
        return np.ascontiguousarray(a.T).reshape(16)
    if a.size != 16:
        raise ValueError("pose must be a (4, 4) matrix or a column-major 16-vector")
    return np.ascontiguousarray(a).reshape(16)


class Context:
    """Owns one dsm_ctx (one per GPU)."""

    def __init__(self, cam, max_batch=1, max_local_surfels=1 << 20, device=0, cuda_stream=None):
        self.lib = load_library()
        self.cam = cam
        self.S = num_seeds(cam.width, cam.height)
        self.max_batch = max_batch
        p = DsmParams(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.far, cam.near,
                      max_batch, max_local_surfels)
        h = ctypes.c_void_p()
        rc = self.lib.dsm_create(ctypes.byref(p), device, cuda_stream, ctypes.byref(h))
        if rc != DSM_OK:
            raise DsmError(rc, self.lib.dsm_strerror(rc).decode())
        self.h = h

    def _ck(self, rc):
        if rc != DSM_OK:
            raise DsmError(rc, self.lib.dsm_last_error(self.h).decode() or self.lib.dsm_strerror(rc).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.dsm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference-identical single frame ----
    def fuse_frame(self, ref_idx, gray, depth, pose, local):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        pose = pose16(pose)
        local = np.array(local, dtype=SURFEL_DTYPE, copy=True)
        new = np.zeros(self.S, dtype=SURFEL_DTYPE)
        n = ctypes.c_int(0)
        self._ck(self.lib.dsm_fuse_frame(self.h, int(ref_idx), _ptr(gray), gray.strides[0], _ptr(depth), depth.strides[0],
                                         _ptr(pose), _ptr(local) if len(local) else None, len(local),
                                         _ptr(new), self.S, ctypes.byref(n)))
        return local, new[:n.value].copy()

    # ---- batch stages ----
    def batch_upload(self, ref_idx, gray, depth, poses, local, offsets):
        self._keep = (np.ascontiguousarray(ref_idx, dtype=np.int32), np.ascontiguousarray(gray, dtype=np.uint8),
                      np.ascontiguousarray(depth, dtype=np.float32), np.ascontiguousarray(poses, dtype=np.float32),
                      np.ascontiguousarray(local, dtype=SURFEL_DTYPE), np.ascontiguousarray(offsets, dtype=np.int32))
        r, g, d, p, l, o = self._keep
        n = len(r)
        assert g.shape == (n, self.cam.height, self.cam.width) and d.shape == g.shape and len(o) == n + 1
        self.nb = n
        self.n_pool = int(o[-1])
        self._ck(self.lib.dsm_batch_upload(self.h, n, _ptr(r), _ptr(g), _ptr(d), _ptr(p), _ptr(l) if len(l) else None, _ptr(o)))

    def batch_run(self):
        self._ck(self.lib.dsm_batch_run(self.h))

    def batch_restore_pool(self):
        self._ck(self.lib.dsm_batch_restore_pool(self.h))

    def sync(self):
        self._ck(self.lib.dsm_sync(self.h))

    def batch_download(self):
        local = np.zeros(self.n_pool, dtype=SURFEL_DTYPE)
        new = np.zeros((self.nb, self.S), dtype=SURFEL_DTYPE)
        cnt = np.zeros(self.nb, dtype=np.int32)
        self._ck(self.lib.dsm_batch_download(self.h, _ptr(local) if self.n_pool else None, _ptr(new), _ptr(cnt)))
        self.sync()
        return local, [new[b, :cnt[b]].copy() for b in range(self.nb)]

    def fuse_batch(self, ref_idx, gray, depth, poses, local, offsets):
        self.batch_upload(ref_idx, gray, depth, poses, local, offsets)
        self.batch_run()
        return self.batch_download()

    def set_constants(self, constants):
        """(huber_range, baseline, disparity_error, min_tolerate_diff): CONSTANTS_DRIVE (default) or CONSTANTS_RGBD."""
        k = DsmConstants(*constants)
        self._ck(self.lib.dsm_set_constants(self.h, ctypes.byref(k)))

    def set_concurrency(self, sub_batches: int):
        """Concurrent sub-batches of batch_run (1..4, default 2); 1 for clean per-kernel profile durations."""
        self._ck(self.lib.dsm_set_concurrency(self.h, int(sub_batches)))

    # ---- multi-GPU gather of the surfel deltas (csrc/dsm_comm.cu) ----
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        assert len(unique_id) == 128
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.lib.dsm_comm_init(self.h, buf, int(rank), int(nranks)))
        self.comm_rank, self.comm_size = rank, nranks

    def gather_deltas(self, root=0):
        self._ck(self.lib.dsm_gather_deltas(self.h, int(root)))

    def gather_wait(self):
        self._ck(self.lib.dsm_gather_wait(self.h))

    def gathered_payload(self, rank):
        """root only: the payload rank `rank` sent in the last gather, as a uint8 array (see gather.unpack_payload)."""
        n = ctypes.c_size_t(0)
        self._ck(self.lib.dsm_gathered_rank_bytes(self.h, int(rank), ctypes.byref(n)))
        out = np.zeros(n.value, np.uint8)
        self._ck(self.lib.dsm_gathered_download(self.h, int(rank), _ptr(out), n.value))
        return out

    # ---- GPU-resident pool (stream mode) ----
    def pool_upload(self, local):
        local = np.ascontiguousarray(local, dtype=SURFEL_DTYPE)
        self._ck(self.lib.dsm_pool_upload(self.h, _ptr(local) if len(local) else None, len(local)))

    def fuse_frame_resident(self, ref_idx, gray, depth, pose, want_count=False):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        pose = pose16(pose)
        n = ctypes.c_int(0)
        self._ck(self.lib.dsm_fuse_frame_resident(self.h, int(ref_idx), _ptr(gray), gray.strides[0], _ptr(depth), depth.strides[0],
                                                  _ptr(pose), ctypes.byref(n) if want_count else None))
        return n.value if want_count else None

    def fuse_stream_resident(self, ref_idx, gray, depth, poses, want_counts=False):
        """n consecutive frames of one stream in one call (see include/dsm.h); gray [n,H,W] u8, depth [n,H,W] f32."""
        ref = np.ascontiguousarray(ref_idx, dtype=np.int32)
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(len(ref), 16)
        assert gray.shape == (len(ref), self.cam.height, self.cam.width) and depth.shape == gray.shape
        cnt = np.zeros(len(ref), dtype=np.int32)
        self._ck(self.lib.dsm_fuse_stream_resident(self.h, len(ref), _ptr(ref), _ptr(gray), _ptr(depth), _ptr(poses),
                                                   _ptr(cnt) if want_counts else None))
        return cnt if want_counts else None

    def pool_transform(self, W_colmajor):
        w = pose16(W_colmajor)
        self._ck(self.lib.dsm_pool_transform(self.h, _ptr(w)))

    def pool_retire(self, keyframe_index, cap=None):
        cap = self.pool_size() if cap is None else cap
        out = np.zeros(max(cap, 1), dtype=SURFEL_DTYPE)
        n = ctypes.c_int(0)
        self._ck(self.lib.dsm_pool_retire(self.h, int(keyframe_index), _ptr(out), cap, ctypes.byref(n)))
        return out[:min(n.value, cap)].copy()

    def pool_append(self, surfels):
        surfels = np.ascontiguousarray(surfels, dtype=SURFEL_DTYPE)
        self._ck(self.lib.dsm_pool_append(self.h, _ptr(surfels) if len(surfels) else None, len(surfels)))

    def pool_size(self):
        n = ctypes.c_int(0)
        self._ck(self.lib.dsm_pool_size(self.h, ctypes.byref(n)))
        return n.value

    def pool_download(self):
        n = self.pool_size()
        out = np.zeros(n, dtype=SURFEL_DTYPE)
        c = ctypes.c_int(0)
        self._ck(self.lib.dsm_pool_download(self.h, _ptr(out) if n else None, n, ctypes.byref(c)))
        return out[:c.value]

    def pool_export_cloud(self, min_update_times=5, cap=None):
        """publish_active_pointcloud & co. (surfel_map.cpp:1398-1417): POINT_DTYPE array in pool order."""
        return self._pool_export(self.lib.dsm_pool_export_cloud, POINT_DTYPE, min_update_times, cap)

    def pool_export_surfels(self, min_update_times=5, cap=None):
        return self._pool_export(self.lib.dsm_pool_export_surfels, SURFEL_DTYPE, min_update_times, cap)

    def _pool_export(self, fn, dtype, min_update_times, cap):
        cap = self.pool_size() if cap is None else cap
        out = np.zeros(max(cap, 1), dtype=dtype)
        n = ctypes.c_int(0)
        self._ck(fn(self.h, int(min_update_times), _ptr(out), cap, ctypes.byref(n)))
        return out[:min(n.value, cap)].copy()

    # ---- inactive store ----
    def inactive_reserve(self, n):
        self._ck(self.lib.dsm_inactive_reserve(self.h, int(n)))

    def inactive_retire(self, keyframe_index):
        n = ctypes.c_int(0)
        self._ck(self.lib.dsm_inactive_retire(self.h, int(keyframe_index), ctypes.byref(n)))
        return n.value

    def inactive_reactivate(self, keyframe_index):
        n = ctypes.c_int(0)
        self._ck(self.lib.dsm_inactive_reactivate(self.h, int(keyframe_index), ctypes.byref(n)))
        return n.value

    def inactive_transform(self, keyframe_index, W_colmajor):
        w = pose16(W_colmajor)
        self._ck(self.lib.dsm_inactive_transform(self.h, int(keyframe_index), _ptr(w)))

    def inactive_size(self):
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        self._ck(self.lib.dsm_inactive_size(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def inactive_download(self, keyframe_index=-1):
        cap = self.inactive_size()[0]
        out = np.zeros(max(cap, 1), dtype=SURFEL_DTYPE)
        n = ctypes.c_int(0)
        self._ck(self.lib.dsm_inactive_download(self.h, int(keyframe_index), _ptr(out), cap, ctypes.byref(n)))
        return out[:min(n.value, cap)].copy()

    def inactive_export_cloud(self):
        cap = self.inactive_size()[0]
        out = np.zeros(max(cap, 1), dtype=POINT_DTYPE)
        n = ctypes.c_int(0)
        self._ck(self.lib.dsm_inactive_export_cloud(self.h, _ptr(out), cap, ctypes.byref(n)))
        return out[:min(n.value, cap)].copy()

    # ---- parity readback ----
    def labels(self, frame=0):
        out = np.empty((self.cam.height, self.cam.width), dtype=np.int32)
        self._ck(self.lib.dsm_get_labels(self.h, frame, _ptr(out)))
        return out

    def seeds(self, frame=0):
        out = np.zeros(self.S, dtype=SEED_DTYPE)
        self._ck(self.lib.dsm_get_seeds(self.h, frame, _ptr(out)))
        return out

    def debug_stop_after(self, n):
        self._ck(self.lib.dsm_debug_stop_after(self.h, int(n)))

    def invariant_violations(self):
        c = ctypes.c_int(0)
        self._ck(self.lib.dsm_debug_invariant_violations(self.h, ctypes.byref(c)))
        return c.value

    # ---- measurement ----
    def profile_enable(self, mask):
        self._ck(self.lib.dsm_profile_enable(self.h, mask))

    def profile_reset(self):
        self._ck(self.lib.dsm_profile_reset(self.h))

    def profile_read(self):
        ms = np.zeros(NUM_KERNELS, dtype=np.float32)
        n = np.zeros(NUM_KERNELS, dtype=np.int32)
        self._ck(self.lib.dsm_profile_read(self.h, _ptr(ms), _ptr(n)))
        return ms, n

    def device_buffer(self, which):
        p = ctypes.c_void_p()
        sz = ctypes.c_size_t()
        self._ck(self.lib.dsm_device_buffer(self.h, which, ctypes.byref(p), ctypes.byref(sz)))
        return p.value, sz.value


class FusionFunctions:
    """Python mirror of the reference's `class FusionFunctions` public surface
    (fusion_functions.h:84-94): initialize(...) then fuse_initialize_map(...)."""

    def __init__(self, device=0, max_local_surfels=1 << 20):
        self._device = device
        self._cap = max_local_surfels
        self._ctx = None

    def initialize(self, width, height, fx, fy, cx, cy, fuse_far, fuse_near):
        from .synth import Camera
        self._ctx = Context(Camera(width, height, fx, fy, cx, cy, fuse_near, fuse_far),
                            max_batch=1, max_local_surfels=self._cap, device=self._device)

    def fuse_initialize_map(self, reference_frame_index, image, depth, pose, local_surfels):
        """Returns (local_surfels_updated, new_surfels) — the reference mutates the first in place
        and clears+fills the second (fusion_functions.cpp:30-83)."""
        if self._ctx is None:
            raise RuntimeError("initialize() has not been called")
        return self._ctx.fuse_frame(reference_frame_index, image, depth, pose, local_surfels)


# ---- output files (host only; no context, no GPU) ----
def _ck_io(rc, path):
    if rc != DSM_OK:
        raise DsmError(rc, f"{load_library().dsm_strerror(rc).decode()}: {path}")


def write_pcd(path, points, binary=False):
    """SurfelMap::save_cloud's file (surfel_map.cpp:1171): PCD v0.7 with fields x y z intensity."""
    points = np.ascontiguousarray(points, dtype=POINT_DTYPE)
    _ck_io(load_library().dsm_write_pcd(os.fsencode(path), _ptr(points) if len(points) else None, len(points), int(binary)), path)


def write_ply_mesh(path, surfels):
    """SurfelMap::save_mesh (surfel_map.cpp:1229-1280): one hexagon per surfel, ASCII PLY."""
    surfels = np.ascontiguousarray(surfels, dtype=SURFEL_DTYPE)
    _ck_io(load_library().dsm_write_ply_mesh(os.fsencode(path), _ptr(surfels) if len(surfels) else None, len(surfels)), path)


def mesh_vertices(surfels):
    """push_a_surfel (surfel_map.cpp:1175-1226): float32 [n, 6, 6] = 6 vertices x (x, y, z, c, c, c)."""
    surfels = np.ascontiguousarray(surfels, dtype=SURFEL_DTYPE)
    out = np.zeros((len(surfels), 6, 6), dtype=np.float32)
    _ck_io(load_library().dsm_mesh_vertices(_ptr(surfels) if len(surfels) else None, len(surfels), _ptr(out) if len(surfels) else None), "")
    return out
