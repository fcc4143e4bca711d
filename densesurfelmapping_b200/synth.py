"""Deterministic synthetic depth+gray frames for parity tests and the bench.

There is no dataset access, so every measured or tested frame comes from here.  The wire
format is the one the reference consumes at its boundary (kitti_publisher/scripts/publisher.py:
mono8 gray + 32FC1 metric depth, 0 = invalid).

Scene (world frame = first camera frame, x right, y down, z forward): a ground plane 1.65 m
below the camera, two slightly slanted side walls, boxes standing on the ground every 8 m
(depth discontinuities), and a camera-relative backdrop at 40 m so that every pixel has depth.
Rendered by ray casting from the given pose, then multiplicative Gaussian depth noise
(sigma 0.2 %), ~1 % holes on the lattice ((7u+13v) mod 97)==0 and one >=16x16 hole block
(exercises the seed-initialisation search and the no-depth cost path).  Gray is
128+60 sin(0.05u) cos(0.07v)+N(0,8) clamped to u8.  Seeded with RandomState(1234+frame_id).
"""
from dataclasses import dataclass
import numpy as np


@dataclass(frozen=True)
class Camera:
    width: int
    height: int
    fx: float
    fy: float
    cx: float
    cy: float
    near: float
    far: float


# reference: ORB_SLAM2/Examples/Stereo/KITTI04-12.yaml:8-25, surfel_fusion/launch/kitti_orb.launch:15-16
KITTI = Camera(1226, 370, 707.0912, 707.0912, 601.8873, 183.1104, 0.5, 30.0)
# reference: ORB_SLAM2/Examples/Stereo/KITTI00-02.yaml:8-25, kitti_orb.launch:5-12
KITTI00 = Camera(1241, 376, 718.856, 718.856, 607.1928, 185.2157, 0.5, 30.0)
# TUM-like (ORB_SLAM2/Examples/RGB-D/TUM1.yaml:8-20) but driven with the "drive" constant set
VGA = Camera(640, 480, 525.0, 525.0, 319.5, 239.5, 0.5, 30.0)
HD = Camera(1280, 720, 720.0, 720.0, 639.5, 359.5, 0.5, 30.0)
CAMERAS = {"kitti": KITTI, "kitti00": KITTI00, "vga": VGA, "hd": HD}


def pose_stream(t: int, step_m: float = 0.8, yaw_deg: float = 0.5) -> np.ndarray:
    """T_world<-cam for frame t of the synthetic drive: forward 0.8 m/frame, 0.5 deg yaw/frame.
    Returned column-major flattened float32[16] (Eigen::Matrix4f memory order)."""
    a = np.deg2rad(yaw_deg) * t
    c, s = np.cos(a), np.sin(a)
    T = np.eye(4, dtype=np.float64)
    T[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    T[:3, 3] = [0.15 * t * np.sin(0.1 * t), 0.0, step_m * t]
    return np.ascontiguousarray(T.T.astype(np.float32).reshape(16))  # column-major


def identity_pose() -> np.ndarray:
    return np.eye(4, dtype=np.float32).reshape(16).copy()


def make_frame(cam: Camera, frame_id: int = 0, pose_colmajor=None, flat: bool = False):
    """Returns (gray uint8[H,W], depth float32[H,W]).  ``flat``: noise-free, piecewise-constant
    gray (the hard case for summation-order sensitivity)."""
    W, H = cam.width, cam.height
    rng = np.random.RandomState(1234 + frame_id)
    if pose_colmajor is None:
        pose_colmajor = identity_pose()
    T = np.asarray(pose_colmajor, dtype=np.float64).reshape(4, 4).T
    R, t = T[:3, :3], T[:3, 3]
    u = np.arange(W, dtype=np.float64)[None, :]
    v = np.arange(H, dtype=np.float64)[:, None]
    dc = np.stack(np.broadcast_arrays((u - cam.cx) / cam.fx, (v - cam.cy) / cam.fy, np.ones((H, W))), -1)
    dw = dc @ R.T  # world ray directions, camera-frame z component of the hit == ray parameter

    def plane(nrm, off):
        nrm = np.asarray(nrm, dtype=np.float64)
        den = dw @ nrm
        num = off - float(nrm @ t)
        with np.errstate(divide="ignore", invalid="ignore"):
            s = num / den
        return np.where((s > 0.05) & np.isfinite(s), s, np.inf)

    depth = np.full((H, W), 40.0)                      # camera-relative backdrop
    depth = np.minimum(depth, plane([0, 1, 0], 1.65))  # ground
    depth = np.minimum(depth, plane([1, 0, 0.02], 7.0))   # right wall
    depth = np.minimum(depth, plane([-1, 0, 0.015], 6.5))  # left wall
    # boxes on the ground every 8 m: front faces z = 8k+5, |x-2.5| < 1, y in [0.65, 1.65]
    k0 = int(np.floor(t[2] / 8.0))
    for k in range(k0, k0 + 5):
        zf = 8.0 * k + 5.0
        s = plane([0, 0, 1], zf)
        with np.errstate(invalid="ignore"):
            hit = t[None, None, :] + dw * np.where(np.isfinite(s), s, 0.0)[..., None]
        inside = np.isfinite(s) & (np.abs(hit[..., 0] - 2.5) < 1.0) & (hit[..., 1] > 0.65) & (hit[..., 1] < 1.65)
        depth = np.where(inside, np.minimum(depth, s), depth)
    if not flat:
        depth = depth * (1.0 + 0.002 * rng.standard_normal((H, W)))
    ui = np.arange(W)[None, :]
    vi = np.arange(H)[:, None]
    depth = np.where(((7 * ui + 13 * vi) % 97) == 0, 0.0, depth)
    hb_u, hb_v = int(0.3 * W), int(0.55 * H)
    depth[hb_v:hb_v + 22, hb_u:hb_u + 30] = 0.0       # hole block >= 16x16
    depth = depth.astype(np.float32)

    if flat:
        gray = 96.0 + 64.0 * (((ui // 40) + (vi // 40)) % 2)
    else:
        gray = 128.0 + 60.0 * np.sin(0.05 * ui) * np.cos(0.07 * vi) + 8.0 * rng.standard_normal((H, W))
    gray = np.clip(np.rint(gray), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(gray), np.ascontiguousarray(depth)
