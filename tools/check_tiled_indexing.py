"""Desk check (CPU, numpy) of the index arithmetic of the experimental tiled gathers (k_gather_depths_tiled /
k_gather_points_tiled, csrc/dsm_kernels.cu): models the block's bulk copies into the shared-memory tile and
the per-lane tile reads, and checks (a) every cp.async.bulk has 16-byte aligned source, destination and size,
(b) every lane that the kernel does not mask reads exactly the value the direct-load kernel reads from global
memory.  Run: python tools/check_tiled_indexing.py"""
import numpy as np

GT_STRIDE, GT_GSTRIDE = 80, 144


def check(W, H, points=False):
    Wp = (W + 15) // 16 * 16
    spw, sph = W // 8, H // 8
    rng = np.random.RandomState(W * 7 + H)
    lab = rng.randint(0, 1 << 30, (H, Wp)).astype(np.int64)
    gry = rng.randint(0, 256, (H, Wp)).astype(np.int64)
    n_checked = 0
    for by in range(sph):
        for bx in range((spw + 7) // 8):
            X0, Y0 = bx * 64 - 4, by * 8 - 4
            t_lab = np.full(16 * GT_STRIDE, -7, np.int64)       # -7 = never written
            t_gry = np.full(16 * GT_GSTRIDE, -7, np.int64)
            ya, yz = max(Y0, 0), min(Y0 + 16, H)
            xs, xt = max(X0, 0), min(X0 + 72, Wp)
            gs, gt = max(X0 - 12, 0), min(X0 + 84, Wp)
            nb4, nbg = (xt - xs) * 4, gt - gs
            assert nb4 > 0 and nbg > 0 and nb4 % 16 == 0 and nbg % 16 == 0
            assert yz > ya
            for lane in range(16):
                y = Y0 + lane
                if not (ya <= y < yz):
                    continue
                dst = lane * GT_STRIDE + (xs - X0)
                assert (dst * 4) % 16 == 0 and ((y * Wp + xs) * 4) % 16 == 0
                assert dst + (xt - xs) <= (lane + 1) * GT_STRIDE
                t_lab[dst:dst + xt - xs] = lab[y, xs:xt]
                gd = lane * GT_GSTRIDE + (gs - (X0 - 12))
                assert gd % 16 == 0 and (y * Wp + gs) % 16 == 0
                assert gd + nbg <= (lane + 1) * GT_GSTRIDE
                t_gry[gd:gd + nbg] = gry[y, gs:gt]
            for warp in range(8):
                sp_x = bx * 8 + warp
                if sp_x >= spw:
                    continue
                x0, y0 = sp_x * 8 - 4, by * 8 - 4
                if points:   # k_gather_points bounds: rows [0, H), columns < Wp (members are further masked by x < W)
                    yb, ye = 0, H
                else:        # k_gather_depths bounds (:488-489)
                    yb, ye = max(y0, 0), min(y0 + 16, H - 1)
                for lane in range(32):
                    xq = x0 + 4 * (lane & 3)
                    colin = 0 <= xq < Wp
                    for ps in range(2):
                        r = 8 * ps + (lane >> 2)
                        y = y0 + r
                        if not (colin and yb <= y < ye):
                            continue  # masked lane: value irrelevant
                        c = 8 * warp + 4 * (lane & 3)
                        assert (r * GT_STRIDE + c) % 4 == 0
                        got = t_lab[r * GT_STRIDE + c: r * GT_STRIDE + c + 4]
                        assert (got == lab[y, xq:xq + 4]).all(), (W, H, bx, by, warp, lane, ps)
                        gg = t_gry[r * GT_GSTRIDE + c + 12: r * GT_GSTRIDE + c + 16]
                        assert (gg == gry[y, xq:xq + 4]).all(), ("gray", W, H, bx, by, warp, lane, ps)
                        n_checked += 1
    return n_checked


def bank_check():
    """LDS.128 is served per quarter-warp (8 lanes): their 16-byte accesses must cover 32 distinct banks."""
    for warp in range(8):
        for ps in range(2):
            for quarter in range(4):
                banks = set()
                for lane in range(quarter * 8, quarter * 8 + 8):
                    w0 = (8 * ps + (lane >> 2)) * GT_STRIDE + 8 * warp + 4 * (lane & 3)
                    banks |= {(w0 + k) % 32 for k in range(4)}
                assert len(banks) == 32
            banks = set()
            for lane in range(32):   # uchar4 (LDS.32): the whole warp in one go
                a = (8 * ps + (lane >> 2)) * GT_GSTRIDE + 8 * warp + 4 * (lane & 3) + 12
                assert a % 4 == 0
                banks.add((a // 4) % 32)
            assert len(banks) == 32


def compaction_tile_check(cap=228):
    """[seed][k] list tile of the tiled gathers: the copy-out (thread t reads tile[(t & 7) * cap + (t >> 3) + 32 j]) must
    hit 32 distinct banks per warp, and a compaction store whose lanes write consecutive positions is conflict-free."""
    for j in range(8):
        for w in range(8):
            banks = {(((t & 7) * cap + (t >> 3) + 32 * j) % 32) for t in range(32 * w, 32 * w + 32)}
            assert len(banks) == 32
    for w in range(8):
        for start in (0, 5, 31, 100):
            assert len({(w * cap + start + l) % 32 for l in range(32)}) == 32


if __name__ == "__main__":
    bank_check()
    compaction_tile_check()
    for (W, H) in [(1226, 370), (1241, 376), (640, 480), (1280, 720), (64, 48), (36, 28), (24, 24), (68, 44), (132, 100)]:
        for pts in (False, True):
            print(W, H, "points" if pts else "depths", check(W, H, pts), "lane reads verified")
    print("ok")
