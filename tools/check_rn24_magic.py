"""CPU check of the identity behind the experimental assign variant (k_assign_x): rounding a
double to float precision WITHOUT a conversion instruction,

    M = 1.5 * 2^(exponent(x) + 29)        (built from the high word of x with two integer ops)
    rn24(x) = (x + M) - M                 (two fp64 adds; the adder does the round-to-nearest-even)

equals (double)(float)x for every non-negative double whose float image is normal or zero -- including exact ties
and values one ulp either side of a tie.  Run: python tools/check_rn24_magic.py"""
import numpy as np


def rn24_magic(x):
    x = np.asarray(x, np.float64)
    hi = (x.view(np.uint64) >> np.uint64(32)).astype(np.uint32)
    mhi = ((hi & np.uint32(0x7FF00000)) + np.uint32(0x01D80000)).astype(np.uint64)
    M = (mhi << np.uint64(32)).view(np.float64)
    return (x + M) - M


def reference(x):
    return np.asarray(x, np.float64).astype(np.float32).astype(np.float64)


def check(x, what):
    got, want = rn24_magic(x), reference(x)
    bad = np.nonzero(got.view(np.uint64) != want.view(np.uint64))[0]
    assert len(bad) == 0, (what, x[bad[:5]], got[bad[:5]], want[bad[:5]])
    print(f"{what}: {len(x)} values ok")


if __name__ == "__main__":
    rng = np.random.RandomState(0)
    for rep in range(20):  # 1e8 log-uniform doubles over the cost range and far beyond
        e = rng.uniform(-100, 120, 5_000_000)
        m = rng.uniform(1.0, 2.0, 5_000_000)
        check(m * np.exp2(np.floor(e)), f"random batch {rep}")
    # exact ties (halfway between two floats), and their neighbours one double-ulp away
    f = np.abs(rng.standard_normal(2_000_000)).astype(np.float32) * np.float32(1000.0)
    f = f[f > 1e-30]
    up = np.nextafter(f, np.float32(np.inf))
    tie = (f.astype(np.float64) + up.astype(np.float64)) * 0.5
    check(tie, "ties")
    check(np.nextafter(tie, np.inf), "ties + 1 ulp")
    check(np.nextafter(tie, -np.inf), "ties - 1 ulp")
    check(f.astype(np.float64), "floats")
    check(np.array([0.0, 2.0 ** -126, 2.0 ** -100, 1.0, 1.9999999, 2.0, 16777216.0, 16777217.0, 3.0e38]), "specials")
    print("ok")
