"""CPU checks of the bench harness: every kernel of the schedule has an algorithmic-byte model (the roofline
numerator) and the CLI keeps the driver's contract flags."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_every_scheduled_kernel_has_a_byte_model():
    import bench
    from densesurfelmapping_b200 import capi
    P, S = 1226 * 370, 7038
    names = capi.kernel_names()
    assert len(names) == capi.NUM_KERNELS
    for k in names:
        if k != "repack":  # input staging of the host-buffer entry points, not part of the resident step
            assert bench.kernel_alg_bytes(k, P, S, 6000, 2600) > 0, k
    # compulsory bytes of the whole path per frame
    assert abs((9 * P + 60 * S) - 4_504_860) < 10


def test_cli_contract_flags():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert r.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in r.stdout
