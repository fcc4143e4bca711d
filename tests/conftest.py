"""Test configuration.

Markers / switches:
  -m "not gpu"             CPU suite: oracle vs golden vectors, ABI surface, host logic, gloo gather, file writers
  -m gpu                   parity suite proper, through the C ABI on a B200
  DSM_TEST_VARIANTS=1      also run tests/test_gpu_variants.py (experimental kernel variants)
  DSM_TEST_UNVERIFIED=1    also run the GPU tests of code written after round 1's GPU budget was spent
                           (chunked stream, C++ resident-pool helper)
  DSM_EXPERIMENTAL_VARIANTS=<mask>   library switch: run ANY test / bench.py with the given variant mask as default
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
