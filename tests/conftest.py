"""Test configuration.

Markers / switches:
  -m "not gpu"             CPU suite: oracle vs golden vectors, ABI surface, host logic, gloo gather, file writers
  -m gpu                   parity suite proper, through the C ABI on a B200
No test is gated by an environment variable; the only hardware-conditional skip is the two-rank NCCL test, which
needs two GPUs (tests/test_gpu_comm.py, run with `gpurun --gpus 2`).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
