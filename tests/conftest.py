import os
import sys

import os

import pytest

# tests/test_exchange_gpu.py runs several exchange ranks on ONE device: each rank's kernels spin on flags the other ranks'
# kernels set, so their streams must not share a hardware queue (8 by default); ranks on their own GPUs never do
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
