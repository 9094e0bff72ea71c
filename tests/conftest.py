import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~30 s on CPU")


@pytest.fixture(autouse=True)
def _single_device_reference(request, monkeypatch):
    """`ThinCurr.compute_Lmat()` spreads over every visible device (symmetric shards, whose tiles sum in another order than
    the single-device build).  The parity tests compare bits against the single-device matrix, so they pin it to one
    device; tests/test_gpu_multi.py exercises the multi-device paths and manages the variable itself."""
    if 'test_gpu_multi' not in request.node.nodeid:
        monkeypatch.setenv('THINCURR_B200_NDEV', '1')
    else:
        monkeypatch.delenv('THINCURR_B200_NDEV', raising=False)
