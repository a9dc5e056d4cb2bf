import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_generate_tests(metafunc):
    """Every GPU test runs twice: with the schedule the engine picks on its own ("auto": pipelined warps for few
    voices, the fused kernel for many) and with the fused kernel forced for every launch ("fused")."""
    if "schedule" in metafunc.fixturenames:
        gpu = metafunc.definition.get_closest_marker("gpu") is not None
        metafunc.parametrize("schedule", ["auto", "fused"] if gpu else ["auto"], indirect=True)


@pytest.fixture(autouse=True)
def schedule(request, monkeypatch):
    mode = getattr(request, "param", "auto")
    if mode == "fused":
        monkeypatch.setenv("SRK_FUSED", "1")
    return mode


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure), built on demand with oracle/Makefile."""
    from oracle import orc as _orc

    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def srk():
    """The product package; importing it loads libsrack_b200.so (build it if it is missing)."""
    so = os.path.join(ROOT, "s-rack_b200", "libsrack_b200.so")
    if not os.path.exists(so):
        import __graft_entry__

        __graft_entry__.build()
    import srack_b200

    return srack_b200


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return 0
