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
    """Every GPU test runs three times: with the schedule the engine picks on its own ("auto": the fused kernel, cut
    into pipeline stages when an SM holds few voice groups), with the fused kernel as one warp per voice group
    ("fused1") and with the interpreter kernels ("interp": what runs when NVRTC is missing)."""
    if "schedule" in metafunc.fixturenames:
        gpu = metafunc.definition.get_closest_marker("gpu") is not None
        metafunc.parametrize("schedule", ["auto", "fused1", "interp"] if gpu else ["auto"], indirect=True)


@pytest.fixture(autouse=True)
def schedule(request, monkeypatch):
    mode = getattr(request, "param", "auto")
    if mode == "fused1":
        monkeypatch.setenv("SRK_FUSED", "1")
        monkeypatch.setenv("SRK_FUSED_STAGES", "1")
    elif mode == "interp":
        monkeypatch.setenv("SRK_FUSED", "0")
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
