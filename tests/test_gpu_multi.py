"""The multi-GPU RESULT on hardware (VERDICT r1 "missing" item 2): N NCCL ranks, one per GPU, each
rendering its voice range; the NCCL-summed mix equals the single-GPU mix within the mix tolerance and the shards' stems
are the single-GPU stems bit for bit.  Skipped when fewer than two GPUs are visible."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


# (cfg4 is silent until its 2 Hz gate opens at sample 12000)
@pytest.mark.parametrize("name,voices,samples", [("cfg4", 4096, 20000), ("cfg3b", 2048, 6000)])
def test_nccl_reduced_mix_equals_the_single_gpu_mix(name, voices, samples):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_worker.py"),
                        str(voices), str(samples), name], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    sys.stdout.write(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "bit-identical to the single-GPU render: True" in r.stdout
