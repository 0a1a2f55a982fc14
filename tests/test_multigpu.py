"""Multi-GPU parity (-m gpu; skipped on a box with fewer than 2 GPUs): torchrun launches tests/multigpu_worker.py with one
rank per GPU; the N-GPU fit of the cell-sharded matrix is compared with the oracle and with the 1-GPU fit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_sharded_fit_matches_oracle_and_single_gpu(nranks):
    if _gpu_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    port = 29600 + nranks
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    assert "FAIL" not in out.stdout and out.stdout.count("PASS") >= 11, out.stdout[-3000:]
