"""Multi-GPU fused exchange+aggregate: `scripts/mp_check.py` under torchrun on 2 GPUs and on every GPU of the box
(skipped when the box has fewer; the single-GPU tier covers the same kernels through `test_engine_gpu.py`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "mp_check.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)


@pytest.mark.timeout(950)
def test_fused_exchange_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = _run(2, 29533)
    assert res.returncode == 0 and "ALL MULTI-GPU CHECKS PASSED" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.timeout(950)
def test_fused_exchange_all_gpus():
    n = torch.cuda.device_count()
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    res = _run(n, 29534)
    assert res.returncode == 0 and "ALL MULTI-GPU CHECKS PASSED" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
