"""GPU, needs >= 2 devices: partitions on separate GPUs, halo of fct_plus / fct_minus over NVLink
(NCCL send/recv inside fct_ale_step_), overlapped with interior work.  Skipped on a 1-GPU box; the
same partition logic is covered there by test_partitioned_on_one_gpu and on CPU by the gloo test."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("name", ["pi", "core2", "delaunay"])
def test_multi_gpu_step_matches_single_domain_oracle(tmp_path, name):
    n = min(gpu_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    port = 29600 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "nccl_worker.py"), str(tmp_path), name]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for k in range(n):
        assert (tmp_path / f"ok_{k}").exists()
