"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): N ranks over NCCL give exactly the single-GPU
result — counters, number of rows, every (key, count) row and the enrichment marginals — both for the one-exchange UMI de-duplication (DEL, lineage)
and for the final table merge (CRISPR)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["del3", "crispr", "lineage"])
def test_n_ranks_equal_one_gpu(name):
    n = n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29610 + os.getpid() % 300), os.path.join(ROOT, "tests", "multi_gpu_check.py"), name, "300000", "70000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_CHECK" in r.stdout and " OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
