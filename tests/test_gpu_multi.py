"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): z-slab distributed solve == single-GPU solve,
and the owned row blocks concatenate to the exact global CSR.  The host-side partition logic is covered
on CPU by tests/test_host_logic.py::test_world_size_two_gloo_partition."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_slab_solve_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "tools", "dist_check.py"), "20"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_two_rank_general_partition_matches_single_gpu():
    """Unstructured mesh, degree-2 elasticity and a transient run on the general node partition (RCB + send-list halo)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29543", os.path.join(ROOT, "tools", "dist_check_general.py"), "10"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "DIST_GENERAL_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
