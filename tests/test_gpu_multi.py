"""Multi-GPU parity (-m gpu, needs >= 2 GPUs; skipped on a 1-GPU box): CommMPI's brick decomposition + NCCL halo
exchange against the single-rank oracle, by atom id (tests/mgpu_check.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

from oracle_py import REPO


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def run(n, *args, port=29701, extra_env=None):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1", **(extra_env or {}))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), "tests/mgpu_check.py", *map(str, args)], capture_output=True, text=True, env=env, cwd=REPO,
                       timeout=600)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("n,args", [(2, ("lj", 12, 12, 14, 45, "half")), (2, ("lj", 12, 10, 12, 25, "full")), (2, ("snap", 4, 4, 8, 6)),
                                    (4, ("lj", 12, 14, 14, 45, "half")), (8, ("lj", 14, 14, 14, 45, "half")), (8, ("snap", 8, 8, 8, 4))])
def test_decomposed_run_matches_single_rank_oracle(emd, oracle_lib, n, args):
    if _ngpus() < n:
        pytest.skip(f"needs {n} GPUs")
    run(n, *args, port=29701 + n)


@pytest.mark.parametrize("env", [{"EMD_HALO_GATE": "1"}, {"EMD_HALO_TRANSPORT": "nccl"}, {"EMD_NO_OVERLAP": "1"}, {"EMD_MGPU_USE_RUN": "1"}])
def test_decomposed_run_other_halo_schedules(emd, oracle_lib, env):
    """the same parity bar for the schedules that are not the default: the force kernel waiting for the neighbours' arrival flags
    itself (halo gate), NCCL send/recv groups instead of peer stores, the blocking (not overlapped) refresh, and the reference's
    run loop (thermo steps served by the fused force + integrator launch of a decomposed run)"""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    run(2, "lj", 12, 14, 14, 45, "half", port=29731, extra_env=env)
