"""CPU tests of the multi-GPU host logic (no GPU): CommMPI::create_domain_decomposition through the C ABI
(emd_comm_decompose, src/comm_types/comm_mpi.cpp:52-147) and the N>1 plumbing of bench.py, with
world_size-2 gloo process groups."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle_py import REPO


def decompose(emd, nranks, rank, domain):
    d = emd.Decomp()
    emd.check(emd.lib().emd_comm_decompose(nranks, rank, emd.vec3(domain), C.byref(d)))
    return d


@pytest.mark.parametrize("nranks,domain,grid", [
    (1, (10, 10, 10), (1, 1, 1)),
    (2, (10, 10, 10), (1, 1, 2)),     # cubic box: first minimal-surface grid found wins (strict <), SURVEY 8(e)
    (4, (10, 10, 10), (1, 2, 2)),
    (8, (10, 10, 10), (2, 2, 2)),
    (2, (40, 10, 10), (2, 1, 1)),     # long box: cut the long axis
    (6, (30, 20, 10), (3, 2, 1)),
    (8, (10, 10, 80), (1, 1, 8)),
])
def test_processor_grid_and_bricks(emd, nranks, domain, grid):
    seen = set()
    vol = 0.0
    for rank in range(nranks):
        d = decompose(emd, nranks, rank, domain)
        assert tuple(d.grid) == grid
        pos = tuple(d.pos)
        assert rank == pos[0] + grid[0] * (pos[1] + grid[1] * pos[2])  # comm_mpi.cpp:90-92
        seen.add(pos)
        for k in range(3):
            assert d.sub[k] == domain[k] / grid[k]
            assert d.sub_lo[k] == pos[k] * d.sub[k] and d.sub_hi[k] == (pos[k] + 1) * d.sub[k]
        vol += d.sub[0] * d.sub[1] * d.sub[2]
        for phase in range(6):
            dim = phase // 2
            if grid[dim] == 1:
                assert d.neighbor_send[phase] == -1 and d.neighbor_recv[phase] == -1
                continue
            # the neighbor one step up (even phase) / down (odd phase), periodic; recv is the opposite side
            step = 1 if phase % 2 == 0 else -1
            npos = list(pos)
            npos[dim] = (pos[dim] + step) % grid[dim]
            want = npos[0] + grid[0] * (npos[1] + grid[1] * npos[2])
            assert d.neighbor_send[phase] == want
            assert d.neighbor_recv[phase] == d.neighbor_send[phase ^ 1]
            back = decompose(emd, nranks, want, domain)
            assert back.neighbor_recv[phase] == rank  # reciprocity: whoever I send to in a phase receives from me
    assert len(seen) == nranks
    assert abs(vol - domain[0] * domain[1] * domain[2]) < 1e-9 * vol


WORKER = r"""
import ctypes as C, os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import torch, torch.distributed as dist
import examinimd_b200 as emd
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
d = emd.Decomp()
emd.check(emd.lib().emd_comm_decompose(world, rank, emd.vec3((20.0, 10.0, 10.0)), C.byref(d)))
mine = torch.tensor([d.rank, d.neighbor_send[0], d.neighbor_recv[0], d.grid[0]], dtype=torch.int64)
allv = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
dist.all_gather(allv, mine)
for r, v in enumerate(allv):
    assert int(v[0]) == r and int(v[3]) == world
    assert int(allv[int(v[1])][2]) == r      # my +x send peer lists me as its +x recv peer
lo = torch.tensor([d.sub_lo[0], d.sub_hi[0]], dtype=torch.float64)
ext = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
dist.all_gather(ext, lo)
assert float(ext[0][0]) == 0.0 and float(ext[-1][1]) == 20.0 and all(float(ext[k][1]) == float(ext[k + 1][0]) for k in range(world - 1))
dist.barrier()
if rank == 0:
    print("DECOMP_OK", world)
dist.destroy_process_group()
"""


def _torchrun(args, timeout=240):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", "29631", *args], capture_output=True, text=True, env=env, timeout=timeout, cwd=REPO)


def test_two_rank_decomposition_is_consistent_gloo(emd, tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    r = _torchrun([str(w), str(REPO)])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "DECOMP_OK 2" in r.stdout


def test_bench_reference_arm_two_ranks_prints_one_line(emd):
    """bench.py --impl reference under a 2-rank launch: rank 0 alone runs the CPU reference and prints ONE JSON line"""
    r = _torchrun(["bench.py", "--gpus", "2", "--impl", "reference", "--steps", "2", "--warmup", "1", "--region", "10", "10", "10"])
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["value"] > 0 and j["e2e"]["h2d_bytes_per_step"] == 0
