#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs (tests/test_gpu_multi.py launches it):

    python -m torch.distributed.run --nproc-per-node N tests/mgpu_check.py lj|snap NX NY NZ STEPS [half|full]

Every rank runs its brick of the SAME global system through CommMPI (3-D decomposition, NCCL halo exchange);
rank 0 gathers the owned atoms of all ranks and compares x, v, f BY ATOM ID with the single-rank CPU oracle
(which is bit-identical to the reference): the decomposition must not change the physics beyond summation
order (1e-10 of the global RMS).  Also checks that every atom is owned exactly once and sits inside its brick."""
import os
import re
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
import examinimd_b200 as emd
from oracle_py import OracleMD

kind, nx, ny, nz, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
iteration = {"half": "NEIGH_HALF", "full": "NEIGH_FULL"}[sys.argv[6] if len(sys.argv) > 6 else "half"]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("gloo")

td = Path(tempfile.mkdtemp())
if kind == "lj":
    src = REPO / "input" / "in.lj"
else:
    src = REPO / "input" / "snap" / "in.snap.W"
    iteration = "NEIGH_FULL"
    for f in (REPO / "input" / "snap").glob("*.snap*"):
        (td / f.name).write_bytes(f.read_bytes())
txt = re.sub(r"region\s+box block.*", "region\t\tbox block 0 %d 0 %d 0 %d" % (nx, ny, nz), src.read_text())
txt = re.sub(r"run\s+\d+", "run\t\t%d" % steps, txt)
deck = td / "in.deck"
deck.write_text(txt)

app = emd.App(["-il", str(deck), "--neigh-type", "CSR", "--force-iteration", iteration, "--comm-type", "MPI"], device=local)
assert app.get("nranks") == world and app.get("rank") == rank
# atom ids are assigned brick by brick (src/input.cpp:578-584), so they label different atoms than in a single-rank
# run unless the cut happens to follow the lattice loop order: pair the two labelings through the step-0 lattice
# positions, which are bit-identical on both sides
start = app.download()
gathered0 = [None] * world
dist.gather_object({k: start[k] for k in ("id", "x")}, gathered0 if rank == 0 else None, dst=0)
app.advance(steps)
cur = app.download()
T, PE, KE = app.thermo()
gathered = [None] * world
dist.gather_object({k: cur[k] for k in ("id", "x", "v", "f")}, gathered if rank == 0 else None, dst=0)
ok = True
if rank == 0:
    ids = np.concatenate([g["id"] for g in gathered])
    x = np.concatenate([g["x"] for g in gathered]); v = np.concatenate([g["v"] for g in gathered]); f = np.concatenate([g["f"] for g in gathered])
    md = OracleMD.from_deck(deck, "CSR", iteration, coeff_dir=td if kind == "snap" else None)
    n = md.geti("N_local")
    assert ids.size == n and np.unique(ids).size == n, f"atoms owned: {ids.size} vs {n}"
    ids0 = np.concatenate([g["id"] for g in gathered0]); x0 = np.concatenate([g["x"] for g in gathered0])
    xo0, ido0 = md.arr("x")[:n], md.arr("id")[:n]
    ka, kb = np.lexsort(x0.T[::-1]), np.lexsort(xo0.T[::-1])
    assert np.array_equal(x0[ka], xo0[kb]), "step-0 lattice positions differ"
    to_oracle_id = np.zeros(ids0.max() + 1, np.int64)
    to_oracle_id[ids0[ka]] = ido0[kb]
    ids = to_oracle_id[ids]
    md.step(steps)
    o, r = np.argsort(ids), np.argsort(md.arr("id")[:n])
    xo, vo, fo = md.arr("x")[:n][r], md.arr("v")[:n][r], md.arr("f")[:n][r]
    L = np.array([md.getd("domain_x"), md.getd("domain_y"), md.getd("domain_z")])
    dx = x[o] - xo
    dx -= np.round(dx / L) * L
    ex = np.abs(dx).max() / np.sqrt((xo ** 2).mean())
    ev = np.abs(v[o] - vo).max() / np.sqrt((vo ** 2).mean())
    ef = np.abs(f[o] - fo).max() / max(np.sqrt((fo ** 2).mean()), 1e-3)
    To, PEo, KEo = md.thermo()
    print(f"MGPU {kind} {nx}x{ny}x{nz} ranks={world} steps={steps} {iteration}: atoms={n} per-rank={[g['id'].size for g in gathered]} "
          f"err x={ex:.2e} v={ev:.2e} f={ef:.2e} T={T:.9f}/{To:.9f} PE={PE:.9f}/{PEo:.9f}", flush=True)
    ok = ex < 1e-10 and ev < 1e-10 and ef < 1e-10 and abs(T - To) < 1e-9 * max(To, 1) and abs(PE - PEo) < 1e-9 * max(abs(PEo), 1)
    print("MGPU_OK" if ok else "MGPU_FAIL", flush=True)
app.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
