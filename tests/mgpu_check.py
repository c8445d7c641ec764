#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs (tests/test_gpu_multi.py launches it; bench.py --gpus N
calls parity() before its timed region):

    python -m torch.distributed.run --nproc-per-node N tests/mgpu_check.py lj|snap NX NY NZ STEPS [half|full]

Every rank runs its brick of the SAME global system through CommMPI (3-D decomposition, halo exchange over NVLink);
rank 0 gathers the owned atoms of all ranks and compares x, v, f BY ATOM ID with the single-rank CPU oracle
(which is bit-identical to the reference): the decomposition must not change the physics beyond summation
order (1e-10 of the global RMS).  Also checks that every atom is owned exactly once."""
import os
import re
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))


def parity(kind, nx, ny, nz, steps, iteration="NEIGH_HALF", group=None, device=None):
    """returns (on rank 0) {"case", "atoms", "err_x", "err_v", "err_f", "ok"}; None on the other ranks.  `group`: a
    torch.distributed group whose backend can move python objects (gloo)."""
    import torch
    import torch.distributed as dist
    import examinimd_b200 as emd
    from oracle_py import OracleMD

    rank, world = dist.get_rank(), dist.get_world_size()
    local = device if device is not None else int(os.environ.get("LOCAL_RANK", "0"))
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
    dist.gather_object({k: start[k] for k in ("id", "x")}, gathered0 if rank == 0 else None, dst=0, group=group)
    if os.environ.get("EMD_MGPU_USE_RUN"):  # the reference's loop (thermo at the deck's cadence: thermo steps served by the fused force launch)
        app.run(steps)
    else:
        app.advance(steps)
    cur = app.download()
    T, PE, KE = app.thermo()
    gathered = [None] * world
    dist.gather_object({k: cur[k] for k in ("id", "x", "v", "f")}, gathered if rank == 0 else None, dst=0, group=group)
    app.close()
    if rank != 0:
        return None
    ids = np.concatenate([g["id"] for g in gathered])
    x = np.concatenate([g["x"] for g in gathered]); v = np.concatenate([g["v"] for g in gathered]); f = np.concatenate([g["f"] for g in gathered])
    md = OracleMD.from_deck(deck, "CSR", iteration, coeff_dir=td if kind == "snap" else None)
    n = md.geti("N_local")
    case = f"{kind} {nx}x{ny}x{nz} ranks={world} steps={steps} {iteration}"
    if ids.size != n or np.unique(ids).size != n:
        return {"case": case, "atoms": int(n), "ok": False, "why": f"atoms owned: {ids.size} vs {n}"}
    ids0 = np.concatenate([g["id"] for g in gathered0]); x0 = np.concatenate([g["x"] for g in gathered0])
    xo0, ido0 = md.arr("x")[:n], md.arr("id")[:n]
    ka, kb = np.lexsort(x0.T[::-1]), np.lexsort(xo0.T[::-1])
    if not np.array_equal(x0[ka], xo0[kb]):
        return {"case": case, "atoms": int(n), "ok": False, "why": "step-0 lattice positions differ"}
    to_oracle_id = np.zeros(ids0.max() + 1, np.int64)
    to_oracle_id[ids0[ka]] = ido0[kb]
    ids = to_oracle_id[ids]
    md.step(steps)
    o, r = np.argsort(ids), np.argsort(md.arr("id")[:n])
    xo, vo, fo = md.arr("x")[:n][r], md.arr("v")[:n][r], md.arr("f")[:n][r]
    L = np.array([md.getd("domain_x"), md.getd("domain_y"), md.getd("domain_z")])
    dx = x[o] - xo
    dx -= np.round(dx / L) * L
    ex = float(np.abs(dx).max() / np.sqrt((xo ** 2).mean()))
    ev = float(np.abs(v[o] - vo).max() / np.sqrt((vo ** 2).mean()))
    ef = float(np.abs(f[o] - fo).max() / max(np.sqrt((fo ** 2).mean()), 1e-3))
    To, PEo, KEo = md.thermo()
    md.close()
    ok = bool(ex < 1e-10 and ev < 1e-10 and ef < 1e-10 and abs(T - To) < 1e-9 * max(To, 1) and abs(PE - PEo) < 1e-9 * max(abs(PEo), 1))
    return {"case": case, "atoms": int(n), "per_rank": [int(g["id"].size) for g in gathered], "err_x": ex, "err_v": ev, "err_f": ef,
            "T": [T, To], "PE": [PE, PEo], "ok": ok}


def main():
    import torch
    import torch.distributed as dist
    kind, nx, ny, nz, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    iteration = {"half": "NEIGH_HALF", "full": "NEIGH_FULL"}[sys.argv[6] if len(sys.argv) > 6 else "half"]
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("gloo")
    res = parity(kind, nx, ny, nz, steps, iteration)
    ok = True
    if dist.get_rank() == 0:
        ok = res["ok"]
        line = (f"MGPU {res['case']}: atoms={res['atoms']} per-rank={res.get('per_rank')} err x={res.get('err_x', float('nan')):.2e} "
                f"v={res.get('err_v', float('nan')):.2e} f={res.get('err_f', float('nan')):.2e} T={res.get('T')} PE={res.get('PE')} {res.get('why', '')}")
        print(line, flush=True)
        print("MGPU_OK" if ok else "MGPU_FAIL", flush=True)
        out = REPO / "gpurun_out"
        if out.is_dir():  # evidence for profiles/: kept when the run happens under gpurun
            with open(out / "mgpu_parity.log", "a") as fh:
                fh.write(line + (" MGPU_OK\n" if ok else " MGPU_FAIL\n"))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
