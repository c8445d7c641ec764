#!/usr/bin/env python
"""Per-step device time of the LJ 2M-atom deck around the first re-neighborings after the lattice start (what a short bench
window sees against the steady state)."""
import os, sys, time
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import examinimd_b200 as emd
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
grid = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[world]
app = emd.App(["-il", str(REPO / "input" / "in.lj"), "--neigh-type", "CSR", "--force-iteration", "NEIGH_HALF", "--comm-type", "MPI" if world > 1 else "SERIAL",
               "--region", *[str(80 * g) for g in grid]], device=local)
app.run(5); app.thermo(); app.run(14)
for s in range(20, 66):
    t0 = time.time()
    ph = app.advance_timed(1)
    app.sync()
    wall = (time.time() - t0) * 1e3
    tot = sum(ph.values()) * 1e3
    if rank == 0 and (tot > 0.7 or s % 20 in (0, 1)):
        print(f"step {s}: device {tot:.3f} ms wall {wall:.3f} ms", {k: round(v * 1e3, 3) for k, v in ph.items()})
    if s % 10 == 0:
        t0 = time.time(); app.thermo()
        if rank == 0: print(f"   thermo at {s}: wall {(time.time()-t0)*1e3:.3f} ms")
