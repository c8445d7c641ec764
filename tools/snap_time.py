#!/usr/bin/env python
"""Times the SNAP path on the GPU box: ExaMiniMD in.snap.W at a given region, per-phase device timers.
    python tools/snap_time.py [nx ny nz] [steps]"""
import re
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import examinimd_b200 as emd

region = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (50, 50, 100)
steps = int(sys.argv[4]) if len(sys.argv) >= 5 else 10
snap = REPO / "input" / "snap"
with tempfile.TemporaryDirectory() as td:
    td = Path(td)
    txt = (snap / "in.snap.W").read_text()
    txt = re.sub(r"region\s+box block.*", "region\t\tbox block 0 %d 0 %d 0 %d" % region, txt)
    (td / "in.deck").write_text(txt)
    for f in snap.glob("*.snap*"):
        (td / f.name).write_bytes(f.read_bytes())
    t0 = time.time()
    app = emd.App(["-il", str(td / "in.deck"), "--neigh-type", "CSR", "--comm-type", "SERIAL"])
    n = app.get("N")
    print(f"atoms {n}  init {time.time() - t0:.2f} s  neighs {app.get('total_neighs')}")
    app.advance(2)
    ph = app.advance_timed(steps)
    tot = sum(ph.values())
    print({k: round(1e3 * v / steps, 3) for k, v in ph.items()}, "ms/step;", f"{n * steps / tot:.4g} atom-steps/s")
    app.close()
