"""GPU parity against the REFERENCE's own dumps (-m gpu): the application on the B200 (C++ host classes + CUDA
kernels through the session C ABI), free-running from the same deck, against tests/golden/*.npz -- binary dumps
written by the unmodified reference compiled over the host-only Kokkos stand-in (tests/golden/make_golden.py).
Step 0 must be exact in x and v (bit-identical lattice and velocities); later steps within 1e-10 of the global
RMS (BASELINE.json north_star), matched by atom id."""
import re
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle_py import REPO

GOLDEN = sorted((REPO / "tests" / "golden").glob("*.npz"))
sys.path.insert(0, str(REPO / "tests" / "golden"))
TOL = 1e-10


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_app_matches_reference_dumps(emd, tmp_path, path):
    import make_golden
    g = np.load(path)
    idial = int(g["idial"]) if "idial" in g.files else 0  # pair_style lj/cut/idial (ForceLJIDialNeigh) with this nrepeat
    deck = tmp_path / "in.deck"
    snap = "deck" in g.files  # SNAP fixtures derive from a shipped input/snap deck; its coefficient files sit beside the deck
    src = make_golden.SNAP_DIR / str(g["deck"]) if snap else make_golden.DECK
    make_golden.make_deck(deck, tuple(int(r) for r in g["region"]), int(g["nsteps"]), "on" if int(g["newton"]) else "off", src, idial)
    if snap:
        for f in make_golden.SNAP_DIR.glob("*.snap*"):
            (tmp_path / f.name).write_bytes(f.read_bytes())
    app = emd.App(["-il", str(deck), "--neigh-type", str(g["neigh"]), "--force-iteration", str(g["iteration"]), "--comm-type", "SERIAL"])
    steps = sorted(int(m.group(1)) for k in g.files if (m := re.match(r"s(\d+)_x", k)))
    L = np.array([float(r) for r in g["region"]]) * make_golden.lattice_constant(src)
    done = 0
    for s in steps:
        app.advance(s - done)
        done = s
        cur = app.download()
        o, r = np.argsort(cur["id"]), np.argsort(g[f"s{s}_id"])
        np.testing.assert_array_equal(cur["id"][o], g[f"s{s}_id"][r])
        if s == 0:
            np.testing.assert_array_equal(cur["x"][o], g["s0_x"][r])
            np.testing.assert_array_equal(cur["v"][o], g["s0_v"][r])
            np.testing.assert_array_equal(cur["id"], g["s0_id"])  # and the same cell-sorted atom order
        dx = cur["x"][o] - g[f"s{s}_x"][r]
        dx -= np.round(dx / L) * L
        assert np.abs(dx).max() / np.sqrt((g[f"s{s}_x"] ** 2).mean()) < TOL
        assert np.abs(cur["v"][o] - g[f"s{s}_v"][r]).max() / np.sqrt((g[f"s{s}_v"] ** 2).mean()) < TOL
        assert np.abs(cur["f"][o] - g[f"s{s}_f"][r]).max() / max(np.sqrt((g[f"s{s}_f"] ** 2).mean()), 1.0) < TOL
    # thermo table of the reference at print precision (the last printed row is at nsteps)
    T, PE, KE = app.thermo()
    row = g["thermo"][-1]
    if int(row[0]) != done:  # short SNAP fixtures: the reference printed only the step-0 row (thermo every 10)
        app.close()
        return
    assert int(row[0]) == done and abs(T - row[1]) < 2e-6 and abs(PE - row[2]) < 2e-6 and abs(PE + KE - row[3]) < 2e-6
    app.close()
