"""GPU parity AT THE BASELINE SIZES (-m gpu): the two single-GPU configurations BASELINE.json quotes the metric on, compared
atom by atom (by id) with a whole run of the reference's own CPU path on the box's host cores -- oracle/_ref/ExaMiniMD_ref_omp
(the unmodified reference sources over the OpenMP Kokkos stand-in) with --dumpbinary, or the OpenMP restatement when that
binary is missing.  Tolerance 1e-10 of the global RMS (north_star); tile-shape selection, list capacities and 32-bit offsets are
exercised at full size here, not only through invariants."""
import re
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle_py import REPO, cpu_run_dump

TOL = 1e-10


def deck_text(src, region, nsteps):
    txt = src.read_text()
    txt = re.sub(r"region\s+box block.*", "region\t\tbox block 0 %d 0 %d 0 %d" % tuple(region), txt)
    return re.sub(r"run\s+\d+", "run\t\t%d" % nsteps, txt)


def compare_by_id(cur, ref, box):
    o, r = np.argsort(cur["id"]), np.argsort(ref["id"])
    np.testing.assert_array_equal(cur["id"][o], ref["id"][r])
    dx = cur["x"][o] - ref["x"][r]
    dx -= np.round(dx / box) * box
    ex = np.abs(dx).max() / np.sqrt((ref["x"] ** 2).mean())
    ev = np.abs(cur["v"][o] - ref["v"][r]).max() / np.sqrt((ref["v"] ** 2).mean())
    ef = np.abs(cur["f"][o] - ref["f"][r]).max() / max(np.sqrt((ref["f"] ** 2).mean()), 1e-3)
    return ex, ev, ef


def test_lj_2M_half_csr_vs_reference(emd, tmp_path):
    """configs[1]: LJ fcc 2 048 000 atoms, half CSR list; 25 steps = 20 steps on the step-0 lists + one re-neighboring (sort,
    ghosts, bins, lists rebuilt off-lattice) + 5 more"""
    region, nsteps = (80, 80, 80), 25
    src = REPO / "input" / "in.lj"
    t0 = time.time()
    ref, out = cpu_run_dump(deck_text(src, region, nsteps), [nsteps], neigh="CSR", iteration="NEIGH_HALF")
    t_cpu = time.time() - t0
    deck = tmp_path / "in.deck"
    deck.write_text(deck_text(src, region, nsteps))
    app = emd.App(["-il", str(deck), "--neigh-type", "CSR", "--force-iteration", "NEIGH_HALF", "--comm-type", "SERIAL"])
    assert app.get("N") == 2048000
    assert app.device_ptr("tiles"), "the 2 M-atom configuration must run on the tile path"
    app.advance(nsteps)
    cur = app.download()
    a = (4.0 / 0.8442) ** (1.0 / 3.0)
    ex, ev, ef = compare_by_id(cur, ref[nsteps], np.array(region) * a)
    print(f"FULLSIZE lj 2048000 atoms x {nsteps} steps vs CPU reference ({t_cpu:.1f} s): err x={ex:.2e} v={ev:.2e} f={ef:.2e}")
    assert ex < TOL and ev < TOL and ef < TOL, (ex, ev, ef)
    # thermo rows the reference printed (6 decimals) at steps 10 and 20 are reproduced by a second, observed run below the
    # print precision: compare the last printed row with our thermo at the same step through a fresh application
    rows = [[float(t) for t in l.split()[:4]] for l in out.splitlines() if re.match(r"^\d+ -?\d+\.\d+ ", l)]
    app2 = emd.App(["-il", str(deck), "--neigh-type", "CSR", "--force-iteration", "NEIGH_HALF", "--comm-type", "SERIAL"])
    last = rows[-1]
    th = app2.run(int(last[0]))
    assert th is not None and abs(th[0] - last[1]) < 2e-6 and abs(th[1] - last[2]) < 2e-6 and abs(th[1] + th[2] - last[3]) < 2e-6, (th, last)
    app.close(); app2.close()


def test_snap_250k_vs_reference(emd, tmp_path):
    """configs[2]: SNAP tungsten (in.snap.W, 2J=8), full list, 250 000 atoms; one step (two force evaluations and one
    re-neighboring: the deck rebuilds every step)"""
    region, nsteps = (50, 50, 100), 1
    sdir = REPO / "input" / "snap"
    files = sorted(sdir.glob("*.snap*"))
    txt = deck_text(sdir / "in.snap.W", region, nsteps)
    t0 = time.time()
    ref, _ = cpu_run_dump(txt, [nsteps], neigh="CSR", iteration="NEIGH_FULL", files=files)
    t_cpu = time.time() - t0
    deck = tmp_path / "in.deck"
    deck.write_text(txt)
    for f in files:
        (tmp_path / f.name).write_bytes(f.read_bytes())
    app = emd.App(["-il", str(deck), "--neigh-type", "CSR", "--comm-type", "SERIAL"])
    assert app.get("N") == 250000
    app.advance(nsteps)
    cur = app.download()
    ex, ev, ef = compare_by_id(cur, ref[nsteps], np.array(region) * 3.1803)
    print(f"FULLSIZE snap 250000 atoms x {nsteps} step vs CPU reference ({t_cpu:.1f} s): err x={ex:.2e} v={ev:.2e} f={ef:.2e}")
    assert ex < TOL and ev < TOL and ef < TOL, (ex, ev, ef)
    app.close()
