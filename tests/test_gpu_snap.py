"""GPU parity of the SNAP force (-m gpu): kernels/snap.cu through the C ABI against the CPU oracle
(oracle/oracle_snap.c, itself bit-identical to the reference).  FP64 throughout; the GPU evaluates the same
force in the adjoint order (Y = sum beta Z, then conj(dU).Y), so sums are re-associated: tolerance 1e-10
of the global force scale (BASELINE.json north_star), U_tot to 1e-12 relative."""
import ctypes as C
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle_py import OracleMD, REPO

SNAP_DIR = REPO / "input" / "snap"
TOL = 1e-10


def snap_deck(tmp_path, name, region, nsteps):
    txt = (SNAP_DIR / name).read_text()
    txt = re.sub(r"region\s+box block.*", "region\t\tbox block 0 %d 0 %d 0 %d" % tuple(region), txt)
    txt = re.sub(r"run\s+\d+", "run\t\t%d" % nsteps, txt)
    p = tmp_path / "in.deck"
    p.write_text(txt)
    for f in SNAP_DIR.glob("*.snap*"):
        (tmp_path / f.name).write_bytes(f.read_bytes())
    return p


def half_index(j, mb, ma):
    return sum((jj // 2 + 1) * (jj + 1) for jj in range(j)) + mb * (j + 1) + ma


@pytest.mark.parametrize("deck,region,neigh", [("in.snap.W", (4, 4, 5), "CSR"), ("in.snap.Ta06A", (4, 4, 4), "2D"), ("in.snap.W", (5, 6, 4), "2D")])
def test_snap_run_matches_oracle(emd, oracle_lib, tmp_path, deck, region, neigh):
    """free-running trajectories, thermalised lattice: x, v, f by atom id at steps 0, 1, 5, 12"""
    d = snap_deck(tmp_path, deck, region, 12)
    app = emd.App(["-il", str(d), "--neigh-type", neigh, "--comm-type", "SERIAL"])
    md = OracleMD.from_deck(d, neigh, "NEIGH_FULL", coeff_dir=tmp_path)
    n = md.geti("N_local")
    assert app.get("N_local") == n
    done = 0
    for s in (0, 1, 5, 12):
        app.advance(s - done)
        md.step(s - done)
        done = s
        cur = app.download()
        o, r = np.argsort(cur["id"]), np.argsort(md.arr("id")[:n])
        fo = md.arr("f")[:n][r]
        fscale = max(np.sqrt((fo ** 2).mean()), 1e-3)
        assert np.abs(cur["f"][o] - fo).max() / fscale < TOL, f"step {s}: forces"
        assert np.abs(cur["v"][o] - md.arr("v")[:n][r]).max() / np.sqrt((md.arr("v")[:n] ** 2).mean()) < TOL, f"step {s}: v"
        assert np.abs(cur["x"][o] - md.arr("x")[:n][r]).max() < 1e-9, f"step {s}: x"
    T, PE, KE = app.thermo()
    To, PEo, KEo = md.thermo()
    assert PE == 0.0 and PEo == 0.0  # Force::compute_energy is not overridden by ForceSNAP (src/force.h:54)
    assert abs(T - To) < 1e-7 * To
    app.close()
    md.close()


def test_snap_utot_and_pair_list_match_oracle(emd, oracle_lib, tmp_path):
    """function level, after 6 steps of motion: U_tot of every atom (half range) and the in-cutoff pair list"""
    import torch
    d = snap_deck(tmp_path, "in.snap.W", (4, 5, 4), 6)
    app = emd.App(["-il", str(d), "--neigh-type", "CSR", "--comm-type", "SERIAL"])
    md = OracleMD.from_deck(d, "CSR", "NEIGH_FULL", coeff_dir=tmp_path)
    app.advance(6)
    md.step(6)
    cur = app.download()
    n = md.geti("N_local")
    # same atom order on both sides (cell sort is bit-exact; positions agree to 1e-12) -> compare index by index
    np.testing.assert_array_equal(cur["id"], md.arr("id")[:n])
    L = emd.lib()
    snap = C.c_void_p(L.emd_app_device_ptr(app.handle, b"snap"))
    assert snap.value, "no ForceSNAP in the app"
    ncoeff, nuh, ntri, npairs, ustride = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rcutmax = C.c_double()
    emd.check(L.emd_snap_info(snap, C.byref(ncoeff), C.byref(nuh), C.byref(ntri), C.byref(rcutmax), C.byref(npairs), C.byref(ustride)))
    assert (ncoeff.value, nuh.value, ntri.value) == (55, 155, 125)
    assert abs(rcutmax.value - 4.73442) < 1e-12

    def dl(name, count, dtype):
        out = np.empty(count, dtype)
        emd.check(L.emd_memcpy_d2h(app.ctx, out.ctypes.data_as(C.c_void_p), C.c_void_p(L.emd_snap_device_ptr(snap, name.encode())),
                                   out.nbytes))
        app.sync()
        return out

    ulist = dl("ulist", nuh.value * ustride.value * 2, np.float64).reshape(nuh.value, ustride.value, 2)
    poff = dl("pair_offsets", n + 1, np.int32)
    pj = dl("pair_j", npairs.value, np.int32)
    assert poff[-1] == npairs.value
    worst = 0.0
    for i in range(0, n, 3):
        U, inside, _ = md.snap_probe(i, with_forces=False)
        np.testing.assert_array_equal(pj[poff[i]:poff[i + 1]], inside)  # same neighbors, same order
        for j in range(9):
            for mb in range(j // 2 + 1):
                for ma in range(j + 1):
                    g = ulist[half_index(j, mb, ma), i]
                    worst = max(worst, abs(complex(g[0], g[1]) - U[j, ma, mb]))
    assert worst < 1e-12, worst
    app.close()
    md.close()


@pytest.mark.parametrize("direct_only", [False, True])
def test_snap_deidrj_direct_recursion(emd, oracle_lib, tmp_path, monkeypatch, direct_only):
    """snap_deidrj evaluates dU through unit tangents and Euler's relation, dividing by a_r; pairs with |a_r| < 0.25 (theta0
    near pi/2, r near half the cutoff) go to the second launch with the direct recursion.  A lattice compressed to a
    nearest-neighbor distance of 2.39 (a_r ~ 0) puts the 6 nearest neighbors of every atom on that path and the other 20
    on the first one; EMD_SNAP_DEIDRJ_DIRECT=1 sends every pair there."""
    if direct_only:
        monkeypatch.setenv("EMD_SNAP_DEIDRJ_DIRECT", "1")
    d = snap_deck(tmp_path, "in.snap.W", (4, 5, 6), 3)
    if not direct_only:
        d.write_text(re.sub(r"lattice\s+sc\s+\S+", "lattice         sc 2.39", d.read_text()))
    app = emd.App(["-il", str(d), "--neigh-type", "CSR", "--comm-type", "SERIAL"])
    md = OracleMD.from_deck(d, "CSR", "NEIGH_FULL", coeff_dir=tmp_path)
    n = md.geti("N_local")
    done = 0
    for s_ in (0, 1, 3):
        app.advance(s_ - done)
        md.step(s_ - done)
        done = s_
        cur = app.download()
        o, r = np.argsort(cur["id"]), np.argsort(md.arr("id")[:n])
        fo = md.arr("f")[:n][r]
        fscale = max(np.sqrt((fo ** 2).mean()), 1e-3)
        assert np.abs(cur["f"][o] - fo).max() / fscale < TOL, f"step {s_}: forces"
        assert np.abs(cur["x"][o] - md.arr("x")[:n][r]).max() < 1e-9, f"step {s_}: x"
    app.close()
    md.close()
