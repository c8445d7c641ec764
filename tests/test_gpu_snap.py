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


# ----------------------------------------------------------------------------- SNAP energy
def _cg(j1, j2, j, m1, m2):
    """Clebsch-Gordan coefficient in the doubled-index convention of SNA::init_clebsch_gordan (sna_impl.hpp:991-1046)"""
    from math import factorial as f, sqrt
    aa2, bb2 = 2 * m1 - j1, 2 * m2 - j2
    m = (aa2 + bb2 + j) // 2
    if m < 0 or m > j:
        return 0.0
    s = 0.0
    for z in range(max(0, max(-(j - j2 + aa2) // 2, -(j - j1 - bb2) // 2)), min((j1 + j2 - j) // 2, min((j1 - aa2) // 2, (j2 + bb2) // 2)) + 1):
        s += (-1 if z % 2 else 1) / (f(z) * f((j1 + j2 - j) // 2 - z) * f((j1 - aa2) // 2 - z) * f((j2 + bb2) // 2 - z) *
                                     f((j - j2 + aa2) // 2 + z) * f((j - j1 - bb2) // 2 + z))
    cc2 = 2 * m - j
    dcg = sqrt(f((j1 + j2 - j) // 2) * f((j1 - j2 + j) // 2) * f((-j1 + j2 + j) // 2) / f((j1 + j2 + j) // 2 + 1))
    sf = sqrt(f((j1 + aa2) // 2) * f((j1 - aa2) // 2) * f((j2 + bb2) // 2) * f((j2 - bb2) // 2) * f((j + cc2) // 2) * f((j - cc2) // 2) * (j + 1))
    return s * dcg * sf


def _snap_energy_numpy(U, beta, twojmax=8, bzero=False, wself=1.0):
    """sum_i [beta_0 + sum_k beta_k (B_k(i) - bzero_k)] from U_tot[atom, j, ma, mb]: a plain restatement of the published
    SNA::compute_zi / compute_bi (LAMMPS src/SNAP/sna.cpp; compute_zi = sna_impl.hpp:196-283), vectorised over atoms only"""
    n = U.shape[0]
    e = np.full(n, beta[0])
    k = 0
    for j1 in range(twojmax + 1):
        for j2 in range(j1 + 1):
            for j in range(j1 - j2, min(twojmax, j1 + j2) + 1, 2):
                if j < j1:
                    continue
                k += 1
                b = np.zeros(n)
                for mb in range(j // 2 + 1):
                    for ma in range(j + 1):
                        w = 1.0
                        if 2 * mb == j:
                            w = 1.0 if ma < mb else (0.5 if ma == mb else 0.0)
                        if w == 0.0:
                            continue
                        z = np.zeros(n, complex)
                        for ma1 in range(max(0, (2 * ma - j - j2 + j1) // 2), min(j1, (2 * ma - j + j2 + j1) // 2) + 1):
                            ma2 = (2 * ma - j - (2 * ma1 - j1) + j2) // 2
                            ca = _cg(j1, j2, j, ma1, ma2)
                            for mb1 in range(max(0, (2 * mb - j - j2 + j1) // 2), min(j1, (2 * mb - j + j2 + j1) // 2) + 1):
                                mb2 = (2 * mb - j - (2 * mb1 - j1) + j2) // 2
                                z += ca * _cg(j1, j2, j, mb1, mb2) * U[:, j1, ma1, mb1] * U[:, j2, ma2, mb2]
                        b += w * (U[:, j, ma, mb].conjugate() * z).real
                b *= 2.0
                if bzero:
                    b -= wself ** 3 * (j + 1)
                e += beta[k] * b
    assert k == len(beta) - 1
    return e.sum()


def test_snap_energy(emd, oracle_lib, tmp_path, monkeypatch):
    """EMD_SNAP_ENERGY=1 (an extension: the reference's ForceSNAP has no energy).  (i) the value against a plain numpy
    restatement of the published compute_zi / compute_bi on the oracle's U_tot plus the pair term of the force kernel's
    1.5e6 / r^14; (ii) it is the potential of the forces: d(PE)/dt = -sum F.v along the trajectory (central difference over
    two steps: 2e-5 of the scale measured); (iii) total energy is conserved to 2e-5 of the kinetic energy (3e-6 measured)
    over the 40 steps before the first pair crosses the cutoff, where the reference's unshifted 1/r^12 term jumps by
    2 x 1.25e5 / rcut^12 = 1.97e-3 eV (tools/dbg_snap_energy.py shows the steps)."""
    monkeypatch.setenv("EMD_SNAP_ENERGY", "1")
    d = snap_deck(tmp_path, "in.snap.W", (4, 4, 4), 80)
    app = emd.App(["-il", str(d), "--neigh-type", "CSR", "--comm-type", "SERIAL"])
    md = OracleMD.from_deck(d, "CSR", "NEIGH_FULL", coeff_dir=tmp_path)
    n = md.geti("N_local")
    app.advance(6)
    md.step(6)
    # (i)
    coeff = [float(t) for t in " ".join(l for l in (SNAP_DIR / "W.snapcoeff").read_text().splitlines() if l.strip() and not l.startswith("#")).split()[5:]]
    assert len(coeff) == 56
    U = np.stack([md.snap_probe(i, with_forces=False)[0] for i in range(n)])
    x = md.arr("x")
    rep = 0.0
    for i in range(n):
        inside = md.snap_probe(i, with_forces=False)[1]
        r2 = ((x[inside] - x[i]) ** 2).sum(1)
        rep += (1.25e5 / r2 ** 6).sum()
    e_np = _snap_energy_numpy(U, coeff) + rep
    T, pe, ke = app.thermo()
    assert abs(pe * n - e_np) < 1e-10 * abs(e_np), (pe * n, e_np)
    # (ii), (iii)
    dt = 0.001  # units metal (input.cpp)
    pes, kes, fv = [], [], []
    for s_ in range(32):
        T, pe, ke = app.thermo()
        cur = app.download()
        pes.append(pe * n); kes.append(ke * n); fv.append((cur["f"] * cur["v"]).sum())
        app.advance(1)
    pes, kes, fv = np.array(pes), np.array(kes), np.array(fv)
    dpe = (pes[2:] - pes[:-2]) / (2 * dt)
    scale = np.abs(fv).max()
    assert np.abs(dpe + fv[1:-1]).max() < 1e-4 * scale, (np.abs(dpe + fv[1:-1]).max(), scale)
    etot = pes + kes
    assert np.abs(etot - etot[0]).max() < 2e-5 * kes.mean(), (np.abs(etot - etot[0]).max(), kes.mean())
    app.close()
    md.close()
