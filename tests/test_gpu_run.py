"""GPU whole-run parity (-m gpu): the application (C++ host classes + CUDA kernels, driven through
the session C ABI and through the ExaMiniMD binary with --dumpbinary) free-running against the
CPU oracle: per-atom x, v, f within 1e-10 of the global RMS after 100 steps, thermo to print
precision, and size-independent invariants at BASELINE.json's full 2 M-atom size."""
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle_py import OracleMD, REPO, lj_energy_envelope

DECK = REPO / "input" / "in.lj"
TOL = 1e-10


def by_id(d, n=None):
    o = np.argsort(d["id"][:n])
    return {k: np.asarray(v)[:n][o] for k, v in d.items()}


def compare(app_state, md, tol=TOL):
    n = md.geti("N_local")
    ref = by_id({k: md.arr(k) for k in ("id", "x", "v", "f")}, n)
    cur = by_id(app_state)
    np.testing.assert_array_equal(cur["id"], ref["id"])
    L = md.getd("domain_x")
    dx = cur["x"] - ref["x"]
    dx -= np.round(dx / L) * L  # an atom may be wrapped on one side and not yet on the other
    ex, ev = np.abs(dx).max() / np.sqrt((ref["x"] ** 2).mean()), np.abs(cur["v"] - ref["v"]).max() / np.sqrt((ref["v"] ** 2).mean())
    print(f"PARITY x {ex:.2e} v {ev:.2e} f {np.abs(cur['f'] - ref['f']).max() / max(np.sqrt((ref['f'] ** 2).mean()), 1.0):.2e}")
    assert ex < tol and ev < tol
    # on the perfect lattice (step 0) every force is zero to roundoff (~1e-13), so the scale is floored at 1
    # (LJ units; the liquid's RMS force is ~10) -- SURVEY.md section 7, hard part 6
    assert np.abs(cur["f"] - ref["f"]).max() / max(np.sqrt((ref["f"] ** 2).mean()), 1.0) < tol


@pytest.mark.parametrize("neigh,iteration", [("CSR", "NEIGH_HALF"), ("CSR", "NEIGH_FULL"), ("2D", "NEIGH_FULL"), ("2D", "NEIGH_HALF")])
def test_100_steps_vs_oracle(emd, neigh, iteration):
    region = (20, 20, 20)  # 32 000 atoms, rebuilds at steps 20..100
    md = OracleMD.from_deck(DECK, neigh, iteration, region=region)
    app = emd.App(["-il", str(DECK), "--neigh-type", neigh, "--force-iteration", iteration, "--comm-type", "SERIAL",
                   "--region", *map(str, region)])
    assert app.get("N") == md.geti("N") and app.get("N_ghost") == md.geti("N_ghost")
    compare(app.download(), md)  # step 0: bit-identical lattice, forces ~ 0
    T0, PE0, KE0 = app.thermo()
    n_at = md.geti("N")  # T0 = 1.4 exactly, PE0 = -6.332812, KE0 = 1.5*T0*(N-1)/N
    assert (f"{T0:.6f}", f"{PE0:.6f}") == ("1.400000", "-6.332812") and abs(KE0 - 2.1 * (n_at - 1) / n_at) < 1e-12
    for _ in range(5):
        app.advance(20)
        md.step(20)
        compare(app.download(), md)
        Ta, PEa, KEa = app.thermo()
        To, PEo, KEo = md.thermo()
        assert abs(Ta - To) < 1e-9 and abs(PEa - PEo) < 1e-9 and abs(KEa - KEo) < 1e-9
    assert app.get("N_ghost") == md.geti("N_ghost")
    if neigh == "CSR":
        assert app.get("total_neighs") == md.geti("total_neighs")
    # total-energy drift within the reference's own run-to-run envelope (north_star): the CPU path run with one thread and with
    # all cores (atomic force accumulation = another summation order) ends 100 steps at energies e1 and en; the GPU changes the
    # order of every row sum, contracts FMAs and replaces the IEEE divide, i.e. it is held to the FP64 parity bar of the
    # trajectory itself: 100 spreads or 1e-10 of |E| per atom, whichever is larger -- five orders of magnitude below the
    # drift (~1e-5 per atom over these 100 steps).  The measured numbers are printed (and kept in profiles/).
    e1, en = lj_energy_envelope(region, 100, iteration)
    envelope = max(100.0 * abs(e1 - en), 1e-10 * abs(e1))
    drift_ref, drift_gpu = e1 - (PE0 + KE0), (PEa + KEa) - (PE0 + KE0)
    print(f"DRIFT {neigh} {iteration}: reference {drift_ref:.6e} gpu {drift_gpu:.6e} |diff| {abs(drift_gpu - drift_ref):.2e} "
          f"thread-to-thread spread {abs(e1 - en):.2e} envelope {envelope:.2e}")
    assert abs(e1 - (PEo + KEo)) <= envelope  # the serial library oracle is one of the reference's own runs
    assert abs(drift_gpu - drift_ref) <= envelope
    app.close(); md.close()


def test_fused_integrator_is_bit_identical(emd):
    """advance(n) folds final_integrate of a step and initial_integrate of the next into the force launch
    (Force::compute_with_nve) or into one kernel (Integrator::final_initial_integrate); advance(1) n times never does: x, v, f must agree bit for bit, across a
    re-neighboring, and the launch count must show the fusion"""
    argv = ["-il", str(DECK), "--neigh-type", "CSR", "--force-iteration", "NEIGH_HALF", "--comm-type", "SERIAL", "--region", "12", "12", "12"]
    a, b = emd.App(argv), emd.App(argv)
    la, lb = a.launches(), b.launches()
    a.advance(45)
    for _ in range(45):
        b.advance(1)
    sa, sb = a.download(), b.download()
    for k in ("id", "x", "v", "f"):
        np.testing.assert_array_equal(sa[k], sb[k])
    # per fused step boundary: the two integrator kernels disappear into the force launch (88), or -- when the force module
    # cannot take the integrator along -- become one kernel (44)
    assert (b.launches() - lb) - (a.launches() - la) in (44, 88)
    a.close(); b.close()


def test_thermo_step_keeps_the_fused_integrator(emd, oracle_lib):
    """run(n) (the reference's loop: thermo every 10 steps) serves a thermo step from the force launch itself: potential energy
    at the step's positions and sum m v^2 of the velocities between the two kicks (what the reference's Temperature / PotE /
    KinE read after final_integrate), with the integrator kicks still fused.  T, PE, KE of steps 10..40 against the oracle,
    the trajectory bit for bit against the unfused sequence, and one launch per thermo step instead of force + energy +
    integrator + reduction kernels."""
    argv = ["-il", str(DECK), "--neigh-type", "CSR", "--force-iteration", "NEIGH_HALF", "--comm-type", "SERIAL", "--region", "12", "12", "12"]
    a, b = emd.App(argv), emd.App(argv)
    md = OracleMD.from_deck(DECK, "CSR", "NEIGH_HALF", region=(12, 12, 12))
    # a thermo step takes the fused path unless it is the last step of the call: run(15) returns the thermo of step 10, the next
    # run(10) (steps 16..25) that of step 20 (a re-neighboring step as well), ...
    done = 0
    for n_run, at in ((15, 10), (10, 20), (10, 30)):
        T, PE, KE = a.run(n_run)
        md.step(at - done)
        To, PEo, KEo = md.thermo()
        assert abs(T - To) < 1e-10 * abs(To) and abs(PE - PEo) < 1e-10 * abs(PEo) and abs(KE - KEo) < 1e-10 * abs(KEo), (at, T, To, PE, PEo, KE, KEo)
        done += n_run
        md.step(done - at)
    a.run(60 - done)
    for _ in range(60):
        b.advance(1)
    sa, sb = a.download(), b.download()
    for k in ("id", "x", "v", "f"):
        np.testing.assert_array_equal(sa[k], sb[k])
    a.close(); b.close(); md.close()


@pytest.mark.parametrize("no_tiles", ["0", "1"])
def test_binary_dump_and_correctness_flags(emd, tmp_path, no_tiles):
    """the ExaMiniMD executable with the reference's own record/replay flags (README.md:99-107); once on the
    tile fast path and once (EMD_NO_TILES=1) on the generic list kernels."""
    import os
    region = ["--region", "10", "10", "10", "--nsteps", "40"]
    refdir = tmp_path / "ref"; refdir.mkdir()
    md = OracleMD.from_deck(DECK, "CSR", "NEIGH_HALF", region=(10, 10, 10))
    md.L.orcf_dump(md.h, str(refdir).encode(), 0)
    for s in range(1, 41):
        md.step(1)
        if s % 20 == 0:
            md.L.orcf_dump(md.h, str(refdir).encode(), s)
    out = tmp_path / "correctness.dat"
    mine = tmp_path / "mine"; mine.mkdir()
    r = subprocess.run([str(emd.EXE_PATH), "-il", str(DECK), "--neigh-type", "CSR", "--force-iteration", "NEIGH_HALF",
                        "--comm-type", "SERIAL", *region, "--dumpbinary", "20", str(mine), "--correctness", "20", str(refdir), str(out)],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, EMD_NO_TILES=no_tiles))
    assert r.returncode == 0, r.stderr
    assert "Using: ForceLJNeighHalf NeighborCSR CommSerial BinningKKSort" in r.stdout
    l0 = [l.split() for l in r.stdout.splitlines() if l.startswith("0 1.400000 -6.332812 ")]
    assert l0 and abs(float(l0[0][3]) - (-6.332811993 + 2.1 * 3999 / 4000)) < 2e-6  # KE0 = 1.5*T0*(N-1)/N, N = 4000
    assert "PERFORMANCE" in r.stdout
    rows = [l.split() for l in out.read_text().splitlines() if not l.startswith("#")]
    assert [int(r_[0]) for r_ in rows] == [0, 20, 40]
    for r_ in rows:  # absolute norms / max deltas of r, v, f against the oracle's dumps
        assert max(float(t) for t in r_[1:]) < 1e-8
    # dump format: int n; id; type; q; x; v; f (examinimd.cpp:337-343)
    raw = (mine / "output.0000000040.000").read_bytes()
    n = struct.unpack("i", raw[:4])[0]
    assert n == 4000 and len(raw) == 4 + n * 88
    md.close()


def test_full_size_invariants(emd):
    """BASELINE configs[1]: LJ fcc 2 048 000 atoms, half CSR list.  Too big for the oracle in seconds,
    so check size-independent properties: known step-0 answers, 39+ neighbors per atom, momentum and
    energy conservation across two rebuilds, sorted/consistent CSR."""
    import torch
    app = emd.App(["-il", str(DECK), "--neigh-type", "CSR", "--force-iteration", "NEIGH_HALF", "--comm-type", "SERIAL",
                   "--region", "80", "80", "80"])
    n = app.get("N")
    assert n == 2048000 and (app.get("nbinx"), app.get("nbiny"), app.get("nbinz")) == (49, 49, 49)
    T0, PE0, KE0 = app.thermo()
    assert (f"{T0:.6f}", f"{PE0:.6f}") == ("1.400000", "-6.332812")
    st = app.download()
    assert np.abs(st["f"]).max() < 1e-10
    app.advance(45)
    T, PE, KE = app.thermo()
    assert abs((PE + KE) - (PE0 + KE0)) < 5e-4  # sanity only: the drift is compared with the reference's in tests/test_gpu_fullsize.py
    st = app.download()
    assert np.array_equal(np.sort(st["id"]), np.arange(1, n + 1, dtype=np.int32))  # a permutation of all atoms
    p = st["v"].sum(0) * 2.0
    assert np.abs(p).max() / n < 1e-12  # momentum stays zero (Newton's third law in the half list)
    total = app.get("total_neighs")
    assert total >= 39 * n * 0.9
    app.close()


@pytest.mark.parametrize("deck,region", [(DECK, (12, 10, 14)), (REPO / "input" / "snap" / "in.snap.W", (5, 6, 7))])
def test_device_lattice_is_bit_identical_to_the_host_loops(emd, deck, region, monkeypatch):
    """Input::create_lattice / create_velocities on the device (kernels/lattice.cu) against the host loops they replace
    (src/input.cpp:460-792): atom order, positions, ids, types and velocities bit for bit"""
    import os
    cwd = os.getcwd()
    os.chdir(deck.parent)  # the SNAP deck names its coefficient files relative to the working directory
    try:
        states = []
        for host in ("1", "0"):
            monkeypatch.setenv("EMD_HOST_LATTICE", host)
            app = emd.App(["-il", str(deck), "--comm-type", "SERIAL", "--region", *map(str, region)])
            states.append(app.download())
            app.close()
    finally:
        os.chdir(cwd)
    a, b = states
    assert a["x"].shape[0] > 0
    for k in ("id", "type", "x", "v"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
