"""CPU tests: the oracle restatement against the derived known answers of SURVEY.md App. C
(the reference ships no golden vectors of its own) and against its own cross-variant invariants
(full vs half lists, CSR vs 2D lists must give the same physics -- src/force_types/force_lj_neigh_impl.h)."""
import math
import struct

import numpy as np
import pytest

from oracle_py import OracleMD, REPO

DECK = REPO / "input" / "in.lj"
A_FCC = 1.6795961913825073  # pow(4/0.8442, 1/3), src/input.cpp:319


@pytest.fixture(scope="module")
def small_full():
    md = OracleMD.from_deck(DECK, "CSR", "NEIGH_FULL", region=(12, 12, 12))
    yield md
    md.close()


def test_lattice_known_answers(oracle_lib):
    md = OracleMD.from_deck(DECK, "CSR", "NEIGH_FULL", region=(6, 6, 6), setup=False)
    assert md.geti("N") == 4 * 6 ** 3
    assert md.getd("domain_x") == A_FCC * 6
    x = md.arr("x")
    # first fcc cell: basis (0,0,0),(.5,.5,0),(.5,0,.5),(0,.5,.5) times a (src/input.cpp:600-603)
    np.testing.assert_array_equal(x[:4], A_FCC * np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]]))
    np.testing.assert_array_equal(md.arr("id"), np.arange(1, 865, dtype=np.int32))
    v = md.arr("v")
    # zero total momentum and T rescaled exactly to the target (src/input.cpp:760-785)
    assert np.abs(v.sum(0)).max() < 1e-10
    m = md.getd("mass0")
    Tm = (v * v).sum() * m / (3 * md.geti("N") - 3)
    assert abs(Tm - 1.4) < 1e-12
    md.close()


def test_in_lj_geometry_known_answers(oracle_lib):
    md = OracleMD.from_deck(DECK, "CSR", "NEIGH_FULL", setup=False)
    assert md.geti("N") == 256000
    assert abs(md.getd("domain_x") - 67.1838476553003) < 1e-12
    md.stage("exchange", "bin_sort")
    assert (md.geti("nbinx"), md.geti("nbiny"), md.geti("nbinz")) == (25, 25, 25)  # 23 interior + 2 halo
    assert abs((md.getd("maxx") - md.getd("minx")) / 25 - 2.921036854578274) < 1e-2
    bc = md.arr("bincount")
    assert bc.sum() == 256000
    bo = md.arr("binoffsets")
    np.testing.assert_array_equal(bo, np.concatenate([[0], np.cumsum(bc)[:-1]]))
    md.close()


def test_step0_thermo_in_lj(oracle_lib):
    md = OracleMD.from_deck(DECK, "CSR", "NEIGH_HALF")
    T, PE, KE = md.thermo()
    # `0 1.400000 -6.332812 -4.232820`
    assert f"{T:.6f}" == "1.400000"
    assert f"{PE:.6f}" == "-6.332812"
    assert f"{PE + KE:.6f}" == "-4.232820"
    rm = md.arr("row_map")
    counts = np.diff(rm)
    # half list, newton off: 39 owned-pair partners + every ghost partner
    assert counts.min() >= 39
    md.close()


def test_neighbor_counts_full(small_full):
    rows = small_full.rows()
    assert all(len(r) == 78 for r in rows)  # fcc shells 12+6+24+12+24 inside 2.8
    # no self, no duplicates
    for i in (0, 17, 5000):
        assert i not in rows[i] and len(set(rows[i].tolist())) == 78


def test_step0_forces_vanish(small_full):
    f = small_full.arr("f")[: small_full.geti("N_local")]
    assert np.abs(f).max() < 1e-11  # lattice inversion symmetry


def test_binning_is_stable_counting_sort(small_full):
    md = small_full
    perm = md.arr("permute")
    bc, bo = md.arr("bincount"), md.arr("binoffsets")
    n = md.geti("bin_range")
    assert sorted(perm.tolist()) == list(range(n))
    for c in np.nonzero(bc)[0][:200]:
        seg = perm[bo[c]: bo[c] + bc[c]]
        assert np.all(np.diff(seg) > 0)  # ascending index inside a bin = 1-thread arrival order


def _run(neigh, iteration, nsteps, region=(8, 8, 8)):
    md = OracleMD.from_deck(DECK, neigh, iteration, region=region)
    md.step(nsteps)
    n = md.geti("N_local")
    order = np.argsort(md.arr("id")[:n])
    out = (md.arr("x")[:n][order], md.arr("v")[:n][order], md.arr("f")[:n][order], md.thermo())
    md.close()
    return out


def test_variants_agree_and_energy_conserved(oracle_lib):
    ref = _run("CSR", "NEIGH_FULL", 45)
    for neigh, it in (("CSR", "NEIGH_HALF"), ("2D", "NEIGH_FULL"), ("2D", "NEIGH_HALF")):
        x, v, f, th = _run(neigh, it, 45)
        scale = np.sqrt((ref[2] ** 2).mean())
        assert np.abs(x - ref[0]).max() < 1e-10
        assert np.abs(f - ref[2]).max() / scale < 1e-10
    E0 = sum(_run("CSR", "NEIGH_FULL", 0)[3][1:])
    E = ref[3][1] + ref[3][2]
    assert abs(E - E0) < 2e-3  # NVE drift over 45 steps at dt=0.005


def test_2d_list_resize_rule(oracle_lib):
    md = OracleMD.from_deck(DECK, "2D", "NEIGH_FULL", region=(8, 8, 8))
    assert md.geti("maxneighs") == int(78 * 1.2)  # 16 -> overflow -> 78*1.2 = 93 (neighbor_2d.h:322-326)
    assert md.geti("fill_passes") == 2
    md.close()


def test_dump_binary_format(oracle_lib, tmp_path):
    md = OracleMD.from_deck(DECK, "CSR", "NEIGH_FULL", region=(5, 5, 5))
    md.L.orcf_dump(md.h, str(tmp_path).encode(), 7)
    raw = (tmp_path / "output.0000000007.000").read_bytes()
    n = struct.unpack("i", raw[:4])[0]
    assert n == 500 and len(raw) == 4 + n * (4 + 4 + 8 + 72)
    ids = np.frombuffer(raw, np.int32, n, 4)
    x = np.frombuffer(raw, np.float64, 3 * n, 4 + 16 * n).reshape(n, 3)
    np.testing.assert_array_equal(ids, md.arr("id")[:n])
    np.testing.assert_array_equal(x, md.arr("x")[:n])
    md.close()


def test_velocity_rng_is_position_hashed(oracle_lib):
    """loop-geom velocities depend only on (seed, position): a sub-lattice reproduces them (input.h:100-132)."""
    a = OracleMD.from_deck(DECK, "CSR", "NEIGH_FULL", region=(4, 4, 4), setup=False)
    L = a.L
    import ctypes as C
    seed = C.c_int(0)
    L.orc_random_reset.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_void_p]
    L.orc_random_uniform_state.argtypes = [C.POINTER(C.c_int)]
    L.orc_random_uniform_state.restype = C.c_double
    pos = np.array([0.0, 0.0, 0.0])
    L.orc_random_reset(C.byref(seed), 87287, pos.ctypes.data)
    u = [L.orc_random_uniform_state(C.byref(seed)) for _ in range(3)]
    assert all(0.0 < t < 1.0 for t in u)
    # Park-Miller recurrence: seed_{k+1} = 16807*seed_k mod (2^31-1)
    s0 = seed.value
    L.orc_random_uniform_state(C.byref(seed))
    assert seed.value == (16807 * s0) % 2147483647
    a.close()
