"""CPU tests that PIN the oracle: the C restatement (oracle/*.c) against the reference's own code.

The golden fixtures in tests/golden/*.npz are dumps written by the UNMODIFIED reference sources compiled over
the host-only Kokkos stand-in (oracle/Makefile.ref, generator: tests/golden/make_golden.py).  The oracle must
reproduce them BIT FOR BIT -- same atom order, same x, v, f -- at every dumped step, for CSR/2D x half/full
lists and for newton on; where oracle/_ref/ExaMiniMD_ref exists (it is built in the container that has
/root/reference and travels to the GPU box) a live run at another size is compared as well."""
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle_py import OracleMD, REPO

GOLDEN = sorted((REPO / "tests" / "golden").glob("*.npz"))
REF_EXE = REPO / "oracle" / "_ref" / "ExaMiniMD_ref"
sys.path.insert(0, str(REPO / "tests" / "golden"))


SNAP_DIR = REPO / "input" / "snap"


def deck_for(tmp_path, region, nsteps, newton, deck=None, idial=0):
    import make_golden
    p = tmp_path / "in.deck"
    make_golden.make_deck(p, tuple(int(r) for r in region), int(nsteps), "on" if newton else "off", deck or make_golden.DECK, idial)
    return p


def check_against(md, g, steps):
    n = md.geti("N_local")
    done = 0
    for s in steps:
        md.step(s - done)
        done = s
        np.testing.assert_array_equal(md.arr("id")[:n], g[f"s{s}_id"], err_msg=f"step {s}: atom order")
        for k in ("x", "v", "f"):
            np.testing.assert_array_equal(md.arr(k)[:n], g[f"s{s}_{k}"], err_msg=f"step {s}: {k} not bit-identical to the reference")


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_oracle_reproduces_reference_dumps_bit_for_bit(oracle_lib, tmp_path, path):
    g = np.load(path)
    snap = "deck" in g.files  # SNAP fixtures name the shipped deck they derive from (input/snap/)
    idial = int(g["idial"]) if "idial" in g.files else 0  # pair_style lj/cut/idial fixtures (ForceLJIDialNeigh)
    deck = deck_for(tmp_path, g["region"], g["nsteps"], int(g["newton"]), SNAP_DIR / str(g["deck"]) if snap else None, idial)
    md = OracleMD.from_deck(deck, str(g["neigh"]), str(g["iteration"]), coeff_dir=SNAP_DIR if snap else None)
    steps = sorted(int(m.group(1)) for k in g.files if (m := re.match(r"s(\d+)_x", k)))
    assert steps[0] == 0
    np.testing.assert_array_equal(md.arr("type")[: md.geti("N_local")], g["s0_type"])
    check_against(md, g, steps)
    md.close()


def test_oracle_thermo_matches_reference_table(oracle_lib, tmp_path):
    g = np.load(REPO / "tests" / "golden" / "lj_10x10x10_csr_half_100.npz")
    deck = deck_for(tmp_path, g["region"], g["nsteps"], 0)
    md = OracleMD.from_deck(deck, "CSR", "NEIGH_HALF")
    done = 0
    for row in g["thermo"]:  # step T PE ETot as printed with %lf by src/examinimd.cpp:159,259
        md.step(int(row[0]) - done)
        done = int(row[0])
        T, PE, KE = md.thermo()
        assert (f"{T:.6f}", f"{PE:.6f}", f"{PE + KE:.6f}") == (f"{row[1]:.6f}", f"{row[2]:.6f}", f"{row[3]:.6f}")
    md.close()


@pytest.mark.skipif(not REF_EXE.exists(), reason="oracle/_ref/ExaMiniMD_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("neigh,iteration,newton", [("CSR", "NEIGH_FULL", 0), ("2D", "NEIGH_HALF", 0), ("CSR", "NEIGH_HALF", 1)])
def test_oracle_vs_live_reference_run(oracle_lib, tmp_path, neigh, iteration, newton):
    import make_golden
    region, nsteps, steps = (8, 7, 6), 25, (0, 19, 20, 25)
    g = make_golden.run_reference(region, nsteps, "on" if newton else "off", neigh, iteration, steps, exe=REF_EXE)
    md = OracleMD.from_deck(deck_for(tmp_path, region, nsteps, newton), neigh, iteration)
    check_against(md, g, steps)
    md.close()
