#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (ECP-copa/ExaMiniMD) compiled over the
host-only Kokkos stand-in (oracle/Makefile.ref -> oracle/_ref/ExaMiniMD_ref, 1 thread).

    python tests/golden/make_golden.py          # needs /root/reference (this container only)

Each fixture is one run of the reference binary on a restricted-LAMMPS deck derived from the shipped
input/in.lj (only `region`, `run`, `newton` edited) with `--dumpbinary`: the per-step binary dumps
(src/examinimd.cpp:296-346: id, type, q, x, v, f of the owned atoms) and the thermo table it prints.
They pin the CPU oracle (tests/test_oracle_vs_reference.py, no GPU) and, through it and directly
(tests/test_gpu_golden.py), the CUDA path on the GPU box, where /root/reference does not exist."""
import re
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = REPO / "oracle" / "_ref" / "ExaMiniMD_ref"
DECK = REPO / "input" / "in.lj"

CASES = [  # name, region, nsteps, newton, neigh, iteration, dump steps
    ("lj_6x6x6_csr_half", (6, 6, 6), 40, "off", "CSR", "NEIGH_HALF", (0, 1, 20, 40)),
    ("lj_6x6x6_csr_full", (6, 6, 6), 40, "off", "CSR", "NEIGH_FULL", (0, 1, 20, 40)),
    ("lj_6x6x6_2d_half", (6, 6, 6), 40, "off", "2D", "NEIGH_HALF", (0, 20, 40)),
    ("lj_6x6x6_2d_full", (6, 6, 6), 40, "off", "2D", "NEIGH_FULL", (0, 20, 40)),
    ("lj_7x6x5_csr_half_newton", (7, 6, 5), 40, "on", "CSR", "NEIGH_HALF", (0, 20, 40)),
    ("lj_10x10x10_csr_half_100", (10, 10, 10), 100, "off", "CSR", "NEIGH_HALF", (0, 100)),
]


IDIAL_CASES = [  # pair_style lj/cut/idial (ForceLJIDialNeigh): name, region, nsteps, neigh, iteration, nrepeat, dump steps
    ("ljidial_6x6x6_csr_full_r3", (6, 6, 6), 40, "CSR", "NEIGH_FULL", 3, (0, 1, 20, 40)),
    ("ljidial_6x6x6_2d_half_r2", (6, 6, 6), 40, "2D", "NEIGH_HALF", 2, (0, 20, 40)),
]

SNAP_DIR = REPO / "input" / "snap"
SNAP_CASES = [  # name, deck, region, nsteps, neigh, dump steps  (newton on, full list: the only mode ForceSNAP accepts)
    ("snap_W_4x4x4_csr", "in.snap.W", (4, 4, 4), 4, "CSR", (0, 1, 4)),
    ("snap_W_4x5x6_2d", "in.snap.W", (4, 5, 6), 2, "2D", (0, 2)),
    ("snap_Ta06A_4x4x4_csr", "in.snap.Ta06A", (4, 4, 4), 4, "CSR", (0, 1, 4)),
]


def make_deck(path, region, nsteps, newton, deck=DECK, idial=0):
    txt = Path(deck).read_text()
    if idial:  # the reference ships no lj/cut/idial deck: in.lj with the pair lines of src/input.cpp:367-369 / force_lj_idial_neigh_impl.h:50-57
        txt = re.sub(r"pair_style\s+lj/cut\s+(\S+)", r"pair_style\tlj/cut/idial \1", txt)
        txt = re.sub(r"(pair_coeff\s+\S+\s+\S+\s+\S+\s+\S+\s+\S+)", r"\1 %d" % idial, txt)
    txt = re.sub(r"region\s+box block.*", "region\t\tbox block 0 %d 0 %d 0 %d" % region, txt)
    txt = re.sub(r"run\s+\d+", "run\t\t%d" % nsteps, txt)
    txt = re.sub(r"newton \w+", "newton %s" % newton, txt)
    path.write_text(txt)


def lattice_constant(deck):
    """cell edge of the deck's `lattice` command (src/input.cpp:312-330: sc a | fcc rho* -> (4/rho)^(1/3))"""
    m = re.search(r"^lattice\s+(\w+)\s+([\d.eE+-]+)", Path(deck).read_text(), re.M)
    return float(m.group(2)) if m.group(1) == "sc" else (4.0 / float(m.group(2))) ** (1.0 / 3.0)


def read_dump(p):
    raw = p.read_bytes()
    n = int(np.frombuffer(raw[:4], np.int32)[0])
    o, out = 4, {}
    for k, dt, w in (("id", np.int32, 1), ("type", np.int32, 1), ("q", np.float64, 1), ("x", np.float64, 3), ("v", np.float64, 3), ("f", np.float64, 3)):
        sz = n * w * np.dtype(dt).itemsize
        a = np.frombuffer(raw[o:o + sz], dt)
        out[k] = a.reshape(n, 3) if w == 3 else a
        o += sz
    assert o == len(raw)
    return out


def run_reference(region, nsteps, newton, neigh, iteration, dump_steps, exe=REF, deck_src=DECK, idial=0):
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        deck = td / "in.deck"
        make_deck(deck, region, nsteps, newton, deck_src, idial)
        for f in SNAP_DIR.glob("*.snap*"):  # coefficient files are opened relative to the working directory
            (td / f.name).write_bytes(f.read_bytes())
        (td / "dump").mkdir()
        r = subprocess.run([str(exe), "-il", str(deck), "--comm-type", "SERIAL", "--neigh-type", neigh, "--force-iteration", iteration,
                            "--dumpbinary", "1", str(td / "dump")], capture_output=True, text=True, check=True, cwd=td)
        thermo = np.array([[float(t) for t in l.split()[:4]] for l in r.stdout.splitlines() if re.match(r"^\d+ -?\d+\.\d+ ", l)])
        out = {"thermo": thermo, "region": np.array(region), "nsteps": np.array(nsteps)}
        for s in dump_steps:
            for k, v in read_dump(td / "dump" / ("output.%010d.000" % s)).items():
                if k in ("type", "q") and s != dump_steps[0]:
                    continue
                out["s%d_%s" % (s, k)] = v
        return out


def main():
    if not REF.exists():
        sys.exit("build the reference first: make -C oracle -f Makefile.ref (needs /root/reference)")
    for name, region, nsteps, newton, neigh, iteration, steps in CASES:
        out = run_reference(region, nsteps, newton, neigh, iteration, steps)
        out["newton"] = np.array(1 if newton == "on" else 0)
        out["neigh"] = np.array(neigh)
        out["iteration"] = np.array(iteration)
        np.savez_compressed(HERE / (name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith("s0_") or k == "thermo"})
    for name, region, nsteps, neigh, iteration, nrepeat, steps in IDIAL_CASES:
        out = run_reference(region, nsteps, "off", neigh, iteration, steps, idial=nrepeat)
        out["newton"] = np.array(0)
        out["neigh"] = np.array(neigh)
        out["iteration"] = np.array(iteration)
        out["idial"] = np.array(nrepeat)
        np.savez_compressed(HERE / (name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith("s0_") or k == "thermo"})
    for name, deck, region, nsteps, neigh, steps in SNAP_CASES:
        out = run_reference(region, nsteps, "on", neigh, "NEIGH_FULL", steps, deck_src=SNAP_DIR / deck)
        out["newton"] = np.array(1)
        out["neigh"] = np.array(neigh)
        out["iteration"] = np.array("NEIGH_FULL")
        out["deck"] = np.array(deck)
        np.savez_compressed(HERE / (name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith("s0_") or k == "thermo"})


if __name__ == "__main__":
    main()
