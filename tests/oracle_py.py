"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- test infrastructure only."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
ODIR = REPO / "oracle"
_lib = None

NEIGH = {"CSR": 1, "CSR_MAPCONSTR": 2, "2D": 3}
ITER = {"NEIGH_FULL": 1, "NEIGH_HALF": 2}


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    so = ODIR / "liboracle.so"
    srcs = list(ODIR.glob("*.c")) + list(ODIR.glob("*.h"))
    if not so.exists() or so.stat().st_mtime < max(p.stat().st_mtime for p in srcs):
        subprocess.run(["make", "-C", str(ODIR), "liboracle.so"], check=True, capture_output=True)
    L = C.CDLL(str(so))
    P = C.c_void_p
    L.orcf_create_deck.restype = P
    L.orcf_create_deck.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.orcf_create_raw.restype = P
    L.orcf_create_raw.argtypes = [C.c_int, P, P, P, C.c_int, P, P, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                  C.c_double, C.c_double, C.c_int]
    L.orcf_destroy.argtypes = [P]
    L.orcf_setup.argtypes = [P]
    L.orcf_step.argtypes = [P, C.c_int]
    L.orcf_thermo.argtypes = [P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.orcf_stage.argtypes = [P, C.c_char_p]
    L.orcf_stage.restype = C.c_int
    L.orcf_get_int.argtypes = [P, C.c_char_p]
    L.orcf_get_int.restype = C.c_longlong
    L.orcf_get_double.argtypes = [P, C.c_char_p]
    L.orcf_get_double.restype = C.c_double
    L.orcf_copy.argtypes = [P, C.c_char_p, P]
    L.orcf_copy.restype = C.c_longlong
    L.orcf_set.argtypes = [P, C.c_char_p, P, C.c_int]
    L.orcf_dump.argtypes = [P, C.c_char_p, C.c_int]
    L.orcf_snap_probe.argtypes = [P, C.c_int, P, P, P, P]
    L.orcf_snap_probe.restype = C.c_int
    L.orcf_snap_jdim.argtypes = [P]
    L.orcf_snap_jdim.restype = C.c_int
    _lib = L
    return L


_DTYPES = {"x": (np.float64, 3), "v": (np.float64, 3), "f": (np.float64, 3), "q": (np.float64, 1), "mass": (np.float64, 1),
           "type": (np.int32, 1), "id": (np.int32, 1), "bincount": (np.int32, 1), "binoffsets": (np.int32, 1),
           "permute": (np.int32, 1), "row_map": (np.int32, 1), "entries": (np.int32, 1), "num_neighs": (np.int32, 1),
           "neighs2d": (np.int32, 1)}


class OracleMD:
    """One oracle simulation (orc_md) driven stage by stage."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle creation failed")
        self.h = C.c_void_p(handle)
        self.L = load()

    @classmethod
    def from_deck(cls, deck, neigh="CSR", iteration="NEIGH_FULL", region=None, coeff_dir=None, setup=True):
        L = load()
        r = region or (0, 0, 0)
        h = L.orcf_create_deck(str(deck).encode(), NEIGH[neigh], ITER[iteration], r[0], r[1], r[2],
                               str(coeff_dir).encode() if coeff_dir else None, 1 if setup else 0)
        return cls(h)

    @classmethod
    def from_arrays(cls, x, box, force_cutoff=2.5, skin=0.3, v=None, types=None, ntypes=1, mass=None, dt=0.005, newton=0,
                    neigh="CSR", iteration="NEIGH_FULL", eps=1.0, sigma=1.0, exchange_rate=20):
        L = load()
        x = np.ascontiguousarray(x, np.float64)
        n = x.shape[0]
        v_ = np.ascontiguousarray(v, np.float64) if v is not None else None
        t_ = np.ascontiguousarray(types, np.int32) if types is not None else None
        m_ = np.ascontiguousarray(mass if mass is not None else np.ones(ntypes), np.float64)
        b_ = np.ascontiguousarray(box, np.float64)
        h = L.orcf_create_raw(n, x.ctypes.data, v_.ctypes.data if v_ is not None else None,
                              t_.ctypes.data if t_ is not None else None, ntypes, m_.ctypes.data, b_.ctypes.data,
                              force_cutoff, skin, dt, newton, NEIGH[neigh], ITER[iteration], eps, sigma, exchange_rate)
        return cls(h)

    def geti(self, w):
        return int(self.L.orcf_get_int(self.h, w.encode()))

    def getd(self, w):
        return float(self.L.orcf_get_double(self.h, w.encode()))

    def stage(self, *names):
        for nm in names:
            if self.L.orcf_stage(self.h, nm.encode()) != 0:
                raise ValueError(nm)

    def setup(self):
        self.L.orcf_setup(self.h)

    def step(self, n=1):
        self.L.orcf_step(self.h, n)

    def thermo(self):
        T, PE, KE = C.c_double(), C.c_double(), C.c_double()
        self.L.orcf_thermo(self.h, C.byref(T), C.byref(PE), C.byref(KE))
        return T.value, PE.value, KE.value

    def arr(self, w):
        nbytes = self.L.orcf_copy(self.h, w.encode(), None)
        if nbytes < 0:
            raise KeyError(w)
        key = "permute" if w == "permute" else ("permute" if w.startswith("pack") else w)
        dt, width = _DTYPES[key]
        out = np.empty(nbytes // np.dtype(dt).itemsize, dt)
        if nbytes:
            self.L.orcf_copy(self.h, w.encode(), out.ctypes.data)
        return out.reshape(-1, 3) if width == 3 else out

    def set(self, w, a):
        a = np.ascontiguousarray(a, np.float64)
        self.L.orcf_set(self.h, w.encode(), a.ctypes.data, a.shape[0])

    def geom(self):
        return {k: self.geti(k) for k in ("nbinx", "nbiny", "nbinz", "nhalo")} | {k: self.getd(k) for k in
                                                                                 ("minx", "maxx", "miny", "maxy", "minz", "maxz")}

    def rows(self):
        """neighbor rows as a list of np arrays (CSR or 2D)."""
        n = self.geti("N_local")
        if self.geti("total_neighs") > 0 or self.L.orcf_copy(self.h, b"row_map", None) > 4:
            rm = self.arr("row_map")
            if rm.size == n + 1 and rm[-1] == self.geti("total_neighs") and rm[-1] > 0:
                e = self.arr("entries")
                return [e[rm[i]:rm[i + 1]] for i in range(n)]
        nn = self.arr("num_neighs")
        m = self.geti("maxneighs")
        t = self.arr("neighs2d").reshape(n, m)
        return [t[i, :nn[i]] for i in range(n)]

    def snap_probe(self, i, with_forces=True):
        """U_tot [jdim,jdim,jdim] complex (index j,ma,mb), in-cutoff neighbor indices and F_ij of atom i."""
        jd = self.L.orcf_snap_jdim(self.h)
        ur, ui = np.zeros(jd ** 3), np.zeros(jd ** 3)
        inside, fij = np.zeros(512, np.int32), np.zeros((512, 3))
        n = self.L.orcf_snap_probe(self.h, i, ur.ctypes.data, ui.ctypes.data, inside.ctypes.data, fij.ctypes.data if with_forces else None)
        if n < 0:
            raise RuntimeError("no SNAP force in this oracle")
        return (ur + 1j * ui).reshape(jd, jd, jd), inside[:n].copy(), fij[:n].copy()

    def close(self):
        if self.h:
            self.L.orcf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------ whole CPU runs (binaries)
REF_OMP = ODIR / "_ref" / "ExaMiniMD_ref_omp"   # the unmodified reference over the OpenMP Kokkos stand-in (oracle/Makefile.ref)
ORACLE_OMP = ODIR / "oracle_md_omp"             # OpenMP build of the restatement


def cpu_run_dump(deck_text, steps, neigh="CSR", iteration="NEIGH_HALF", threads=None, files=(), exe=None):
    """One whole CPU run of `deck_text` (restricted LAMMPS deck) with --dumpbinary at `steps`: the reference binary itself when
    it was built (oracle/_ref), else the OpenMP restatement.  Returns {step: {id, x, v, f}} and the stdout."""
    import os
    import sys
    import tempfile
    sys.path.insert(0, str(REPO / "tests" / "golden"))
    import make_golden
    exe = exe or (REF_OMP if REF_OMP.exists() else ORACLE_OMP)
    if not Path(exe).exists():
        subprocess.run(["make", "-C", str(ODIR), "oracle_md_omp"], check=True, capture_output=True)
    threads = threads or os.cpu_count() or 1
    rate = steps[0] if len(steps) == 1 else int(np.gcd.reduce(steps))
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "in.deck").write_text(deck_text)
        for f in files:
            (td / Path(f).name).write_bytes(Path(f).read_bytes())
        (td / "dump").mkdir()
        env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="close", OMP_PLACES="cores")
        r = subprocess.run([str(exe), "-il", "in.deck", "--comm-type", "SERIAL", "--neigh-type", neigh, "--force-iteration", iteration,
                            "--dumpbinary", str(max(rate, 1)), "dump"], capture_output=True, text=True, check=True, cwd=td, env=env)
        out = {}
        for s_ in steps:
            d = make_golden.read_dump(td / "dump" / ("output.%010d.000" % s_))
            out[s_] = {k: np.array(d[k]) for k in ("id", "x", "v", "f")}
    return out, r.stdout


def lj_total_energy(x, v, box, iteration="NEIGH_HALF", mass=2.0):
    """PE + KE per atom (full precision) of an LJ state of the in.lj deck, evaluated by the serial oracle"""
    md = OracleMD.from_arrays(x, box, v=v, mass=[mass], neigh="CSR", iteration=iteration)
    md.setup()
    _, pe, ke = md.thermo()
    md.close()
    return pe + ke


def lj_energy_envelope(region, nsteps, iteration="NEIGH_HALF"):
    """The reference's own run-to-run envelope of the total energy after `nsteps` (north_star): the OpenMP restatement run with
    1 thread and with all cores (atomic force accumulation changes the summation order).  Returns (E_1thread, E_allcores, E_0)."""
    import os
    import re
    txt = (REPO / "input" / "in.lj").read_text()
    txt = re.sub(r"region\s+box block.*", "region\t\tbox block 0 %d 0 %d 0 %d" % tuple(region), txt)
    txt = re.sub(r"run\s+\d+", "run\t\t%d" % nsteps, txt)
    a = (4.0 / 0.8442) ** (1.0 / 3.0)
    box = [r * a for r in region]
    one, _ = cpu_run_dump(txt, [nsteps], iteration=iteration, threads=1, exe=ORACLE_OMP)
    many, _ = cpu_run_dump(txt, [nsteps], iteration=iteration, threads=max(2, os.cpu_count() or 2), exe=ORACLE_OMP)
    e1 = lj_total_energy(one[nsteps]["x"], one[nsteps]["v"], box, iteration)
    en = lj_total_energy(many[nsteps]["x"], many[nsteps]["v"], box, iteration)
    return e1, en
