"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the headers
declare, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import re
import subprocess

import pytest

from oracle_py import REPO


def _declared(header):
    txt = (REPO / "include" / header).read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(emd_[a-z0-9_]+)\s*\(", txt)))


def test_headers_declare_and_library_exports(emd):
    names = _declared("emd_b200.h") + _declared("emd_b200_app.h")
    assert len(names) > 40
    out = subprocess.run(["nm", "-D", "--defined-only", str(emd.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (emd_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    # the ctypes table binds exactly the declared functions
    assert sorted(emd.declared_symbols()) == sorted(names)


def test_abi_version_and_struct_layout(emd):
    L = emd.lib()
    assert L.emd_abi_version() == 1
    assert C.sizeof(emd.BinGeom) == 4 * 4 + 6 * 8
    assert C.sizeof(emd.NeighList) == 3 * 8 + 8


def test_binning_geometry_matches_oracle(emd, oracle_lib):
    """host-side arithmetic of binning_kksort.cpp:77-99 through the C ABI vs the oracle, bit for bit."""
    from oracle_py import OracleMD
    md = OracleMD.from_deck(REPO / "input" / "in.lj", "CSR", "NEIGH_FULL", region=(7, 9, 11), setup=False)
    md.stage("exchange", "bin_sort")
    g = emd.BinGeom()
    box = [md.getd("domain_x"), md.getd("domain_y"), md.getd("domain_z")]
    c = md.getd("neigh_cutoff")
    rc = emd.lib().emd_binning_geometry(emd.vec3(box), emd.vec3([0, 0, 0]), emd.vec3(box), c, c, c, 1, C.byref(g))
    assert rc == 0
    og = md.geom()
    for k in ("nbinx", "nbiny", "nbinz", "nhalo", "minx", "maxx", "miny", "maxy", "minz", "maxz"):
        assert getattr(g, k) == og[k], k
    md.close()


def test_no_cpu_fallback(emd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = emd.lib().emd_ctx_create(C.byref(h), 0, None)
    assert rc != 0 and b"no CPU fallback" in emd.lib().emd_last_error()
    r = subprocess.run([str(emd.EXE_PATH), "-il", str(REPO / "input" / "in.lj")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_product_does_not_reference_oracle():
    """the product tree must not include, link or execute anything under oracle/."""
    bad = []
    for p in (REPO / "examinimd_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".cpp", ".h"} and "oracle" in p.read_text(errors="ignore").lower():
            bad.append(str(p))
    assert not bad, bad
