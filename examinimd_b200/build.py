"""In-tree build of libemd_b200.so (CUDA kernels + C ABI + C++ host layer) and the ExaMiniMD driver.

    python -m examinimd_b200.build [--force] [--verbose]

nvcc cross-compiles sm_100a without a GPU; objects are cached under examinimd_b200/build/ and
rebuilt when a source or any header is newer.  Outputs (git-ignored, shipped to the GPU box by
gpurun): examinimd_b200/lib/libemd_b200.so, examinimd_b200/bin/ExaMiniMD.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REPO = ROOT.parent
CSRC = ROOT / "csrc"
BUILD = ROOT / "build"
LIB = ROOT / "lib" / "libemd_b200.so"
EXE = ROOT / "bin" / "ExaMiniMD"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
INCLUDES = ["-I", str(REPO / "include"), "-I", str(CSRC / "host"), "-I", str(CSRC / "kernels")]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")
    return exe


def host_compiler_flags() -> list[str]:
    # the image exports CC=/opt/gcc/bin/gcc; the distro g++ is the one with a complete runtime
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    return ["-ccbin", cxx] if cxx else []


def sources() -> list[Path]:
    return sorted((CSRC / "kernels").glob("*.cu")) + sorted(p for p in (CSRC / "host").rglob("*.cpp") if p.name != "main.cpp")


def headers_mtime() -> float:
    hs = list(CSRC.rglob("*.h")) + list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.inc")) + list((REPO / "include").glob("*.h"))
    return max(p.stat().st_mtime for p in hs)


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd[:6]) + " ...")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    LIB.parent.mkdir(exist_ok=True)
    EXE.parent.mkdir(exist_ok=True)
    hm = headers_mtime()
    cc = nvcc()
    ccbin = host_compiler_flags()
    jobs = []
    objs = []
    for src in sources():
        obj = BUILD / (src.relative_to(CSRC).as_posix().replace("/", "__") + ".o")
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hm):
            cmd = [cc, *ccbin, *ARCH, *NVCC_FLAGS, *INCLUDES, "-c", str(src), "-o", str(obj)]
            jobs.append(cmd)
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda c: _run(c, verbose), jobs))
    if force or jobs or not LIB.exists():
        _run([cc, *ccbin, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"], verbose)
    main_src = CSRC / "host" / "main.cpp"
    if force or jobs or not EXE.exists() or EXE.stat().st_mtime < main_src.stat().st_mtime:
        _run([cc, *ccbin, *ARCH, *NVCC_FLAGS, *INCLUDES, str(main_src), "-o", str(EXE), "-L", str(LIB.parent), "-lemd_b200",
              "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../lib"], verbose)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
