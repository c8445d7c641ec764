"""examinimd_b200 -- B200-native ExaMiniMD hot path.

The product is ``lib/libemd_b200.so`` (hand-written sm_100a CUDA kernels behind the C ABI of
``include/emd_b200.h`` + the C++ host classes mirroring ExaMiniMD's module API) and the
``bin/ExaMiniMD`` driver.  This Python package is only the ctypes binding that tests and
``bench.py`` use; PyTorch is used there for device buffers and ``torch.distributed`` plumbing.

There is no CPU fallback: :func:`lib` raises if the CUDA library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")  # load all kernels with the module, not at their first launch (if CUDA is not up yet)
from pathlib import Path

_ROOT = Path(__file__).resolve().parent
REPO = _ROOT.parent
LIB_PATH = _ROOT / "lib" / "libemd_b200.so"
EXE_PATH = _ROOT / "bin" / "ExaMiniMD"

_lib = None


class EmdError(RuntimeError):
    pass


class BinGeom(C.Structure):
    """emd_bin_geom (include/emd_b200.h)."""

    _fields_ = [("nbinx", C.c_int), ("nbiny", C.c_int), ("nbinz", C.c_int), ("nhalo", C.c_int),
                ("minx", C.c_double), ("maxx", C.c_double), ("miny", C.c_double), ("maxy", C.c_double),
                ("minz", C.c_double), ("maxz", C.c_double)]

    @property
    def nbins(self) -> int:
        return self.nbinx * self.nbiny * self.nbinz


class NeighList(C.Structure):
    """emd_neigh_list (include/emd_b200.h)."""

    _fields_ = [("d_row_map", C.c_void_p), ("d_num_neighs", C.c_void_p), ("d_neighs", C.c_void_p), ("stride", C.c_int)]


class SnapParams(C.Structure):
    """emd_snap_params (include/emd_b200.h)."""

    _fields_ = [("twojmax", C.c_int), ("switchflag", C.c_int), ("ntypes", C.c_int), ("nelements", C.c_int), ("ncoeffall", C.c_int),
                ("rcutfac", C.c_double), ("rfac0", C.c_double), ("rmin0", C.c_double), ("wself", C.c_double),
                ("elem_of_type", C.c_int * 12), ("radelem", C.c_void_p), ("wjelem", C.c_void_p), ("coeffelem", C.c_void_p)]


class Decomp(C.Structure):
    """emd_decomp (include/emd_b200.h)."""

    _fields_ = [("nranks", C.c_int), ("rank", C.c_int), ("grid", C.c_int * 3), ("pos", C.c_int * 3), ("neighbor_send", C.c_int * 6),
                ("neighbor_recv", C.c_int * 6), ("sub", C.c_double * 3), ("sub_lo", C.c_double * 3), ("sub_hi", C.c_double * 3)]


_P = C.c_void_p
_D3 = C.c_double * 3
_SIGS = {
    # name: (restype, argtypes)
    "emd_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int, _P]),
    "emd_ctx_destroy": (None, [_P]),
    "emd_ctx_stream": (_P, [_P]),
    "emd_ctx_sync": (C.c_int, [_P]),
    "emd_last_error": (C.c_char_p, []),
    "emd_microbench_fp64": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "emd_abi_version": (C.c_int, []),
    "emd_ctx_launch_count": (C.c_ulonglong, [_P]),
    "emd_ctx_tic": (C.c_int, [_P]),
    "emd_ctx_toc": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "emd_event_create": (C.c_int, [C.POINTER(_P)]),
    "emd_event_destroy": (C.c_int, [_P]),
    "emd_event_record": (C.c_int, [_P, _P]),
    "emd_event_elapsed_ms": (C.c_int, [_P, _P, C.POINTER(C.c_float)]),
    "emd_malloc": (C.c_int, [C.POINTER(_P), C.c_ulonglong]),
    "emd_free": (C.c_int, [_P]),
    "emd_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_ulonglong]),
    "emd_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_ulonglong]),
    "emd_memcpy_d2d": (C.c_int, [_P, _P, _P, C.c_ulonglong]),
    "emd_memset_zero": (C.c_int, [_P, _P, C.c_ulonglong]),
    "emd_binning_geometry": (C.c_int, [_D3, _D3, _D3, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(BinGeom)]),
    "emd_binning_build": (C.c_int, [_P, _P, C.c_int, C.POINTER(BinGeom), _P, _P, _P]),
    "emd_binning_permute": (C.c_int, [_P, _P, C.c_int] + [_P] * 12),
    "emd_neigh_csr_count": (C.c_int, [_P, _P, C.c_int, C.POINTER(BinGeom), _P, _P, _P, C.c_double, C.c_int, C.c_int, _P,
                                      C.POINTER(C.c_int)]),
    "emd_neigh_csr_fill": (C.c_int, [_P, _P, C.c_int, C.POINTER(BinGeom), _P, _P, _P, C.c_double, C.c_int, C.c_int, _P, _P]),
    "emd_neigh_2d_fill": (C.c_int, [_P, _P, C.c_int, C.POINTER(BinGeom), _P, _P, _P, C.c_double, C.c_int, C.c_int, C.c_int,
                                    _P, _P, C.POINTER(C.c_int)]),
    "emd_force_lj_set_params": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "emd_force_lj_compute": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.POINTER(NeighList), C.c_int, C.c_int]),
    "emd_force_lj_idial_compute": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.POINTER(NeighList), C.c_int, C.c_int, _P]),
    "emd_force_lj_energy": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(NeighList), C.c_int, C.POINTER(C.c_double)]),
    "emd_lattice_count": (C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    "emd_lattice_fill": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "emd_velocity_sums": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "emd_velocity_shift": (C.c_int, [_P, _P, C.c_int, C.c_double, C.c_double, C.c_double]),
    "emd_velocity_scale": (C.c_int, [_P, _P, C.c_int, C.c_double]),
    "emd_tiles_create": (C.c_int, [C.POINTER(_P)]),
    "emd_tiles_destroy": (None, [_P]),
    "emd_tiles_valid": (C.c_int, [_P]),
    "emd_tiles_invalidate": (None, [_P]),
    "emd_tiles_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int)]),
    "emd_tiles_lists": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "emd_neigh_tiles_build": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.POINTER(BinGeom), _P, _P, _P, C.c_double]),
    "emd_neigh_tiles_count": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.POINTER(C.c_int)]),
    "emd_neigh_tiles_fill_csr": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "emd_neigh_tiles_fill_2d": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.POINTER(C.c_int)]),
    "emd_force_lj_compute_tiles": (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(C.c_double)]),
    "emd_force_lj_compute_tiles_with_energy": (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(C.c_double)]),
    "emd_force_lj_compute_tiles_part": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int]),
    "emd_force_lj_compute_tiles_nve": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_double, C.c_double]),
    "emd_force_lj_compute_tiles_nve_thermo": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "emd_force_lj_compute_tiles_part_nve": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P, C.c_double, C.c_double]),
    "emd_tiles_complete": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "emd_tiles_halo_split": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "emd_ctx_side_mark": (C.c_int, [_P]),
    "emd_ctx_side_begin": (C.c_int, [_P]),
    "emd_ctx_side_end": (C.c_int, [_P]),
    "emd_ctx_side_join": (C.c_int, [_P]),
    "emd_ctx_side_sms": (C.c_int, [_P]),
    "emd_snap_create": (C.c_int, [C.POINTER(_P), C.POINTER(SnapParams)]),
    "emd_snap_destroy": (None, [_P]),
    "emd_snap_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "emd_snap_device_ptr": (_P, [_P, C.c_char_p]),
    "emd_force_snap_compute": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.POINTER(NeighList)]),
    "emd_snap_yi_plan_stats": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "emd_force_snap_energy": (C.c_int, [_P, _P, _P, _P, C.c_int, C.POINTER(NeighList), C.c_int, C.POINTER(C.c_double)]),
    "emd_nve_initial_integrate": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_double, C.c_double]),
    "emd_nve_final_initial_integrate": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_double, C.c_double]),
    "emd_nve_final_integrate": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_double]),
    "emd_comm_wrap": (C.c_int, [_P, _P, C.c_int, _D3]),
    "emd_comm_halo_phase": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _D3, _D3, _D3,
                                      C.c_double, C.POINTER(C.c_int)]),
    "emd_comm_halo_update_phase": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _D3]),
    "emd_comm_halo_resolve": (C.c_int, [_P, _P * 6, C.c_int * 6, C.c_int, _D3, _P, _P]),
    "emd_comm_halo_refresh": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "emd_comm_force_fold_phase": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "emd_comm_decompose": (C.c_int, [C.c_int, C.c_int, _D3, C.POINTER(Decomp)]),
    "emd_comm_wrap_dims": (C.c_int, [_P, _P, C.c_int, _D3, C.c_int * 3]),
    "emd_comm_exchange_pack": (C.c_int, [_P, C.c_int, C.POINTER(Decomp), _D3, _P, _P, _P, _P, _P, C.c_int, _P, C.c_int, C.POINTER(C.c_int)]),
    "emd_comm_halo_pack": (C.c_int, [_P, C.c_int, C.POINTER(Decomp), _D3, C.c_double, _P, _P, _P, _P, _P, C.c_int, _P, _P, C.c_int,
                                     C.POINTER(C.c_int)]),
    "emd_comm_unpack": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "emd_comm_exchange_compact": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int]),
    "emd_comm_halo_update_pack": (C.c_int, [_P, C.c_int, C.POINTER(Decomp), _D3, _P, _P, C.c_int, _P]),
    "emd_comm_halo_update_unpack": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "emd_comm_force_unpack": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "emd_net_unique_id": (C.c_int, [_P]),
    "emd_net_create": (C.c_int, [C.POINTER(_P), _P, C.c_int, C.c_int, _P]),
    "emd_net_destroy": (None, [_P]),
    "emd_net_sendrecv": (C.c_int, [_P, _P, C.c_ulonglong, C.c_int, _P, C.c_ulonglong, C.c_int]),
    "emd_net_group_begin": (C.c_int, [_P]),
    "emd_net_group_end": (C.c_int, [_P]),
    "emd_net_exchange_count": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "emd_net_allreduce": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int]),
    "emd_net_scan_int": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "emd_net_barrier": (C.c_int, [_P]),
    "emd_net_allgather_bytes": (C.c_int, [_P, _P, C.c_int, _P]),
    "emd_net_exchange_counts2": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "emd_peer_create": (C.c_int, [C.POINTER(_P), _P, _P, C.c_int, C.c_int]),
    "emd_peer_destroy": (None, [_P]),
    "emd_peer_publish": (C.c_int, [_P, _P, _P, _P, C.POINTER(C.c_int)]),
    "emd_peer_ready": (C.c_int, [_P]),
    "emd_peer_begin_update": (C.c_int, [_P, _P, _P]),
    "emd_peer_update_dim": (C.c_int, [_P, _P, C.POINTER(C.c_double), C.c_int, _P, _P, C.c_int, _P, C.c_int, C.c_int]),
    "emd_peer_wait_all": (C.c_int, [_P]),
    "emd_ctx_set_halo_gate": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "emd_ctx_halo_gate_wait": (C.c_int, [_P]),
    "emd_ctx_halo_gate_pending": (C.c_int, [_P]),
    "emd_reduce_mv2": (C.c_int, [_P, _P, _P, _P, C.c_int, C.POINTER(C.c_double)]),
    # session API (include/emd_b200_app.h)
    "emd_app_create": (C.c_int, [C.POINTER(_P), C.c_int, C.POINTER(C.c_char_p), C.c_int, _P]),
    "emd_app_destroy": (None, [_P]),
    "emd_app_ctx": (_P, [_P]),
    "emd_app_advance": (C.c_int, [_P, C.c_int]),
    "emd_app_run": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "emd_app_advance_timed": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "emd_app_thermo": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "emd_app_get": (C.c_longlong, [_P, C.c_char_p]),
    "emd_app_download": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "emd_app_upload": (C.c_int, [_P, _P, _P, _P]),
    "emd_app_device_ptr": (_P, [_P, C.c_char_p]),
    "emd_app_neigh_stride": (C.c_int, [_P]),
    "emd_app_dump_binary": (C.c_int, [_P, C.c_char_p, C.c_int]),
}


def lib() -> C.CDLL:
    """Load libemd_b200.so (built by ``python -m examinimd_b200.build``).  No fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise EmdError(f"{LIB_PATH} is missing: run `python -m examinimd_b200.build` (there is no CPU fallback)")
        L = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else C.DEFAULT_MODE)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def declared_symbols() -> list[str]:
    return sorted(_SIGS)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise EmdError(f"{what} failed (rc={rc}): {lib().emd_last_error().decode(errors='replace')}")


def vec3(a) -> "C.Array":
    return _D3(float(a[0]), float(a[1]), float(a[2]))


class Context:
    """RAII wrapper of emd_ctx bound to a CUDA stream (default: torch's current stream)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = _P()
        check(lib().emd_ctx_create(C.byref(self._h), device, _P(stream) if stream else None), "emd_ctx_create")

    @property
    def handle(self):
        return self._h

    def sync(self):
        check(lib().emd_ctx_sync(self._h), "emd_ctx_sync")

    @property
    def launches(self) -> int:
        return int(lib().emd_ctx_launch_count(self._h))

    def close(self):
        if self._h:
            lib().emd_ctx_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class App:
    """The whole application through the session C API (emd_b200_app.h): ExaMiniMD::init + run loop."""

    def __init__(self, argv: list[str], device: int = 0, stream: int | None = None):
        self._h = _P()
        arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
        check(lib().emd_app_create(C.byref(self._h), len(argv), arr, device, _P(stream) if stream else None), "emd_app_create")

    def get(self, what: str) -> int:
        return int(lib().emd_app_get(self._h, what.encode()))

    @property
    def ctx(self):
        return _P(lib().emd_app_ctx(self._h))

    def advance(self, nsteps: int) -> None:
        check(lib().emd_app_advance(self._h, nsteps), "emd_app_advance")

    def run(self, nsteps: int):
        """the reference's run loop: steps + thermo reductions at the deck's cadence; returns the last (T, PE, KE) or None"""
        out = (C.c_double * 3)(float("nan"), float("nan"), float("nan"))
        check(lib().emd_app_run(self._h, nsteps, out), "emd_app_run")
        return None if out[0] != out[0] else (out[0], out[1], out[2])

    def advance_timed(self, nsteps: int) -> dict:
        """advance with the reference's phase timers; seconds per phase over the nsteps"""
        out = (C.c_double * 4)()
        check(lib().emd_app_advance_timed(self._h, nsteps, out), "emd_app_advance_timed")
        return {"force": out[0], "neigh": out[1], "comm": out[2], "other": out[3]}

    def sync(self) -> None:
        check(lib().emd_ctx_sync(self.ctx), "emd_ctx_sync")

    def thermo(self) -> tuple[float, float, float]:
        T, PE, KE = C.c_double(), C.c_double(), C.c_double()
        check(lib().emd_app_thermo(self._h, C.byref(T), C.byref(PE), C.byref(KE)), "emd_app_thermo")
        return T.value, PE.value, KE.value

    def download(self):
        import numpy as np
        n = self.get("N_local")
        out = {"id": np.empty(n, np.int32), "type": np.empty(n, np.int32), "q": np.empty(n), "x": np.empty((n, 3)),
               "v": np.empty((n, 3)), "f": np.empty((n, 3))}
        check(lib().emd_app_download(self._h, *[out[k].ctypes.data_as(_P) for k in ("id", "type", "q", "x", "v", "f")]),
              "emd_app_download")
        return out

    def device_ptr(self, what: str) -> int:
        p = lib().emd_app_device_ptr(self._h, what.encode())
        return int(p) if p else 0

    def launches(self) -> int:
        return int(lib().emd_ctx_launch_count(self.ctx))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            lib().emd_app_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
