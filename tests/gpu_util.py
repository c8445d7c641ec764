"""Helpers for the -m gpu parity tests: torch owns the device buffers, every compute call goes
through the C ABI (examinimd_b200.lib())."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

import examinimd_b200 as emd

P = C.c_void_p


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def ptr(t):
    return P(t.data_ptr()) if t is not None else None


def new_ctx():
    # torch's default stream has handle 0, which the C ABI reads as "create a private stream"; pass
    # cudaStreamLegacy (0x1) instead so the kernels are ordered with torch's own copies/allocations
    s = torch.cuda.current_stream().cuda_stream
    return emd.Context(torch.cuda.current_device(), s if s else 1)


def geom_from(d):
    g = emd.BinGeom()
    for k, v in d.items():
        setattr(g, k, v)
    return g


def binning_build(ctx, x_dev, n, g):
    L = emd.lib()
    bc = torch.empty(g.nbins, dtype=torch.int32, device="cuda")
    bo = torch.empty(g.nbins, dtype=torch.int32, device="cuda")
    pv = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
    emd.check(L.emd_binning_build(ctx.handle, ptr(x_dev), n, C.byref(g), ptr(bc), ptr(bo), ptr(pv)), "emd_binning_build")
    return bc, bo, pv[:n]


def neigh_csr(ctx, x_dev, n_local, g, bc, bo, pv, cut, half, newton):
    L = emd.lib()
    rm = torch.empty(n_local + 1, dtype=torch.int32, device="cuda")
    total = C.c_int()
    emd.check(L.emd_neigh_csr_count(ctx.handle, ptr(x_dev), n_local, C.byref(g), ptr(bc), ptr(bo), ptr(pv), cut, half, newton,
                                    ptr(rm), C.byref(total)), "emd_neigh_csr_count")
    ent = torch.empty(max(total.value, 1), dtype=torch.int32, device="cuda")
    emd.check(L.emd_neigh_csr_fill(ctx.handle, ptr(x_dev), n_local, C.byref(g), ptr(bc), ptr(bo), ptr(pv), cut, half, newton,
                                   ptr(rm), ptr(ent)), "emd_neigh_csr_fill")
    return rm, ent[: total.value], total.value


def neigh_2d(ctx, x_dev, n_local, g, bc, bo, pv, cut, half, newton, maxneighs=16):
    L = emd.lib()
    passes = 0
    while True:
        nn = torch.empty(n_local + 1, dtype=torch.int32, device="cuda")
        tab = torch.empty((n_local + 1, maxneighs), dtype=torch.int32, device="cuda")
        mx = C.c_int()
        emd.check(L.emd_neigh_2d_fill(ctx.handle, ptr(x_dev), n_local, C.byref(g), ptr(bc), ptr(bo), ptr(pv), cut, half, newton,
                                      maxneighs, ptr(nn), ptr(tab), C.byref(mx)), "emd_neigh_2d_fill")
        passes += 1
        if mx.value <= maxneighs:
            return nn[:n_local], tab, maxneighs, passes
        maxneighs = int(mx.value * 1.2)


def csr_list(rm, ent):
    return emd.NeighList(ptr(rm), None, ptr(ent), 1)


def table_list(nn, tab, stride):
    return emd.NeighList(None, ptr(nn), ptr(tab), stride)


class Tiles:
    """emd_tiles built from an oracle state (the B200 fast path of kernels/tiles.cu)."""

    def __init__(self, ctx, x_dev, n_local, n_all, g, bc, bo, pv, cut):
        L = emd.lib()
        self.ctx, self.n_local = ctx, n_local
        self.keep = (x_dev, bc, bo, pv)  # the build captures these pointers
        self.h = P()
        emd.check(L.emd_tiles_create(C.byref(self.h)), "emd_tiles_create")
        self.rc = L.emd_neigh_tiles_build(ctx.handle, self.h, ptr(x_dev), n_local, n_all, C.byref(g), ptr(bc), ptr(bo), ptr(pv), cut)
        if self.rc not in (0, 3):
            emd.check(self.rc, "emd_neigh_tiles_build")

    @property
    def ok(self):
        return self.rc == 0 and emd.lib().emd_tiles_valid(self.h) == 1

    def info(self):
        d = (C.c_int * 3)()
        nt, st, mr, cap = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        emd.check(emd.lib().emd_tiles_info(self.h, d, C.byref(nt), C.byref(st), C.byref(mr), C.byref(cap)))
        return {"dims": tuple(d), "ntiles": nt.value, "stride": st.value, "maxrow": mr.value, "cap": cap.value}

    def lists(self):
        """host copies of the tile lists: exact rows (csr16/ncsr, None before a CSR/2D list was asked for) and the force
        kernel's rows (ell_s/nell_s), both [ntiles][words][stride][8] uint16, + own slots and the slot -> atom table"""
        inf = self.info()
        csr16, ncsr, ell_s, nell_s, islot, stgj = P(), P(), P(), P(), P(), P()
        emd.check(emd.lib().emd_tiles_lists(self.h, C.byref(csr16), C.byref(ncsr), C.byref(ell_s), C.byref(nell_s), C.byref(islot), C.byref(stgj)))
        nt, st, mr, cap = inf["ntiles"], inf["stride"], inf["maxrow"], inf["cap"]

        def grab(p, count, dtype):
            h = np.empty(count, dtype=dtype)
            emd.check(emd.lib().emd_memcpy_d2h(self.ctx.handle, h.ctypes.data_as(P), p, h.nbytes), "emd_memcpy_d2h")
            return h

        out = {"maxrow": mr, "stride": st, "ntiles": nt, "cap": cap}
        if csr16:
            out["csr16"] = grab(csr16, nt * mr * st, np.uint16).reshape(nt, mr // 8, st, 8)
            out["ncsr"] = grab(ncsr, nt * st, np.int32).reshape(nt, st)
        out["ell_s"] = grab(ell_s, nt * mr * st, np.uint16).reshape(nt, mr // 8, st, 8)
        out["nell_s"] = grab(nell_s, nt * st, np.int32).reshape(nt, st)
        out["int_slot"] = grab(islot, nt * st, np.uint16).reshape(nt, st)
        out["stg_j"] = grab(stgj, nt * cap, np.int32).reshape(nt, cap)
        return out

    def csr(self, half, newton):
        L = emd.lib()
        rm = torch.empty(self.n_local + 1, dtype=torch.int32, device="cuda")
        total = C.c_int()
        rc = L.emd_neigh_tiles_count(self.ctx.handle, self.h, half, newton, ptr(rm), C.byref(total))
        if rc == 3:  # the build is asynchronous: "fast path not applicable" may only show when the flags are read back
            self.rc = 3
            return None, None, -1
        emd.check(rc, "tiles_count")
        ent = torch.empty(max(total.value, 1), dtype=torch.int32, device="cuda")
        emd.check(L.emd_neigh_tiles_fill_csr(self.ctx.handle, self.h, half, newton, ptr(rm), ptr(ent)), "tiles_fill_csr")
        return rm, ent[: total.value], total.value

    def table(self, half, newton, maxneighs=16):
        L = emd.lib()
        passes = 0
        while True:
            nn = torch.empty(self.n_local + 1, dtype=torch.int32, device="cuda")
            tab = torch.empty((self.n_local + 1, maxneighs), dtype=torch.int32, device="cuda")
            mx = C.c_int()
            emd.check(L.emd_neigh_tiles_fill_2d(self.ctx.handle, self.h, half, newton, maxneighs, ptr(nn), ptr(tab), C.byref(mx)))
            passes += 1
            if mx.value <= maxneighs:
                return nn[: self.n_local], tab, maxneighs, passes
            maxneighs = int(mx.value * 1.2)

    def force(self, x_dev, type_dev, f_dev, energy=False):
        pe = C.c_double()
        emd.check(emd.lib().emd_force_lj_compute_tiles(self.ctx.handle, self.h, ptr(x_dev), ptr(type_dev), ptr(f_dev),
                                                       C.byref(pe) if energy else None), "force_tiles")
        return pe.value

    def force_part(self, x_dev, type_dev, f_dev, part, reserve=0):
        emd.check(emd.lib().emd_force_lj_compute_tiles_part(self.ctx.handle, self.h, ptr(x_dev), ptr(type_dev), ptr(f_dev), part, reserve),
                  "force_tiles_part")

    def halo_split(self):
        a, b = C.c_int(), C.c_int()
        emd.check(emd.lib().emd_tiles_halo_split(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        if self.h:
            emd.lib().emd_tiles_destroy(self.h)
            self.h = P()
