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
