"""GPU parity tests (-m gpu): every CUDA kernel family, called through the C ABI, against the CPU
oracle on the same inputs.  Integer/index outputs (bins, permutations, neighbor rows, ghost lists)
must be bit-exact; FP64 forces/positions within 1e-10 of the global RMS;
the integrator is bit-exact given identical forces."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle_py import OracleMD, REPO

DECK = REPO / "input" / "in.lj"
TOL = 1e-10


@pytest.fixture(scope="module")
def gu():
    import gpu_util
    return gpu_util


@pytest.fixture(scope="module")
def ctx(emd, gu):
    c = gu.new_ctx()
    yield c
    c.close()


def liquid(nsteps=25, region=(9, 9, 9), neigh="CSR", iteration="NEIGH_FULL", newton_deck=None):
    """an oracle state off the lattice: nsteps of MD from the in.lj melt (rebuild at step 20)."""
    md = OracleMD.from_deck(newton_deck or DECK, neigh, iteration, region=region)
    md.step(nsteps)
    return md


def rebuilt(md):
    md.stage("exchange", "bin_sort", "halo", "bin_all", "neigh")
    return md


# ----------------------------------------------------------------------------- binning (a1)
@pytest.mark.parametrize("state", ["lattice", "liquid"])
def test_binning_bit_exact(emd, gu, ctx, state):
    md = OracleMD.from_deck(DECK, "CSR", "NEIGH_FULL", region=(10, 10, 10), setup=False) if state == "lattice" else liquid()
    md.stage("exchange")
    x0 = md.arr("x")[: md.geti("N_local")].copy()
    md.stage("bin_nosort")  # same permutation, atoms left in place
    g = gu.geom_from(md.geom())
    bc, bo, pv = gu.binning_build(ctx, gu.dev(x0), x0.shape[0], g)
    np.testing.assert_array_equal(bc.cpu().numpy(), md.arr("bincount"))
    np.testing.assert_array_equal(bo.cpu().numpy(), md.arr("binoffsets"))
    np.testing.assert_array_equal(pv.cpu().numpy(), md.arr("permute"))
    md.close()


def test_binning_with_ghosts_and_sort_payload(emd, gu, ctx):
    import torch
    md = liquid()
    md.stage("exchange")
    n = md.geti("N_local")
    before = {k: md.arr(k)[:n].copy() for k in ("x", "v", "f", "type", "id", "q")}
    md.stage("bin_sort")
    g = gu.geom_from(md.geom())
    d = {k: gu.dev(v) for k, v in before.items()}
    bc, bo, pv = gu.binning_build(ctx, d["x"], n, g)
    out = {k: torch.empty_like(v) for k, v in d.items()}
    order = ("x", "v", "f", "type", "id", "q")
    emd.check(emd.lib().emd_binning_permute(ctx.handle, gu.ptr(pv), n, *[gu.ptr(d[k]) for k in order], *[gu.ptr(out[k]) for k in order]))
    for k in order:
        np.testing.assert_array_equal(out[k].cpu().numpy(), md.arr(k)[:n], err_msg=k)
    # second pass: locals + ghosts, no sort (examinimd.cpp:217)
    md.stage("halo", "bin_all")
    xa = md.arr("x")
    g2 = gu.geom_from(md.geom())
    bc2, bo2, pv2 = gu.binning_build(ctx, gu.dev(xa), xa.shape[0], g2)
    np.testing.assert_array_equal(bc2.cpu().numpy(), md.arr("bincount"))
    np.testing.assert_array_equal(pv2.cpu().numpy(), md.arr("permute"))
    md.close()


def test_binning_empty_and_ragged(emd, gu, ctx):
    rng = np.random.default_rng(5)
    box = np.array([17.0, 9.0, 23.0])
    # ragged: a dense blob, a sparse gas and many empty bins
    x = np.concatenate([rng.uniform(0, 1, (300, 3)) * [2, 2, 2] + [3, 3, 3], rng.uniform(0, 1, (200, 3)) * box])
    md = OracleMD.from_arrays(x, box)
    md.stage("exchange", "bin_nosort")
    g = gu.geom_from(md.geom())
    bc, bo, pv = gu.binning_build(ctx, gu.dev(md.arr("x")[:500]), 500, g)
    np.testing.assert_array_equal(bc.cpu().numpy(), md.arr("bincount"))
    np.testing.assert_array_equal(pv.cpu().numpy(), md.arr("permute"))
    assert (md.arr("bincount") == 0).sum() > 10 and md.arr("bincount").max() > 40
    # empty range
    bc0, bo0, pv0 = gu.binning_build(ctx, gu.dev(np.zeros((1, 3))), 0, g)
    assert int(bc0.sum()) == 0 and int(bo0.max()) == 0
    md.close()


def test_binning_lost_atom_is_an_error(emd, gu, ctx):
    md = OracleMD.from_deck(DECK, "CSR", "NEIGH_FULL", region=(6, 6, 6), setup=False)
    md.stage("exchange", "bin_nosort")
    g = gu.geom_from(md.geom())
    x = md.arr("x")[: md.geti("N_local")].copy()
    x[7, 1] = 1e6
    with pytest.raises(emd.EmdError, match="outside the bin grid"):
        gu.binning_build(ctx, gu.dev(x), x.shape[0], g)
    md.close()


# ---------------------------------------------------------------------------- neighbor (a2)
def _gpu_rows_csr(gu, ctx, md, half, newton):
    x = md.arr("x")
    n = md.geti("N_local")
    g = gu.geom_from(md.geom())
    bc, bo, pv = gu.dev(md.arr("bincount")), gu.dev(md.arr("binoffsets")), gu.dev(md.arr("permute"))
    rm, ent, total = gu.neigh_csr(ctx, gu.dev(x), n, g, bc, bo, pv, md.getd("neigh_cutoff"), half, newton)
    return rm.cpu().numpy(), ent.cpu().numpy(), total


@pytest.mark.parametrize("iteration", ["NEIGH_FULL", "NEIGH_HALF"])
@pytest.mark.parametrize("state", ["lattice", "liquid"])
def test_neighbor_csr_rows_identical(emd, gu, ctx, iteration, state):
    half = iteration == "NEIGH_HALF"
    md = OracleMD.from_deck(DECK, "CSR", iteration, region=(10, 10, 10)) if state == "lattice" else rebuilt(liquid(iteration=iteration))
    rm, ent, total = _gpu_rows_csr(gu, ctx, md, half, 0)
    assert total == md.geti("total_neighs")
    np.testing.assert_array_equal(rm, md.arr("row_map"))
    np.testing.assert_array_equal(ent, md.arr("entries"))  # same row ORDER as the 1-thread reference, not just the same set
    if state == "lattice" and not half:
        assert np.all(np.diff(rm) == 78)
    md.close()


def test_neighbor_half_newton_on(emd, gu, ctx):
    x = OracleMD.from_deck(DECK, "CSR", "NEIGH_HALF", region=(8, 8, 8), setup=False)
    box = [x.getd("domain_x")] * 3
    md = OracleMD.from_arrays(x.arr("x")[: x.geti("N")], box, newton=1, iteration="NEIGH_HALF")
    rebuilt(md)
    rm, ent, total = _gpu_rows_csr(gu, ctx, md, True, 1)
    np.testing.assert_array_equal(rm, md.arr("row_map"))
    np.testing.assert_array_equal(ent, md.arr("entries"))
    assert np.all(np.diff(rm) == 39)  # newton on: every pair exactly once, incl. ghost pairs
    x.close(); md.close()


@pytest.mark.parametrize("iteration", ["NEIGH_FULL", "NEIGH_HALF"])
def test_neighbor_2d_rows_identical(emd, gu, ctx, iteration):
    half = iteration == "NEIGH_HALF"
    md = rebuilt(liquid(neigh="2D", iteration=iteration))
    n = md.geti("N_local")
    g = gu.geom_from(md.geom())
    bc, bo, pv = gu.dev(md.arr("bincount")), gu.dev(md.arr("binoffsets")), gu.dev(md.arr("permute"))
    nn, tab, stride, passes = gu.neigh_2d(ctx, gu.dev(md.arr("x")), n, g, bc, bo, pv, md.getd("neigh_cutoff"), half, 0)
    o_nn = md.arr("num_neighs")
    np.testing.assert_array_equal(nn.cpu().numpy(), o_nn)
    o_tab = md.arr("neighs2d").reshape(n, md.geti("maxneighs"))
    t = tab.cpu().numpy()
    for i in range(0, n, 7):
        np.testing.assert_array_equal(t[i, : o_nn[i]], o_tab[i, : o_nn[i]])
    assert passes == 2 and stride >= o_nn.max()  # 16 -> overflow -> one resize (neighbor_2d.h:304-330)
    md.close()


def test_neighbor_ragged_dense_bins(emd, gu, ctx):
    """a stencil with more candidates than one staging tile (1024) and empty bins around it."""
    rng = np.random.default_rng(11)
    box = np.array([12.0, 12.0, 12.0])
    x = np.concatenate([rng.uniform(4.0, 7.5, (1500, 3)), rng.uniform(0, 12, (60, 3))])
    for it in ("NEIGH_FULL", "NEIGH_HALF"):
        md = OracleMD.from_arrays(x, box, force_cutoff=2.5, skin=0.3, iteration=it)
        rebuilt(md)
        rm, ent, total = _gpu_rows_csr(gu, ctx, md, it == "NEIGH_HALF", 0)
        np.testing.assert_array_equal(rm, md.arr("row_map"))
        np.testing.assert_array_equal(ent, md.arr("entries"))
        assert np.diff(rm).max() > 200
        md.close()


# ------------------------------------------------------- tile lists: the B200 fast path (a2 + a3)
def _tiles_for(gu, ctx, md):
    x = gu.dev(md.arr("x"))
    n, na = md.geti("N_local"), md.geti("N_local") + md.geti("N_ghost")
    g = gu.geom_from(md.geom())
    bc, bo, pv = gu.dev(md.arr("bincount")), gu.dev(md.arr("binoffsets")), gu.dev(md.arr("permute"))
    return gu.Tiles(ctx, x, n, na, g, bc, bo, pv, md.getd("neigh_cutoff")), x


@pytest.mark.parametrize("iteration", ["NEIGH_FULL", "NEIGH_HALF"])
@pytest.mark.parametrize("state", ["lattice", "liquid"])
def test_tiles_emit_reference_rows(emd, gu, ctx, iteration, state):
    """the CSR and 2D lists emitted from the tile lists are the reference's rows, entry by entry"""
    half = iteration == "NEIGH_HALF"
    md = OracleMD.from_deck(DECK, "CSR", iteration, region=(10, 10, 10)) if state == "lattice" else rebuilt(liquid(iteration=iteration))
    t, _ = _tiles_for(gu, ctx, md)
    assert t.ok, t.info()
    rm, ent, total = t.csr(half, 0)
    assert total == md.geti("total_neighs")
    np.testing.assert_array_equal(rm.cpu().numpy(), md.arr("row_map"))
    np.testing.assert_array_equal(ent.cpu().numpy(), md.arr("entries"))
    nn, tab, stride, passes = t.table(half, 0)
    o_rm = md.arr("row_map")
    np.testing.assert_array_equal(nn.cpu().numpy(), np.diff(o_rm))
    tt, oe = tab.cpu().numpy(), md.arr("entries")
    for i in range(0, md.geti("N_local"), 11):
        np.testing.assert_array_equal(tt[i, : o_rm[i + 1] - o_rm[i]], oe[o_rm[i]: o_rm[i + 1]])
    assert passes == 2
    t.close(); md.close()


def test_tiles_half_newton_on_rows(emd, gu, ctx):
    x = OracleMD.from_deck(DECK, "CSR", "NEIGH_HALF", region=(8, 8, 8), setup=False)
    box = [x.getd("domain_x")] * 3
    md = OracleMD.from_arrays(x.arr("x")[: x.geti("N")], box, newton=1, iteration="NEIGH_HALF")
    rebuilt(md)
    t, _ = _tiles_for(gu, ctx, md)
    assert t.ok
    rm, ent, total = t.csr(True, 1)
    np.testing.assert_array_equal(rm.cpu().numpy(), md.arr("row_map"))
    np.testing.assert_array_equal(ent.cpu().numpy(), md.arr("entries"))
    t.close(); x.close(); md.close()


@pytest.mark.parametrize("iteration", ["NEIGH_FULL", "NEIGH_HALF"])
def test_tiles_lj_force_and_energy(emd, gu, ctx, iteration):
    """owned-atom forces and the shifted PE from the tile lists equal the reference's full- and half-list results"""
    import torch
    md = rebuilt(liquid(iteration=iteration))
    md.stage("zero_f", "force")
    n = md.geti("N_local")
    t, x = _tiles_for(gu, ctx, md)
    assert t.ok
    _set_lj(emd, ctx, md)
    typ = gu.dev(md.arr("type"))
    f = torch.full((x.shape[0], 3), 7.0, dtype=torch.float64, device="cuda")
    t.force(x, typ, f)
    fo = md.arr("f")
    assert np.abs(f.cpu().numpy()[:n] - fo[:n]).max() / np.sqrt((fo[:n] ** 2).mean()) < TOL
    assert (f.cpu().numpy()[n:] == 7.0).all()  # ghost rows are not touched
    pe = t.force(x, typ, f, energy=True)
    _, PE, _ = md.thermo()
    assert abs(pe / md.geti("N") - PE) < 1e-12 * abs(PE) * 100
    # positions move between rebuilds: same lists, new x (skin not exceeded)
    rng = np.random.default_rng(2)
    xm = md.arr("x").copy()
    xm[:n] += rng.normal(0, 0.02, (n, 3))
    md.set("x", xm)
    md.stage("update_halo", "zero_f", "force")
    x2 = gu.dev(md.arr("x"))
    t.force(x2, typ, f)
    fo = md.arr("f")
    assert np.abs(f.cpu().numpy()[:n] - fo[:n]).max() / np.sqrt((fo[:n] ** 2).mean()) < TOL
    t.close(); md.close()


@pytest.mark.parametrize("state", ["lattice", "liquid"])
def test_tiles_force_rows_are_a_conflict_light_permutation(emd, gu, ctx, state):
    """the force kernel's rows (tiles_lists_kernel) hold, row by row, every entry of the exact full list once, plus only
    pairs within the FP32 search margin of the list radius; an entry is the byte offset of the neighbor's staged
    coordinates; padding points at the dummy slots; within a column the lanes of a half-warp mostly read different
    shared-memory banks; and the padding costs only a few per cent of columns"""
    md = OracleMD.from_deck(DECK, "CSR", "NEIGH_FULL", region=(10, 10, 10)) if state == "lattice" else rebuilt(liquid(iteration="NEIGH_FULL"))
    t, _ = _tiles_for(gu, ctx, md)
    assert t.ok, t.info()
    rm, ent, total = t.csr(False, 0)  # exact full rows of this build
    assert total == md.geti("total_neighs")
    L = t.lists()
    nt, st, cap = L["ntiles"], L["stride"], L["cap"]
    x, cut = md.arr("x"), md.getd("neigh_cutoff")
    rows = L["csr16"].transpose(0, 2, 1, 3).reshape(nt, st, -1)        # [tile][thread][q]
    rows_s = L["ell_s"].transpose(0, 2, 1, 3).reshape(nt, st, -1)
    cols_total, longest_total, conflicts, pad_conflicts, pairs, extras = 0, 0, 0, 0, 0, 0
    for tile in range(nt):
        for w0 in range(0, st, 32):
            n = L["ncsr"][tile, w0:w0 + 32]
            c8 = L["nell_s"][tile, w0:w0 + 32]
            assert (c8 == c8[0]).all() and c8[0] % 8 == 0 and c8[0] <= L["maxrow"]
            assert c8[0] >= n.max()
            blk = rows_s[tile, w0:w0 + 32, : c8[0]].astype(np.int64)
            assert (blk % 24 == 0).all()
            blk = blk // 24 - 16               # slots; the 16 dummy atoms in front of the buffer are -16..-1
            assert (blk < cap).all()
            nreal = 0
            for l in range(32):
                if n[l] == 0 and not (blk[l] >= 0).any():
                    continue
                real = np.sort(blk[l][blk[l] >= 0])
                exact = np.sort(rows[tile, w0 + l, : n[l]].astype(np.int64))
                assert np.unique(real).size == real.size
                extra = np.setdiff1d(real, exact)
                assert np.setdiff1d(exact, real).size == 0
                if extra.size:  # only pairs in the rounding margin of the search
                    own = L["stg_j"][tile, L["int_slot"][tile, w0 + l]]
                    d = np.linalg.norm(x[L["stg_j"][tile, extra]] - x[own], axis=1)
                    assert (d > cut).all() and (d < cut * (1 + 1e-3)).all(), d
                    extras += extra.size
                nreal = max(nreal, real.size)
            assert c8[0] == (nreal + 7) // 8 * 8
            for h in (slice(0, 16), slice(16, 32)):
                for q in range(c8[0]):
                    col = blk[h, q]
                    u = np.unique(col[col >= 0])             # slots whose pairs are evaluated
                    conflicts += u.size - np.unique(u % 16).size
                    u = np.unique(col)                       # with the padding lanes' reads
                    pad_conflicts += u.size - np.unique(u % 16).size
            cols_total += int(c8[0]); longest_total += nreal; pairs += int(n.sum())
    assert pairs == total
    print(f"force rows/{state}: {conflicts / (2 * cols_total):.3f} extra wavefronts per half-warp column, "
          f"{pad_conflicts / (2 * cols_total):.3f} with padding reads, {cols_total / max(longest_total, 1):.3f} x longest-row columns, "
          f"{extras} entries beyond the list radius")
    assert cols_total <= longest_total + 8 * nt * (st // 32)
    assert conflicts / (2 * cols_total) < 2.0
    t.close(); md.close()


def test_tiles_force_in_two_parts(emd, gu, ctx):
    """the halo-independent tiles and the tiles that read ghosts partition the force: part 1 + part 2 write exactly the rows
    of the single launch, bit for bit, and part 1 really does not depend on the ghost coordinates"""
    import torch
    md = OracleMD.from_deck(DECK, "CSR", "NEIGH_FULL", region=(16, 16, 16))
    md.step(3)
    rebuilt(md)
    n = md.geti("N_local")
    t, x = _tiles_for(gu, ctx, md)
    assert t.ok
    n_free, n_halo = t.halo_split()
    assert n_free > 0 and n_halo > 0 and n_free + n_halo == t.info()["ntiles"]
    _set_lj(emd, ctx, md)
    typ = gu.dev(md.arr("type"))
    f0 = torch.full((x.shape[0], 3), 7.0, dtype=torch.float64, device="cuda")
    t.force(x, typ, f0)
    f1 = torch.full_like(f0, 7.0)
    xg = x.clone()
    xg[n:] = float("nan")                      # part 1 must not read a single ghost coordinate
    t.force_part(xg, typ, f1, 1, reserve=16)
    ctx.sync()
    touched1 = (f1[:n] != 7.0).any(dim=1)
    assert bool(touched1.any()) and not bool(touched1.all())
    assert not bool(torch.isnan(f1).any())
    t.force_part(x, typ, f1, 2)
    ctx.sync()
    assert torch.equal(f0, f1)
    t.close(); md.close()


def test_tiles_two_types_and_ragged(emd, gu, ctx):
    import torch
    rng = np.random.default_rng(3)
    base = OracleMD.from_deck(DECK, "CSR", "NEIGH_FULL", region=(7, 7, 7), setup=False)
    x = base.arr("x")[: base.geti("N")] + rng.normal(0, 0.05, (base.geti("N"), 3))
    box = [base.getd("domain_x")] * 3
    types = rng.integers(0, 2, x.shape[0]).astype(np.int32)
    md = OracleMD.from_arrays(x, box, types=types, ntypes=2, mass=[1.0, 3.0], iteration="NEIGH_FULL")
    rebuilt(md)
    md.stage("zero_f", "force")
    n = md.geti("N_local")
    t, xd = _tiles_for(gu, ctx, md)
    assert t.ok
    _set_lj(emd, ctx, md, ntypes=2)
    f = torch.zeros((xd.shape[0], 3), dtype=torch.float64, device="cuda")
    t.force(xd, gu.dev(md.arr("type")), f)
    fo = md.arr("f")
    assert np.abs(f.cpu().numpy()[:n] - fo[:n]).max() / np.sqrt((fo[:n] ** 2).mean()) < TOL
    t.close(); base.close(); md.close()
    # a dense blob does not fit a tile: the build must say "not applicable" (3), never produce a wrong list
    box = np.array([12.0, 12.0, 12.0])
    xb = np.concatenate([rng.uniform(4.0, 7.5, (1500, 3)), rng.uniform(0, 12, (60, 3))])
    md = OracleMD.from_arrays(xb, box, force_cutoff=2.5, skin=0.3, iteration="NEIGH_FULL")
    rebuilt(md)
    t, _ = _tiles_for(gu, ctx, md)
    rm, ent, total = t.csr(False, 0) if t.ok else (None, None, -1)
    if total >= 0:
        np.testing.assert_array_equal(rm.cpu().numpy(), md.arr("row_map"))
        np.testing.assert_array_equal(ent.cpu().numpy(), md.arr("entries"))
    else:
        assert t.rc == 3
    t.close(); md.close()


# ---------------------------------------------------------------------------- LJ force (a3)
def _set_lj(emd, ctx, md, ntypes=1):
    L = emd.lib()
    arr = lambda v: (C.c_double * (ntypes * ntypes))(*([v] * (ntypes * ntypes)))
    emd.check(L.emd_force_lj_set_params(ctx.handle, ntypes, arr(md.getd("lj1")), arr(md.getd("lj2")), arr(md.getd("cutsq"))))


@pytest.mark.parametrize("iteration", ["NEIGH_FULL", "NEIGH_HALF"])
@pytest.mark.parametrize("kind", ["CSR", "2D"])
def test_lj_force_and_energy(emd, gu, ctx, iteration, kind):
    import torch
    half = iteration == "NEIGH_HALF"
    md = rebuilt(liquid(neigh=kind, iteration=iteration))
    md.stage("zero_f", "force")
    n, na = md.geti("N_local"), md.geti("N_local") + md.geti("N_ghost")
    x, typ = gu.dev(md.arr("x")), gu.dev(md.arr("type"))
    if kind == "CSR":
        rm, ent = gu.dev(md.arr("row_map")), gu.dev(md.arr("entries"))
        lst = gu.csr_list(rm, ent)
    else:
        nn, tab = gu.dev(md.arr("num_neighs")), gu.dev(md.arr("neighs2d"))
        lst = gu.table_list(nn, tab, md.geti("maxneighs"))
    _set_lj(emd, ctx, md)
    f = torch.full((na, 3), 7.0, dtype=torch.float64, device="cuda")  # garbage: zero_f=1 must clear it
    emd.check(emd.lib().emd_force_lj_compute(ctx.handle, gu.ptr(x), gu.ptr(typ), gu.ptr(f), n, na, C.byref(lst), half, 1))
    fo = md.arr("f")
    scale = np.sqrt((fo[:n] ** 2).mean())
    rows = na if half else n  # half mode also scatters onto ghost rows (force_lj_neigh_impl.h:245-247)
    assert np.abs(f.cpu().numpy()[:rows] - fo[:rows]).max() / scale < TOL
    # accumulate semantics (+=) with zero_f=0
    f2 = torch.zeros((na, 3), dtype=torch.float64, device="cuda")
    f2[:n] = 1.0
    emd.check(emd.lib().emd_force_lj_compute(ctx.handle, gu.ptr(x), gu.ptr(typ), gu.ptr(f2), n, na, C.byref(lst), half, 0))
    assert np.abs(f2.cpu().numpy()[:n] - 1.0 - fo[:n]).max() / scale < TOL
    pe = C.c_double()
    emd.check(emd.lib().emd_force_lj_energy(ctx.handle, gu.ptr(x), gu.ptr(typ), n, C.byref(lst), half, C.byref(pe)))
    _, PE, _ = md.thermo()
    assert abs(pe.value / md.geti("N") - PE) < 1e-12 * abs(PE) * 100
    md.close()


def test_lj_force_two_types(emd, gu, ctx):
    import torch
    rng = np.random.default_rng(3)
    base = OracleMD.from_deck(DECK, "CSR", "NEIGH_FULL", region=(7, 7, 7), setup=False)
    x = base.arr("x")[: base.geti("N")] + rng.normal(0, 0.05, (base.geti("N"), 3))
    box = [base.getd("domain_x")] * 3
    types = rng.integers(0, 2, x.shape[0]).astype(np.int32)
    md = OracleMD.from_arrays(x, box, types=types, ntypes=2, mass=[1.0, 3.0], iteration="NEIGH_HALF")
    rebuilt(md)
    md.stage("zero_f", "force")
    n, na = md.geti("N_local"), md.geti("N_local") + md.geti("N_ghost")
    _set_lj(emd, ctx, md, ntypes=2)
    lst = gu.csr_list(*(keep := (gu.dev(md.arr("row_map")), gu.dev(md.arr("entries")))))
    f = torch.zeros((na, 3), dtype=torch.float64, device="cuda")
    xd, td = gu.dev(md.arr("x")), gu.dev(md.arr("type"))  # keep the device buffers alive across the call
    emd.check(emd.lib().emd_force_lj_compute(ctx.handle, gu.ptr(xd), gu.ptr(td), gu.ptr(f), n, na, C.byref(lst), 1, 1))
    fo = md.arr("f")
    assert np.abs(f.cpu().numpy()[:n] - fo[:n]).max() / np.sqrt((fo[:n] ** 2).mean()) < TOL
    base.close(); md.close()


# -------------------------------------------------------------------------- integrator (a5)
def test_integrator_bit_exact(emd, gu, ctx):
    md = liquid(nsteps=7)
    n = md.geti("N_local")
    x, v, f = (gu.dev(md.arr(k)[:n]) for k in ("x", "v", "f"))
    typ, mass = gu.dev(md.arr("type")[:n]), gu.dev(md.arr("mass"))
    dt, mvv2e = md.getd("dt"), md.getd("mvv2e")
    dtf = 0.5 * dt / mvv2e
    L = emd.lib()
    emd.check(L.emd_nve_initial_integrate(ctx.handle, gu.ptr(x), gu.ptr(v), gu.ptr(f), gu.ptr(typ), gu.ptr(mass), n, dtf, dt))
    md.stage("initial_integrate")
    np.testing.assert_array_equal(x.cpu().numpy(), md.arr("x")[:n])
    np.testing.assert_array_equal(v.cpu().numpy(), md.arr("v")[:n])
    emd.check(L.emd_nve_final_integrate(ctx.handle, gu.ptr(v), gu.ptr(f), gu.ptr(typ), gu.ptr(mass), n, dtf))
    md.stage("final_integrate")
    np.testing.assert_array_equal(v.cpu().numpy(), md.arr("v")[:n])
    md.close()


# ---------------------------------------------------------------------------- CommSerial (a6)
def test_comm_serial_kernels(emd, gu, ctx):
    import torch
    md = liquid(nsteps=19)
    md.stage("initial_integrate")  # some atoms now sit outside [0,L)
    n = md.geti("N_local")
    L3 = [md.getd("domain_x"), md.getd("domain_y"), md.getd("domain_z")]
    lib = emd.lib()
    cap = 3 * n
    x = torch.zeros((cap, 3), dtype=torch.float64, device="cuda"); x[:n] = gu.dev(md.arr("x")[:n])
    v = torch.zeros((cap, 3), dtype=torch.float64, device="cuda"); v[:n] = gu.dev(md.arr("v")[:n])
    q = torch.zeros(cap, dtype=torch.float64, device="cuda")
    idt = torch.zeros(cap, dtype=torch.int32, device="cuda"); idt[:n] = gu.dev(md.arr("id")[:n])
    typ = torch.zeros(cap, dtype=torch.int32, device="cuda")
    emd.check(lib.emd_comm_wrap(ctx.handle, gu.ptr(x), n, emd.vec3(L3)))
    md.stage("exchange")
    np.testing.assert_array_equal(x[:n].cpu().numpy(), md.arr("x")[:n])
    assert (md.arr("x")[:n] >= 0).all()
    # six halo phases, bit-exact ghosts in the same order
    md.stage("halo")
    packs, counts, nghost = [], [], 0
    depth = md.getd("neigh_cutoff")
    for phase in range(6):
        nscan = n + nghost - (counts[phase - 1] if phase % 2 == 1 else 0)
        pk = torch.zeros(cap, dtype=torch.int32, device="cuda")
        cnt = C.c_int()
        emd.check(lib.emd_comm_halo_phase(ctx.handle, phase, gu.ptr(x), gu.ptr(v), gu.ptr(q), gu.ptr(idt), gu.ptr(typ), nscan,
                                          n + nghost, cap, gu.ptr(pk), cap, emd.vec3(L3), emd.vec3([0, 0, 0]), emd.vec3(L3), depth,
                                          C.byref(cnt)))
        assert cnt.value == md.geti(f"num_ghost{phase}")
        np.testing.assert_array_equal(pk[: cnt.value].cpu().numpy(), md.arr(f"pack{phase}"))
        packs.append(pk); counts.append(cnt.value); nghost += cnt.value
    na = n + nghost
    assert nghost == md.geti("N_ghost")
    np.testing.assert_array_equal(x[:na].cpu().numpy(), md.arr("x"))
    np.testing.assert_array_equal(idt[:na].cpu().numpy(), md.arr("id"))
    # capacity overflow writes nothing and reports the count (caller grows + redoes)
    cnt = C.c_int()
    x_before = x.clone()
    emd.check(lib.emd_comm_halo_phase(ctx.handle, 0, gu.ptr(x), gu.ptr(v), gu.ptr(q), gu.ptr(idt), gu.ptr(typ), n, n, n + 3,
                                      gu.ptr(packs[0]), cap, emd.vec3(L3), emd.vec3([0, 0, 0]), emd.vec3(L3), depth, C.byref(cnt)))
    assert cnt.value == counts[0] and torch.equal(x, x_before)
    # halo update after moving the owned atoms
    rng = np.random.default_rng(0)
    xm = md.arr("x")[:n] + rng.normal(0, 1e-3, (n, 3))
    md.set("x", xm)
    x[:n] = gu.dev(xm)
    md.stage("update_halo")
    g0 = n
    for phase in range(6):
        emd.check(lib.emd_comm_halo_update_phase(ctx.handle, phase, gu.ptr(x), gu.ptr(v), gu.ptr(q), gu.ptr(idt), gu.ptr(typ),
                                                 gu.ptr(packs[phase]), counts[phase], g0, emd.vec3(L3)))
        g0 += counts[phase]
    np.testing.assert_array_equal(x[:na].cpu().numpy(), md.arr("x"))
    # reverse force fold, phases 5..0
    fr = rng.normal(0, 1, (na, 3))
    md.set("f", fr)
    f = gu.dev(fr)
    md.stage("update_force")
    offs = np.concatenate([[n], n + np.cumsum(counts)[:-1]])
    for phase in range(5, -1, -1):
        emd.check(lib.emd_comm_force_fold_phase(ctx.handle, gu.ptr(f), gu.ptr(packs[phase]), counts[phase], int(offs[phase])))
    np.testing.assert_array_equal(f.cpu().numpy(), md.arr("f"))
    md.close()


# -------------------------------------------------------------------------------- thermo (a8)
def test_reduce_mv2(emd, gu, ctx):
    md = liquid(nsteps=3)
    n = md.geti("N_local")
    s = C.c_double()
    vd, td, md_ = gu.dev(md.arr("v")[:n]), gu.dev(md.arr("type")[:n]), gu.dev(md.arr("mass"))
    emd.check(emd.lib().emd_reduce_mv2(ctx.handle, gu.ptr(vd), gu.ptr(td), gu.ptr(md_), n, C.byref(s)))
    T, _, KE = md.thermo()
    assert abs(s.value * 0.5 * md.getd("mvv2e") / md.geti("N") - KE) < 1e-13 * KE * 10
    md.close()
