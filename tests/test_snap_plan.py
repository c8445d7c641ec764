"""Host logic of the SNAP Y kernel (no GPU): the work plan that emd_snap_create builds for snap_yi -- items of one or two output
rows, their segments per block (j1,j2,j) and the Clebsch-Gordan step table -- against a brute-force enumeration of
SNA::compute_zi's loop nest (src/force_types/sna_impl.hpp:196-283): every (block, output row, ma1) exactly once (checked inside
emd_snap_yi_plan_stats), and the non-zero terms the plan executes = the terms of the loop nest on the half range."""
import ctypes as C

import pytest


def brute_force_terms(twojmax):
    """inner (ma1, mb1) iterations of compute_zi summed over the outputs the force needs: mb <= j/2, without the rows below the
    diagonal of an even level's middle column (their weight in compute_dbidrj is 0, :393-424)"""
    n = 0
    for j1 in range(twojmax + 1):
        for j2 in range(j1 + 1):
            for j in range(j1 - j2, min(twojmax, j1 + j2) + 1, 2):
                for ma in range(j + 1):
                    for mb in range(j // 2 + 1):
                        if j % 2 == 0 and mb == j // 2 and ma > j // 2:
                            continue
                        na = min(j1, (2 * ma - j + j2 + j1) // 2) - max(0, (2 * ma - j - j2 + j1) // 2) + 1
                        nb = min(j1, (2 * mb - j + j2 + j1) // 2) - max(0, (2 * mb - j - j2 + j1) // 2) + 1
                        n += max(na, 0) * max(nb, 0)
    return n


@pytest.mark.parametrize("twojmax", [0, 1, 2, 3, 4, 6, 8])
def test_yi_plan_covers_compute_zi(twojmax):
    import examinimd_b200 as emd
    L = emd.lib()
    nitems, nsegs, ntab = C.c_int(), C.c_int(), C.c_int()
    executed, nonzero = C.c_longlong(), C.c_longlong()
    rc = L.emd_snap_yi_plan_stats(twojmax, C.byref(nitems), C.byref(nsegs), C.byref(ntab), C.byref(executed), C.byref(nonzero))
    assert rc == 0, L.emd_last_error().decode()
    want = brute_force_terms(twojmax)
    # the plan executes every term of the loop nest (plus zero padding of ragged steps: at most a quarter more); a Clebsch-Gordan
    # factor can vanish inside its block (e.g. cg(2,2,2; m1 = m2 = 1)), so the non-zero ones are at most the loop nest's
    assert nonzero.value <= want <= executed.value <= 1.25 * want + 64, (nonzero.value, want, executed.value)
    assert nonzero.value >= 0.8 * want
    assert 0 < nitems.value <= 64 and nsegs.value > 0 and ntab.value % 2 == 0
    if twojmax == 8:
        assert want == 38358
        assert executed.value == 45656
