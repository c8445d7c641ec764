#include "binning_kksort.h"
#include <cstdio>
#include <cstdlib>

BinningKKSort::BinningKKSort(System *s) : Binning(s) {}

static void fail(const char *what) {
  fprintf(stderr, "BinningKKSort: %s: %s\n", what, emd_last_error());
  emd_host_exit(1);
}

// src/binning_types/binning_kksort.cpp:71-140, expressed as three C-ABI calls.
void BinningKKSort::create_binning(T_X_FLOAT dx_in, T_X_FLOAT dy_in, T_X_FLOAT dz_in, int halo_depth, bool do_local,
                                   bool do_ghost, bool sort) {
  if (!(do_local || do_ghost)) return;
  const T_INT begin = do_local ? 0 : system->N_local;
  const T_INT end = do_ghost ? system->N_local + system->N_ghost : system->N_local;

  const double sub[3] = {system->sub_domain_x, system->sub_domain_y, system->sub_domain_z};
  const double lo[3] = {system->sub_domain_lo_x, system->sub_domain_lo_y, system->sub_domain_lo_z};
  const double hi[3] = {system->sub_domain_hi_x, system->sub_domain_hi_y, system->sub_domain_hi_z};
  emd_bin_geom g;
  emd_binning_geometry(sub, lo, hi, dx_in, dy_in, dz_in, halo_depth, &g);
  nbinx = g.nbinx; nbiny = g.nbiny; nbinz = g.nbinz; nhalo = g.nhalo;
  minx = g.minx; maxx = g.maxx; miny = g.miny; maxy = g.maxy; minz = g.minz; maxz = g.maxz;

  const size_t nbins = (size_t)nbinx * nbiny * nbinz;
  if (bincount_buf.extent() < nbins) {
    if (!bincount_buf.alloc(nbins) || !binoffsets_buf.alloc(nbins)) fail("alloc bins");
  }
  const size_t n = (size_t)(end - begin);
  if (permute_buf.extent() < n) { if (!permute_buf.alloc(n + n / 8)) fail("alloc permute"); }
  bincount = bincount_buf.ptr; binoffsets = binoffsets_buf.ptr; permute_vector = permute_buf.ptr;

  if (emd_binning_build(system->ctx, system->x + 3 * (size_t)begin, (int)n, &g, bincount, binoffsets, permute_vector))
    fail("emd_binning_build");

  if (sort) {
    // the reference gathers into a scratch copy and copies back, six times; here one gather
    // writes the alternate buffers and the System swaps pointers (modules re-fetch handles
    // on every call, as they must already because grow() may reallocate)
    if (begin != 0) fail("sort requires do_local");
    if (emd_binning_permute(system->ctx, permute_vector, (int)n, system->x, system->v, system->f, system->type, system->id,
                            system->q, system->x_alt, system->v_alt, system->f_alt, system->type_alt, system->id_alt,
                            system->q_alt))
      fail("emd_binning_permute");
    // atoms behind the sorted range (stale ghosts of the previous halo) keep their slots, as in
    // the reference's in-place sort: carry them over to the new primary buffers
    const size_t tail = (size_t)(system->N_local + system->N_ghost - end), e = (size_t)end;
    if (tail > 0) {
      emd_ctx *c = system->ctx;
      int rc = 0;
      rc |= emd_memcpy_d2d(c, system->x_alt + 3 * e, system->x + 3 * e, sizeof(double) * 3 * tail);
      rc |= emd_memcpy_d2d(c, system->v_alt + 3 * e, system->v + 3 * e, sizeof(double) * 3 * tail);
      rc |= emd_memcpy_d2d(c, system->f_alt + 3 * e, system->f + 3 * e, sizeof(double) * 3 * tail);
      rc |= emd_memcpy_d2d(c, system->type_alt + e, system->type + e, sizeof(int) * tail);
      rc |= emd_memcpy_d2d(c, system->id_alt + e, system->id + e, sizeof(int) * tail);
      rc |= emd_memcpy_d2d(c, system->q_alt + e, system->q + e, sizeof(double) * tail);
      if (rc) fail("carry ghosts");
    }
    system->swap_sorted();
    is_sorted = true;
  }
}

const char *BinningKKSort::name() { return "BinningKKSort"; }
