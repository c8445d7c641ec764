// Host side of the neighbor modules: buffer management around the emd_neigh_* C-ABI calls.
#include "neighbor_2d.h"
#include "neighbor_csr.h"
#include <cstdio>
#include <cstdlib>

static void fail(const char *who, const char *what) {
  fprintf(stderr, "%s: %s: %s\n", who, what, emd_last_error());
  emd_host_exit(1);
}

bool Neighbor::build_tiles(System *system, Binning *binning, T_X_FLOAT neigh_cut) {
  static const bool disabled = getenv("EMD_NO_TILES") && atoi(getenv("EMD_NO_TILES")) != 0;
  if (disabled) { if (tile_lists) emd_tiles_invalidate(tile_lists); return false; }
  if (!tile_lists && emd_tiles_create(&tile_lists)) fail("Neighbor", "tiles create");
  const emd_bin_geom g = binning->geom();
  const int rc = emd_neigh_tiles_build(system->ctx, tile_lists, system->x, system->N_local, system->N_local + system->N_ghost, &g,
                                       binning->bincount, binning->binoffsets, binning->permute_vector, neigh_cut);
  if (rc != 0 && rc != 3) fail("Neighbor", "tiles build");
  return rc == 0;
}

// src/neighbor_types/neighbor_csr.h:370-435
void NeighborCSR::create_neigh_list(System *system, Binning *binning, bool half_neigh_, bool) {
  const T_INT N_local = system->N_local;
  if (neigh_offsets.extent() < (size_t)N_local + 1) {
    if (!neigh_offsets.alloc((size_t)N_local + 1 + N_local / 16)) fail("NeighborCSR", "alloc row_map");
  }
  const emd_bin_geom g = binning->geom();
  int total = 0;
  if (build_tiles(system, binning, neigh_cut)) {
    // fast path: exact CSR rows made from the tile search (same rows, same order); 3 = a tile outgrew the fast path
    const int rc = emd_neigh_tiles_count(system->ctx, tile_lists, half_neigh_, comm_newton, neigh_offsets.ptr, &total);
    if (rc != 0 && rc != 3) fail("NeighborCSR", "tiles count");
    if (rc == 0) {
      if (neighs.extent() < (size_t)total) {
        if (!neighs.alloc((size_t)total + total / 16)) fail("NeighborCSR", "alloc entries");
      }
      if (emd_neigh_tiles_fill_csr(system->ctx, tile_lists, half_neigh_, comm_newton, neigh_offsets.ptr, neighs.ptr))
        fail("NeighborCSR", "tiles fill");
      neigh_list.row_map = neigh_offsets.ptr;
      neigh_list.entries = neighs.ptr;
      neigh_list.N_local = N_local;
      neigh_list.total = total;
      return;
    }
    emd_tiles_invalidate(tile_lists);
  }
  if (emd_neigh_csr_count(system->ctx, system->x, N_local, &g, binning->bincount, binning->binoffsets,
                          binning->permute_vector, neigh_cut, half_neigh_, comm_newton, neigh_offsets.ptr, &total))
    fail("NeighborCSR", "count");
  if (neighs.extent() < (size_t)total) {
    if (!neighs.alloc((size_t)total + total / 16)) fail("NeighborCSR", "alloc entries");
  }
  if (emd_neigh_csr_fill(system->ctx, system->x, N_local, &g, binning->bincount, binning->binoffsets,
                         binning->permute_vector, neigh_cut, half_neigh_, comm_newton, neigh_offsets.ptr, neighs.ptr))
    fail("NeighborCSR", "fill");
  neigh_list.row_map = neigh_offsets.ptr;
  neigh_list.entries = neighs.ptr;
  neigh_list.N_local = N_local;
  neigh_list.total = total;
}

// src/neighbor_types/neighbor_2d.h:280-331
void Neighbor2D::create_neigh_list(System *system, Binning *binning, bool half_neigh_, bool) {
  const T_INT N_local = system->N_local;
  if (num_neighs_buf.extent() < (size_t)N_local + 1) {
    if (!num_neighs_buf.alloc((size_t)N_local + 1)) fail("Neighbor2D", "alloc num_neighs");
  }
  const emd_bin_geom g = binning->geom();
  fill_passes = 0;
  bool resize;
  bool fast = build_tiles(system, binning, neigh_cut);
  do {
    if (rows_cap < (size_t)N_local + 1 || cols_cap != (size_t)neigh_list.maxneighs) {
      rows_cap = (size_t)N_local + 1;
      cols_cap = (size_t)neigh_list.maxneighs;
      if (!neighs_buf.alloc(rows_cap * cols_cap)) fail("Neighbor2D", "alloc neighs");
    }
    int max_count = 0;
    int rc2d = 3;
    if (fast) {
      rc2d = emd_neigh_tiles_fill_2d(system->ctx, tile_lists, half_neigh_, comm_newton, neigh_list.maxneighs, num_neighs_buf.ptr,
                                     neighs_buf.ptr, &max_count);
      if (rc2d != 0 && rc2d != 3) fail("Neighbor2D", "tiles fill");
      if (rc2d == 3) { emd_tiles_invalidate(tile_lists); fast = false; }
    }
    if (rc2d == 0) {
    } else if (emd_neigh_2d_fill(system->ctx, system->x, N_local, &g, binning->bincount, binning->binoffsets,
                          binning->permute_vector, neigh_cut, half_neigh_, comm_newton, neigh_list.maxneighs,
                          num_neighs_buf.ptr, neighs_buf.ptr, &max_count))
      fail("Neighbor2D", "fill");
    fill_passes++;
    resize = max_count > neigh_list.maxneighs;
    // the reference takes the LAST overflowing row's count (racy, :210-216); the max is the
    // deterministic choice and never needs more passes
    if (resize) neigh_list.maxneighs = (T_INT)(max_count * 1.2);
  } while (resize);
  neigh_list.num_neighs = num_neighs_buf.ptr;
  neigh_list.neighs = neighs_buf.ptr;
  neigh_list.N_local = N_local;
  last_total = 0; // not tracked for 2D (would need a reduction); rows are exact in num_neighs
}
