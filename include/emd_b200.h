/*
 * emd_b200.h -- C ABI of the B200-native ExaMiniMD hot path (libemd_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry
 * point names the reference interface it replaces (paths relative to the ExaMiniMD checkout).
 * All `d_` pointers are DEVICE pointers on the context's GPU; `h_` pointers are host pointers.
 * Calls are stream-ordered on the context's stream; a call that returns a value through an
 * `h_` pointer synchronises that stream first (as the reference does with deep_copy-to-host).
 * Every function returns 0 on success, non-zero on error (message: emd_last_error()).
 * There is no CPU fallback: without a CUDA device emd_ctx_create fails.
 *
 * Per-atom arrays use the reference layout (src/types.h:88-113, src/system.h:59-93):
 * x,v,f = double[N][3] row-major, type,id = int[N], q = double[N], mass = double[ntypes].
 */
#ifndef EMD_B200_H
#define EMD_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define EMD_ABI_VERSION 1

typedef struct emd_ctx emd_ctx;

/* ---- context ------------------------------------------------------------------------- */
/* stream: a cudaStream_t to launch on, or NULL to create a private non-blocking stream. */
int emd_ctx_create(emd_ctx **out, int device, void *stream);
void emd_ctx_destroy(emd_ctx *ctx);
void *emd_ctx_stream(emd_ctx *ctx);
int emd_ctx_sync(emd_ctx *ctx);                 /* replaces Kokkos::fence() */
/* Halo gate: emd_peer_update_dim(..., defer_wait = 1) leaves the wait for the neighbours' ghost stores to the consumer.
 * The LJ tile force launches take a pending gate along (they wait only before their first tile that reads a ghost);
 * every other consumer of ghost positions calls emd_ctx_halo_gate_wait first (a one-warp wait kernel on the stream). */
int emd_ctx_set_halo_gate(emd_ctx *ctx, const int *d_arrived6, int seq, int phase_mask);
int emd_ctx_halo_gate_wait(emd_ctx *ctx);
int emd_ctx_halo_gate_pending(const emd_ctx *ctx);
/* measurement aid (bench.py): FP64 FMA peak of the device in TFLOP/s, DFMA loop on every SM, `reps` launches after two
 * warm-ups -- best and mean.  The SNAP roofline's denominator (MEASURED_PEAKS.json holds HBM and bf16 only). */
int emd_microbench_fp64(emd_ctx *ctx, int reps, double *h_tflops_best, double *h_tflops_mean);
const char *emd_last_error(void);
int emd_abi_version(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
unsigned long long emd_ctx_launch_count(emd_ctx *ctx);
/* device event timers on the context's stream: ms between tic and toc (synchronises) */
int emd_ctx_tic(emd_ctx *ctx);
int emd_ctx_toc(emd_ctx *ctx, float *h_ms);
/* opaque device events on the context's stream, for per-phase device timers
 * (the reference's Kokkos::Timer pairs, src/examinimd.cpp:183-189) without host syncs */
int emd_event_create(void **ev);
int emd_event_destroy(void *ev);
int emd_event_record(emd_ctx *ctx, void *ev);
int emd_event_elapsed_ms(void *ev_begin, void *ev_end, float *h_ms); /* synchronises on ev_end */
/* plain device memory helpers for hosts without a CUDA runtime binding of their own */
int emd_malloc(void **d_ptr, unsigned long long bytes);
int emd_free(void *d_ptr);
int emd_memcpy_h2d(emd_ctx *ctx, void *d_dst, const void *h_src, unsigned long long bytes);
int emd_memcpy_d2h(emd_ctx *ctx, void *h_dst, const void *d_src, unsigned long long bytes);
int emd_memcpy_d2d(emd_ctx *ctx, void *d_dst, const void *d_src, unsigned long long bytes);
int emd_memset_zero(emd_ctx *ctx, void *d_dst, unsigned long long bytes); /* Kokkos::deep_copy(view,0) */
/* A second, lower-priority stream for work that may overlap the module stream (the driver computes the halo-independent
 * part of the force on it while CommMPI::update_halo runs on the module stream).  side_begin: the side stream waits for
 * everything queued so far (or up to the last side_mark) and becomes the stream every entry point launches on; side_end: back to the module stream;
 * side_join: the module stream waits for the side work. */
int emd_ctx_side_mark(emd_ctx *ctx);  /* optional: fix the fork point now; the next side_begin waits for THIS point only */
int emd_ctx_side_begin(emd_ctx *ctx);
int emd_ctx_side_end(emd_ctx *ctx);
int emd_ctx_side_join(emd_ctx *ctx);
int emd_ctx_side_sms(const emd_ctx *ctx); /* SMs available to the current stream (side stream: its partition) */

/* ---- binning: BinningKKSort::create_binning, src/binning_types/binning_kksort.cpp:71-140 */
typedef struct {
  int nbinx, nbiny, nbinz, nhalo;          /* src/binning.h:52 (nbin* INCLUDE the 2*nhalo halo bins) */
  double minx, maxx, miny, maxy, minz, maxz; /* src/binning.h:53 */
} emd_bin_geom;

/* host arithmetic of binning_kksort.cpp:77-99, same expressions in the same order */
int emd_binning_geometry(const double sub_domain[3], const double sub_lo[3], const double sub_hi[3],
                         double dx_in, double dy_in, double dz_in, int halo_depth, emd_bin_geom *out);

/* Kokkos::BinSort<...,BinOp3D>::create_permute_vector (binning_kksort.cpp:103-111) +
 * AssignOffsets (:118-125).  d_x points at the first atom of the binned range and n is the
 * range length, so permute indices are relative to the range start like the reference's.
 * Outputs: d_bincount/d_binoffsets int[nbinx*nbiny*nbinz] flattened (x slowest, z fastest =
 * the reference's 3-D views), d_permute int[n].  Within a bin atoms are in ascending index
 * order (the reference's 1-thread arrival order).  Atoms whose bin falls outside the grid are
 * an error (reference: undefined behaviour). */
int emd_binning_build(emd_ctx *ctx, const double *d_x, int n, const emd_bin_geom *geom,
                      int *d_bincount, int *d_binoffsets, int *d_permute);

/* BinSort::sort(x), (v), (f), (type), (id), (q) (binning_kksort.cpp:126-138) as ONE gather
 * kernel: out[i] = in[permute[i]].  in/out must not alias (the host classes swap buffers). */
int emd_binning_permute(emd_ctx *ctx, const int *d_permute, int n,
                        const double *d_x_in, const double *d_v_in, const double *d_f_in,
                        const int *d_type_in, const int *d_id_in, const double *d_q_in,
                        double *d_x_out, double *d_v_out, double *d_f_out,
                        int *d_type_out, int *d_id_out, double *d_q_out);

/* ---- neighbor lists -------------------------------------------------------------------- */
/* NeighborCSR<>::create_neigh_list, src/neighbor_types/neighbor_csr.h:370-435.
 * Step 1 (count_neighbors_{full,half} :176-216/:258-309 + create_offsets :359-368):
 * fills d_row_map[0..n_local] and returns the total entry count in *h_total (host sync,
 * as neighbor_csr.h:413-414).  Step 2 (fill_neigh_list_{full,half} :219-255/:312-358) writes
 * d_entries[0..total).  Rows are in the reference's serial traversal order (27 stencil bins
 * bx-1..bx+1/by/bz, permute order inside each bin), so they compare equal entry by entry with
 * the 1-thread reference, not just as sets.
 * d_type may be NULL (the reference loads type_j but never uses it). */
int emd_neigh_csr_count(emd_ctx *ctx, const double *d_x, int n_local, const emd_bin_geom *geom,
                        const int *d_bincount, const int *d_binoffsets, const int *d_permute,
                        double neigh_cut, int half_neigh, int comm_newton,
                        int *d_row_map, int *h_total);
int emd_neigh_csr_fill(emd_ctx *ctx, const double *d_x, int n_local, const emd_bin_geom *geom,
                       const int *d_bincount, const int *d_binoffsets, const int *d_permute,
                       double neigh_cut, int half_neigh, int comm_newton,
                       const int *d_row_map, int *d_entries);

/* Neighbor2D<>::create_neigh_list fill pass, src/neighbor_types/neighbor_2d.h:175-278,304-318.
 * One pass into d_neighs[(n_local+1)][maxneighs] (row-major, stride maxneighs); entries beyond
 * maxneighs are dropped but still counted in d_num_neighs, and *h_max_count returns the largest
 * row count so the caller can apply the reference's resize rule (maxneighs = max*1.2, :322-326)
 * and call again. */
int emd_neigh_2d_fill(emd_ctx *ctx, const double *d_x, int n_local, const emd_bin_geom *geom,
                      const int *d_bincount, const int *d_binoffsets, const int *d_permute,
                      double neigh_cut, int half_neigh, int comm_newton,
                      int maxneighs, int *d_num_neighs, int *d_neighs, int *h_max_count);

/* A neighbor list as the force kernels consume it (the duck-typed list concept of
 * neighbor_csr.h:81-124 / neighbor_2d.h:80-120): row i starts at d_neighs + row_start(i).
 * CSR: d_row_map != NULL, row i = [row_map[i], row_map[i+1]).
 * 2D : d_row_map == NULL, row i = [i*stride, i*stride + d_num_neighs[i]). */
typedef struct {
  const int *d_row_map;
  const int *d_num_neighs;
  const int *d_neighs;
  int stride;
} emd_neigh_list;

/* ---- LJ force: ForceLJNeigh<>, src/force_types/force_lj_neigh_impl.h -------------------- */
/* init_coeff (:57-98) result, host arrays [ntypes][ntypes]: lj1 = 48 eps sigma^12,
 * lj2 = 24 eps sigma^6, cutsq = rc^2.  Copied into the context. */
int emd_force_lj_set_params(emd_ctx *ctx, int ntypes, const double *h_lj1, const double *h_lj2,
                            const double *h_cutsq);
/* compute (:100-126; functors :161-206 full, :208-254 half).  Accumulates onto d_f (+=), which
 * the caller has zeroed like examinimd.cpp:232; with zero_f != 0 the kernel itself overwrites
 * rows [0,n_local) and zeroes rows [n_local,n_all) first, replacing that deep_copy.
 * half: f_j -= for every listed j including ghosts (:245-247). */
int emd_force_lj_compute(emd_ctx *ctx, const double *d_x, const int *d_type, double *d_f,
                         int n_local, int n_all, const emd_neigh_list *list, int half_neigh,
                         int zero_f);
/* ForceLJIDialNeigh<>::compute, src/force_types/force_lj_idial_neigh_impl.h:90-111 (functors :113-163 full, :165-213 half):
 * the LJ pair force accumulated intensity(ti,tj) times, each term divided by it; h_intensity = host [ntypes][ntypes] table
 * set by init_coeff (:50-88); lj1/lj2/cutsq come from emd_force_lj_set_params.  Half lists subtract from j only if j is
 * owned (:203-207).  zero_f as above.  The module has no energy (Force::compute_energy default, src/force.h:54). */
int emd_force_lj_idial_compute(emd_ctx *ctx, const double *d_x, const int *d_type, double *d_f,
                               int n_local, int n_all, const emd_neigh_list *list, int half_neigh,
                               int zero_f, const double *h_intensity);
/* compute_energy (:128-156; functors :256-296 full, :298-343 half): shifted PE, host result */
int emd_force_lj_energy(emd_ctx *ctx, const double *d_x, const int *d_type, int n_local,
                        const emd_neigh_list *list, int half_neigh, double *h_pe);

/* ---- Input::create_lattice / create_velocities on the device (src/input.cpp:460-792, src/input.h:66-133), bit-identical
 * to the host loops: sites in the reference's loop order (z, y, x, basis) that lie in the brick [lo, hi); velocities from the
 * per-position hashed Park-Miller stream; momentum / temperature sums in atom order (one thread). */
typedef struct emd_lattice {
  long long i0[3];   /* first lattice index per dimension (input.cpp:485-490 / 604-609) */
  int n[3];          /* number of indices per dimension (inclusive ranges of the reference) */
  int fcc;           /* 1: fcc (4-atom basis, a*(1.0*i + basis + offset)); 0: sc (a*(i + offset)) */
  double a, offset[3];
  double lo[3], hi[3];
} emd_lattice;
int emd_lattice_count(emd_ctx *ctx, const emd_lattice *lat, int *h_n);
/* follows emd_lattice_count of the same lattice; one atom type (type 0); id = row + 1 + id_offset; q = 0; v = (u - 0.5)/sqrt(m) */
int emd_lattice_fill(emd_ctx *ctx, const emd_lattice *lat, int seed, int id_offset, const double *d_mass, double *d_x, double *d_v,
                     double *d_q, int *d_type, int *d_id);
/* mode 0: {sum m, sum m vx, sum m vy, sum m vz} (input.cpp:750-753); mode 1: {sum m |v|^2} (property_temperature.cpp:49) */
int emd_velocity_sums(emd_ctx *ctx, const double *d_v, const int *d_type, const double *d_mass, int n, int mode, double *h_out4);
int emd_velocity_shift(emd_ctx *ctx, double *d_v, int n, double sx, double sy, double sz);   /* input.cpp:763-767 */
int emd_velocity_scale(emd_ctx *ctx, double *d_v, int n, double s);                          /* input.cpp:781-785 */

/* ---- tile lists: the B200 fast path of the neighbor build + LJ force (kernels/tiles.cu) -------
 * An emd_tiles object holds a tile-local FULL adjacency (shared-memory slot numbers, ELL layout)
 * built from the same inputs as the reference lists: an FP32 search leaves one bit per stencil
 * candidate; the reference-visible CSR / 2D lists are made FROM those bits with the reference's
 * exact FP64 inclusion rules, so they are bit-identical to emd_neigh_csr_count/fill and
 * emd_neigh_2d_fill; the LJ force then runs on the tile lists with every x[j] served from shared
 * memory and no atomics (each pair is evaluated from both sides).
 * emd_neigh_tiles_build / _count return 0 on success and 3 when the fast path does not apply to
 * this configuration (bins narrower than the list radius, a tile that does not fit in shared
 * memory): the caller then uses the generic entry points above.  A re-neighboring synchronises
 * with the host once (the row total of emd_neigh_tiles_count, read back with the overflow flags).  The binning arrays passed to build must
 * stay alive and unchanged until the next build (they are re-read by every later call). */
typedef struct emd_tiles emd_tiles;
int emd_tiles_create(emd_tiles **out);
void emd_tiles_destroy(emd_tiles *t);
int emd_tiles_valid(const emd_tiles *t);
void emd_tiles_invalidate(emd_tiles *t);
int emd_tiles_info(const emd_tiles *t, int *tile_dims3, int *ntiles, int *stride, int *maxrow, int *cap);
/* Device pointers of the tile lists (inspection / tests).  csr16 / ncsr: the EXACT rows of the list type last asked for
 * (emd_neigh_tiles_count / fill_*), in the reference's order, as staged-slot numbers ([ntiles][maxrow/8][stride][8] uint16,
 * row lengths [ntiles][stride]); NULL if no exact list was made of this build.  ell_s / nell_s: the force kernel's rows, a
 * conservative FP32 superset of the full list re-ordered into shared-memory bank-conflict-free columns; an entry is the byte
 * offset 24*(16 + slot) of the neighbor's staged coordinates, entries < 24*16 are padding (16 dummy atoms lead the buffer).  int_slot: [ntiles][stride] staged
 * slot of the row's own atom; stg_j: [ntiles][cap] atom index of every staged slot.  Any pointer argument may be NULL. */
int emd_tiles_lists(const emd_tiles *t, const unsigned short **d_csr16, const int **d_ncsr, const unsigned short **d_ell_s,
                    const int **d_nell_s, const unsigned short **d_int_slot, const int **d_stg_j);
int emd_neigh_tiles_build(emd_ctx *ctx, emd_tiles *t, const double *d_x, int n_local, int n_all,
                          const emd_bin_geom *geom, const int *d_bincount, const int *d_binoffsets,
                          const int *d_permute, double neigh_cut);
/* = emd_neigh_csr_count / emd_neigh_csr_fill / emd_neigh_2d_fill, from the tile lists */
int emd_neigh_tiles_count(emd_ctx *ctx, emd_tiles *t, int half_neigh, int comm_newton, int *d_row_map,
                          int *h_total);
int emd_neigh_tiles_fill_csr(emd_ctx *ctx, emd_tiles *t, int half_neigh, int comm_newton,
                             const int *d_row_map, int *d_entries);
int emd_neigh_tiles_fill_2d(emd_ctx *ctx, emd_tiles *t, int half_neigh, int comm_newton, int maxneighs,
                            int *d_num_neighs, int *d_neighs, int *h_max_count);
/* ForceLJNeigh::compute (h_pe == NULL: overwrites d_f rows [0,n_local), ghost rows untouched) or
 * ::compute_energy (h_pe != NULL: shifted PE, d_f untouched), force_lj_neigh_impl.h:100-156.
 * Valid for full lists and for half lists with newton off (same forces on owned atoms). */
int emd_force_lj_compute_tiles(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type,
                               double *d_f, double *h_pe);
/* ::compute and ::compute_energy in ONE pass over the pairs (the driver asks for it on the steps whose thermo output follows:
 * src/examinimd.cpp:232-235 then :252-267 evaluate the same pairs twice): d_f as by compute, *h_pe as by compute_energy. */
int emd_force_lj_compute_tiles_with_energy(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type,
                                           double *d_f, double *h_pe);
/* The same force in two launches, for a decomposed run (CommMPI): part 1 = the tiles whose staged cells hold only owned
 * atoms (no dependence on this step's halo exchange, src/comm_types/comm_mpi.cpp:382-423), part 2 = the tiles that read
 * ghosts; part 0 = all.  The two parts write disjoint rows of d_f.  reserve_ctas leaves CTA slots of the persistent grid
 * free for the exchange's pack / transport kernels.  emd_tiles_halo_split reports the two tile counts. */
int emd_force_lj_compute_tiles_part(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type,
                                    double *d_f, int part, int reserve_ctas);
int emd_tiles_halo_split(const emd_tiles *t, int *n_free_tiles, int *n_halo_tiles);
/* 1 if every owned atom has a row in the tile lists (precondition of the *_nve launches) */
int emd_tiles_complete(const emd_tiles *t, int *all_owned_have_rows);
/* ForceLJNeigh::compute followed, per owned atom and in the same launch, by IntegratorNVE::final_integrate of this step and
 * initial_integrate of the next (src/integrator_nve.cpp:77-83,115-121; legal when nothing observes x, v, f between the two
 * steps): d_f and d_v are updated as by the separate calls, the advanced positions are written to d_x_new (rows [0,n_local));
 * d_x is left untouched -- the caller publishes d_x_new as the position array afterwards.  Bit-identical to the three calls.
 * Returns 3 (nothing done) if an owned atom has no row in the tile lists (it sits outside the interior bins). */
int emd_force_lj_compute_tiles_nve(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f,
                                   double *d_v, double *d_x_new, const double *d_mass, double dtf, double dtv);
/* the same on a thermo step: the pass also returns *h_pe = the potential energy of the owned atoms at d_x (as
 * emd_force_lj_compute_tiles_with_energy) and *h_mv2 = sum m v^2 over the owned atoms of the velocities BETWEEN the two kicks,
 * i.e. what Temperature / KinE (src/property_temperature.cpp:43-62, property_kine.cpp:43-61) sum after final_integrate; the
 * thermo step then keeps the fused integrator and needs no reduction pass of its own */
int emd_force_lj_compute_tiles_nve_thermo(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f,
                                          double *d_v, double *d_x_new, const double *d_mass, double dtf, double dtv,
                                          double *h_pe, double *h_mv2);
/* the same for one part of a split force (emd_force_lj_compute_tiles_part): every owned atom belongs to exactly one part */
int emd_force_lj_compute_tiles_part_nve(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f,
                                        int part, int reserve_ctas, double *d_v, double *d_x_new, const double *d_mass,
                                        double dtf, double dtv);

/* ---- SNAP force: ForceSNAP<> + SNA, src/force_types/force_snap_neigh_impl.h, sna_impl.hpp ------- */
/* What init_coeff/read_files (force_snap_neigh_impl.h:227-336, 340-587) leave behind, as plain values.
 * elem_of_type[t] = map[t] exactly as the reference's kernel reads it (:597: indexed by the 0-based atom
 * type); radelem/wjelem/coeffelem are HOST arrays, copied.  Only diagonalstyle 3 (the only style for
 * which the reference builds its index lists, sna_impl.hpp:86-132) and linear SNAP (quadraticflag 0). */
typedef struct {
  int twojmax, switchflag, ntypes, nelements, ncoeffall;
  double rcutfac, rfac0, rmin0, wself;
  int elem_of_type[12];
  const double *radelem;   /* [nelements] */
  const double *wjelem;    /* [nelements] */
  const double *coeffelem; /* [nelements][ncoeffall], coeffelem[e][0] = the constant term (unused by forces) */
} emd_snap_params;
typedef struct emd_snap emd_snap;
/* SNA::SNA + SNA::init (sna_impl.hpp:25-53,136-140): index lists, Clebsch-Gordan and sqrt(p/q) tables,
 * plus the beta-folded block coefficients of the adjoint formulation; uploads them to the device */
int emd_snap_create(emd_snap **out, const emd_snap_params *p);
/* host-only self-check of snap_yi's work plan (items / segments / step table built by emd_snap_create): every (block, output
 * row, ma1) of compute_zi (sna_impl.hpp:196-283) is covered exactly once; returns the plan's sizes, the Clebsch-Gordan terms it
 * executes and the non-zero ones among them.  Needs no device. */
int emd_snap_yi_plan_stats(int twojmax, int *nitems, int *nsegs, int *ntab, long long *terms_executed, long long *terms_nonzero);
void emd_snap_destroy(emd_snap *s);
/* ncoeff (sna.ncoeff), U/Y elements per atom on the half range, Z blocks, rcutmax (:321-327), and of the
 * last compute: in-cutoff pairs and the atom stride of the U array.  Any pointer may be NULL. */
int emd_snap_info(const emd_snap *s, int *ncoeff, int *nuh, int *ntriples, double *rcutmax, int *npairs, int *ustride);
/* device arrays of the last compute, for function-level parity tests: "ulist" double[nuh][ustride][2]
 * (U_tot on mb <= j/2, element (j,mb,ma) at block(j) + mb*(j+1) + ma), "ylist" double[n_local][nuh][2]
 * (weighted adjoint Y), "pair_i"/"pair_j" int[npairs], "pair_offsets" int[n_local+1] */
void *emd_snap_device_ptr(emd_snap *s, const char *what);
/* ForceSNAP::compute (force_snap_neigh_impl.h:159-208 + team kernel :589-725) over a FULL neighbor list
 * with newton on: f_i += F_ij and f_j -= F_ij for every listed j with rsq < rcutmax^2, ghosts included
 * (the caller folds ghost forces back with update_force, examinimd.cpp:241-245).  Accumulates onto d_f,
 * which the caller has zeroed (examinimd.cpp:232).  Synchronises once (pair count read-back), like the
 * reference's max_neighs reduction (:181). */
int emd_force_snap_compute(emd_ctx *ctx, emd_snap *s, const double *d_x, const int *d_type, double *d_f,
                           int n_local, int n_all, const emd_neigh_list *list);
/* SNAP potential energy of the owned atoms -- NOT in the reference: ForceSNAP inherits Force::compute_energy, which
 * returns 0 (src/force.h:54).  E = sum_i [ beta_0 + sum_k beta_k (B_k(i) - bzero_k) ] +
 * sum_i sum_{j: rsq < rcutmax^2} 1.25e5 / r_ij^12, where B_k = 2 sum_{mb <= j/2, ma} w Re(conj(U_tot) Z_k) is the
 * bispectrum component of the published SNA::compute_bi (LAMMPS src/SNAP/sna.cpp, the code sna_impl.hpp was
 * taken from; w = 1, 1/2 on the diagonal of an even level's middle column, 0 below it), bzero_k = wself^3 (j+1)
 * if bzeroflag, and the last term is the potential of the rij * (-1.5e6 / r^14) the force kernel adds from both
 * ends of a pair (force_snap_neigh_impl.h:698-711).  -dE/dx equals the force of emd_force_snap_compute
 * (tests/test_gpu_snap.py: finite differences).  Recomputes the pair list and U_tot from d_x; overwrites the Y
 * array of the last force call.  *h_energy is valid on return. */
int emd_force_snap_energy(emd_ctx *ctx, emd_snap *s, const double *d_x, const int *d_type, int n_local,
                          const emd_neigh_list *list, int bzeroflag, double *h_energy);

/* ---- integrator: IntegratorNVE, src/integrator_nve.cpp:41-121 --------------------------- */
/* dtf = 0.5*dt/mvv2e, dtv = dt (:41-44).  Bit-exact with the reference's CPU arithmetic
 * (separate multiply and add, no FMA contraction). */
int emd_nve_initial_integrate(emd_ctx *ctx, double *d_x, double *d_v, const double *d_f,
                              const int *d_type, const double *d_mass, int n_local,
                              double dtf, double dtv);
int emd_nve_final_integrate(emd_ctx *ctx, double *d_v, const double *d_f, const int *d_type,
                            const double *d_mass, int n_local, double dtf);
/* final_integrate of one step followed by initial_integrate of the next (nothing between the two reads or writes x, v
 * or f when no thermo output / dump sits between the steps): one pass over the atoms, bit-identical results. */
int emd_nve_final_initial_integrate(emd_ctx *ctx, double *d_x, double *d_v, const double *d_f, const int *d_type,
                                    const double *d_mass, int n_local, double dtf, double dtv);

/* ---- single-process periodic comm: CommSerial, src/comm_types/comm_serial.{h,cpp} ------- */
/* exchange (comm_serial.cpp:47-54, TagExchangeSelf comm_serial.h:94-107) */
int emd_comm_wrap(emd_ctx *ctx, double *d_x, int n_local, const double domain[3]);
/* one phase of exchange_halo (comm_serial.cpp:62-94, TagHaloSelf comm_serial.h:110-181).
 * Scans atoms [0,n_scan) for x_dim >= hi-depth (even phase) / <= lo+depth (odd phase), appends
 * a shifted copy (x,v,q,id,type) of each hit at ghost_begin+slot and records the source index
 * in d_pack_indicies[slot]; slots are in ascending source index (the 1-thread arrival order).
 * *h_count returns the hit count (host sync, as comm_serial.cpp:73).  Nothing is written beyond
 * `capacity` atoms / `pack_capacity` indices: the caller grows and redoes, as the reference. */
int emd_comm_halo_phase(emd_ctx *ctx, int phase, double *d_x, double *d_v, double *d_q, int *d_id,
                        int *d_type, int n_scan, int ghost_begin, int capacity,
                        int *d_pack_indicies, int pack_capacity, const double domain[3],
                        const double sub_lo[3], const double sub_hi[3], double comm_depth,
                        int *h_count);
/* one phase of update_halo (comm_serial.cpp:99-110, TagHaloUpdateSelf comm_serial.h:183-197) */
int emd_comm_halo_update_phase(emd_ctx *ctx, int phase, double *d_x, double *d_v, double *d_q,
                               int *d_id, int *d_type, const int *d_pack_indicies, int count,
                               int ghost_begin, const double domain[3]);
/* update_halo (comm_serial.cpp:99-110) as ONE kernel.  emd_comm_halo_resolve follows, once per ghost build, every ghost's
 * chain of sources through the six pack lists (counts[p] ghosts per phase, stored behind n_local in phase order) down to
 * its owned root atom and total periodic shift (d_root int[n_ghost], d_shift double[n_ghost][3]); emd_comm_halo_refresh
 * then rewrites all ghost positions, x[ghost] = x[root] + shift: the same values as the six dependent phase copies
 * (positions only: the other ghost fields do not change between ghost builds). */
int emd_comm_halo_resolve(emd_ctx *ctx, const int *const d_pack_indicies[6], const int counts[6], int n_local,
                          const double domain[3], int *d_root, double *d_shift);
int emd_comm_halo_refresh(emd_ctx *ctx, double *d_x, int n_local, int n_ghost, const int *d_root, const double *d_shift);
/* one phase of update_force (comm_serial.cpp:112-127, TagHaloForceSelf comm_serial.h:199-213) */
int emd_comm_force_fold_phase(emd_ctx *ctx, double *d_f, const int *d_pack_indicies, int count,
                              int ghost_begin);

/* ---- multi-GPU comm: CommMPI, src/comm_types/comm_mpi.{h,cpp} ----------------------------------
 * One process per GPU; the host class (csrc/host/comm_types/comm_mpi.cpp) sequences the six phases and
 * owns the buffers, these entry points are the pack/unpack kernels and the transport (kernels/comm_mpi.cu). */
typedef struct {
  int nranks, rank;
  int grid[3], pos[3];                   /* proc_grid, proc_pos (comm_mpi.cpp:58-92) */
  int neighbor_send[6], neighbor_recv[6]; /* per phase 0:+x 1:-x 2:+y 3:-y 4:+z 5:-z; -1 = dimension not decomposed (:94-130) */
  double sub[3], sub_lo[3], sub_hi[3];   /* brick extent and bounds (:132-140) */
} emd_decomp;
/* CommMPI::create_domain_decomposition (comm_mpi.cpp:52-147): minimum-surface processor grid (first found wins
 * on ties), rank = x + px*(y + py*z), periodic neighbors, brick bounds.  Pure host arithmetic (no GPU needed). */
int emd_comm_decompose(int nranks, int rank, const double domain[3], emd_decomp *out);
/* TagExchangeSelf (comm_mpi.h:134-153): periodic wrap in the dimensions with wrap_dim[d] != 0 (grid[d] == 1) */
int emd_comm_wrap_dims(emd_ctx *ctx, double *d_x, int n_local, const double domain[3], const int wrap_dim[3]);
/* TagExchangePack (comm_mpi.h:155-226): atoms of [0,n_scan) with type >= 0 strictly beyond the phase's face are
 * packed (72-byte Particles, ascending index, shifted by the box length on the boundary rank) and marked
 * type = -1.  *h_count = how many (host sync, comm_mpi.cpp:229); nothing is written if count > capacity. */
int emd_comm_exchange_pack(emd_ctx *ctx, int phase, const emd_decomp *dec, const double domain[3], double *d_x, double *d_v,
                           double *d_q, int *d_id, int *d_type, int n_scan, void *d_pack_buffer, int capacity, int *h_count);
/* TagHaloPack (comm_mpi.h:349-430): atoms of [0,n_scan) within comm_depth of the phase's face; also records the
 * source indices for update_halo / update_force */
int emd_comm_halo_pack(emd_ctx *ctx, int phase, const emd_decomp *dec, const double domain[3], double comm_depth, double *d_x,
                       double *d_v, double *d_q, int *d_id, int *d_type, int n_scan, int *d_pack_indicies, void *d_pack_buffer,
                       int capacity, int *h_count);
/* TagUnpack (comm_mpi.h:432-436): set_particle(dst_begin + i, buffer[i]) */
int emd_comm_unpack(emd_ctx *ctx, const void *d_unpack_buffer, int count, int dst_begin, double *d_x, double *d_v, double *d_q,
                    int *d_id, int *d_type);
/* TagExchangeCreateDestList + TagExchangeCompact (comm_mpi.h:229-250, comm_mpi.cpp:262-284): the holes (type < 0)
 * below n_new are filled with the live atoms of [n_new, n_end), taken from the end downwards */
int emd_comm_exchange_compact(emd_ctx *ctx, double *d_x, double *d_v, double *d_q, int *d_id, int *d_type, int n_new, int n_end);
/* TagHaloUpdatePack / TagHaloUpdateUnpack (comm_mpi.h:462-490): positions only, 24 B per ghost */
int emd_comm_halo_update_pack(emd_ctx *ctx, int phase, const emd_decomp *dec, const double domain[3], const double *d_x,
                              const int *d_pack_indicies, int count, double *d_buffer);
int emd_comm_halo_update_unpack(emd_ctx *ctx, double *d_x, int ghost_begin, int count, const double *d_buffer);
/* TagHaloForceUnpack (comm_mpi.h:487-497): f[pack_indicies[ii]] += buffer[ii].  (TagHaloForcePack needs no kernel:
 * the ghost rows of f are contiguous and are sent in place.) */
int emd_comm_force_unpack(emd_ctx *ctx, double *d_f, const int *d_pack_indicies, int count, const double *d_buffer);

/* transport: NCCL send/recv over NVLink on the context's stream (takes the place of MPI in comm_mpi.cpp) */
typedef struct emd_net emd_net;
/* 128-byte NCCL unique id, for hosts that distribute it themselves (e.g. through torch.distributed) */
int emd_net_unique_id(void *out128);
/* unique_id128 == NULL: rank 0 creates the id and hands it to the others over TCP (MASTER_ADDR, MASTER_PORT+29
 * or EMD_RENDEZVOUS_PORT) */
int emd_net_create(emd_net **out, emd_ctx *ctx, int nranks, int rank, const void *unique_id128);
void emd_net_destroy(emd_net *n);
/* one phase's message pair (MPI_Irecv + MPI_Send + MPI_Wait, comm_mpi.cpp:245-251,333-337,401-404,449-452),
 * stream-ordered, no host synchronisation; either side may be 0 bytes */
int emd_net_sendrecv(emd_net *n, const void *d_send, unsigned long long send_bytes, int peer_send, void *d_recv,
                     unsigned long long recv_bytes, int peer_recv);
/* message pairs issued between begin and end travel as one NCCL group (one fused send/recv kernel on the stream) */
int emd_net_group_begin(emd_net *n);
int emd_net_group_end(emd_net *n);
/* the handshakes of the two phases of one dimension (which do not depend on each other) in one group; synchronises */
int emd_net_exchange_counts2(emd_net *n, const int send_count[2], const int peer_send[2], const int peer_recv[2], int h_recv_count[2]);
/* MPI_Allgather of one small record per rank, host memory to host memory (nbytes a multiple of 4); synchronises */
int emd_net_allgather_bytes(emd_net *n, const void *h_in, int nbytes, void *h_out_all);
/* ---- CommMPI::update_halo (comm_mpi.cpp:382-423) by peer stores over NVLink (kernels/comm_peer.cu): the pack kernel of a
 * phase writes the shifted positions straight into the neighbour's ghost rows through CUDA IPC mappings and raises a
 * sequence flag there; no transport library on the per-step path.  emd_peer_publish is collective and follows every
 * exchange_halo (it ships the IPC handles of the two position arrays and the first ghost row of each phase);
 * per refresh: emd_peer_begin_update once, then emd_peer_update_dim for every decomposed dimension in order 0,1,2
 * (the dimensions that are not decomposed use emd_comm_halo_update_phase in between, as before). */
typedef struct emd_peer emd_peer;
int emd_peer_create(emd_peer **out, emd_net *net, emd_ctx *ctx, int nranks, int rank);
void emd_peer_destroy(emd_peer *p);
int emd_peer_publish(emd_peer *p, const emd_decomp *dec, double *d_x0, double *d_x1, const int ghost_begin[6]);
int emd_peer_ready(const emd_peer *p);
int emd_peer_begin_update(emd_peer *p, const emd_decomp *dec, const double *d_x_current);
int emd_peer_update_dim(emd_peer *p, const emd_decomp *dec, const double domain[3], int dim, const double *d_x,
                        const int *d_pack_idx_a, int count_a, const int *d_pack_idx_b, int count_b, int defer_wait);
/* blocks the stream until every message of the refresh in progress has landed (needed before a dimension that is not
 * decomposed forwards ghosts of an earlier, decomposed one) */
int emd_peer_wait_all(emd_peer *p);
/* the count handshake of a phase (comm_mpi.cpp:235-238, 325-328); synchronises */
int emd_net_exchange_count(emd_net *n, int send_count, int peer_send, int peer_recv, int *h_recv_count);
/* MPI_Allreduce(IN_PLACE) / MPI_Scan on HOST scalars (comm_mpi.cpp:150-191): is_double 0 int / 1 double, op 0 sum / 1 max */
int emd_net_allreduce(emd_net *n, void *h_values, int count, int is_double, int op);
int emd_net_scan_int(emd_net *n, int *h_value);
int emd_net_barrier(emd_net *n);

/* ---- thermo: Temperature/KinE functor, src/property_temperature.h:55-57 ------------------ */
/* sum_i m[type_i] * |v_i|^2 over [0,n_local) (deterministic two-stage reduction), host result */
int emd_reduce_mv2(emd_ctx *ctx, const double *d_v, const int *d_type, const double *d_mass,
                   int n_local, double *h_sum);

#ifdef __cplusplus
}
#endif
#endif
