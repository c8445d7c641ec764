/*
 * oracle.h -- CPU restatement of the ExaMiniMD per-timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under examinimd_b200/ may include, link or
 * execute this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference checkout, src/...).  The reference itself needs Kokkos (+MPI), neither
 * of which exists in this image, so the reference is not directly buildable; the
 * restatement is pinned against the unmodified reference translation units compiled
 * over a serial Kokkos-API shim (oracle/kokkos_shim -> oracle/_ref/, see
 * oracle/Makefile) and against the derived known answers of SURVEY.md App. C.
 * Kokkos::BinSort/BinOp3D (third-party kokkos/kokkos >=3.0, Kokkos_Sort.hpp, not
 * vendored by the reference) is restated from its published algorithm.
 *
 * Plain C11, no dependencies.  Compile with -ffp-contract=off: the reference is
 * built with plain -O3 for x86-64 (src/Makefile:19-20), i.e. without FMA contraction.
 */
#ifndef EMD_ORACLE_H
#define EMD_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_UNITS_REAL, ORC_UNITS_LJ, ORC_UNITS_METAL };
enum { ORC_LATTICE_SC, ORC_LATTICE_FCC };
enum { ORC_FORCE_LJ, ORC_FORCE_LJ_IDIAL, ORC_FORCE_SNAP };
enum { ORC_ITER_CELL_FULL, ORC_ITER_NEIGH_FULL, ORC_ITER_NEIGH_HALF };
enum { ORC_NEIGH_NONE, ORC_NEIGH_CSR, ORC_NEIGH_CSR_MAPCONSTR, ORC_NEIGH_2D };

/* src/system.h:57-143 (System), arrays row-major [N][3] */
typedef struct {
  int N, N_max, N_local, N_ghost, ntypes;
  double *x, *v, *f, *q;
  int *type, *id;
  double *mass; /* [ntypes] */
  double domain_x, domain_y, domain_z;
  double sub_domain_x, sub_domain_y, sub_domain_z;
  double sub_domain_lo_x, sub_domain_lo_y, sub_domain_lo_z;
  double sub_domain_hi_x, sub_domain_hi_y, sub_domain_hi_z;
  double boltz, mvv2e, dt;
} orc_system;

void orc_system_init(orc_system *s);
void orc_system_destroy(orc_system *s);
void orc_system_grow(orc_system *s, int n_new); /* src/system.cpp:95-109 */

/* src/input.h:135-183 (Input) -- the subset the hot path consumes */
#define ORC_MAX_COEFF_LINES 8
#define ORC_MAX_WORDS 32
#define ORC_WORD 32
typedef struct {
  int units, lattice_style;
  double lattice_constant, lattice_offset_x, lattice_offset_y, lattice_offset_z;
  int lattice_nx, lattice_ny, lattice_nz;
  double temperature_target;
  int temperature_seed;
  int nsteps, comm_exchange_rate, comm_newton;
  int force_type, force_iteration_type, neighbor_type;
  double force_cutoff, neighbor_skin;
  int thermo_rate;
  int timestepflag;
  char pair_style_words[ORC_MAX_WORDS][ORC_WORD];
  int n_coeff_lines;
  char coeff_words[ORC_MAX_COEFF_LINES][ORC_MAX_WORDS][ORC_WORD];
  int coeff_nwords[ORC_MAX_COEFF_LINES];
} orc_input;

void orc_input_defaults(orc_input *in);                                   /* src/input.cpp:120-149 */
int orc_input_read_deck(orc_input *in, orc_system *s, const char *file);  /* src/input.cpp:226-458 */
/* src/input.cpp:460-792, single rank (Comm::create_domain_decomposition, src/comm.cpp:56-63)
 * or brick `rank` of a px*py*pz grid (CommMPI::create_domain_decomposition,
 * src/comm_types/comm_mpi.cpp:52-147).  id_offset = exclusive prefix of N_local over ranks. */
void orc_create_lattice(const orc_input *in, orc_system *s);
void orc_lattice_positions(const orc_input *in, orc_system *s); /* positions/type/id only */
void orc_lattice_velocities_raw(const orc_input *in, orc_system *s, double *mom4); /* before momentum zero */
double orc_random_uniform_state(int *seed);
void orc_random_reset(int *seed, int ibase, const double *coord); /* src/input.h:66-133 */

/* src/binning.h:43-67 + src/binning_types/binning_kksort.cpp:71-140 */
typedef struct {
  int nbinx, nbiny, nbinz, nhalo;
  double minx, maxx, miny, maxy, minz, maxz;
  int *bincount, *binoffsets; /* flattened [nbinx][nbiny][nbinz] */
  int *permute_vector;
  int nbins_cap, perm_cap, range_begin, range_end;
} orc_binning;
void orc_binning_init(orc_binning *b);
void orc_binning_destroy(orc_binning *b);
void orc_create_binning(orc_binning *b, orc_system *s, double dx_in, double dy_in, double dz_in,
                        int halo_depth, int do_local, int do_ghost, int sort);

/* src/comm_types/comm_serial.{h,cpp} */
typedef struct {
  double comm_depth;
  int num_ghost[6];
  int *pack_indicies[6];
  int pack_cap[6];
} orc_comm_serial;
void orc_comm_serial_init(orc_comm_serial *c, double comm_depth);
void orc_comm_serial_destroy(orc_comm_serial *c);
void orc_comm_exchange(orc_comm_serial *c, orc_system *s);      /* comm_serial.cpp:47-54 */
void orc_comm_exchange_halo(orc_comm_serial *c, orc_system *s); /* comm_serial.cpp:56-97 */
void orc_comm_update_halo(orc_comm_serial *c, orc_system *s);   /* comm_serial.cpp:99-110 */
void orc_comm_update_force(orc_comm_serial *c, orc_system *s);  /* comm_serial.cpp:112-127 */

/* neighbor lists: CSR (src/neighbor_types/neighbor_csr.h) and 2D (neighbor_2d.h) share
 * one row accessor: row i = neighs + row_start(i), length num(i). */
typedef struct {
  int kind; /* ORC_NEIGH_CSR or ORC_NEIGH_2D */
  double neigh_cut;
  int comm_newton;
  int N_local;
  /* CSR */
  int *row_map; /* [N_local+1] */
  int *entries;
  int entries_cap, rows_cap;
  int total;
  /* 2D */
  int *num_neighs; /* [N_local+1] */
  int *neighs2d;   /* [N_local+1][maxneighs] row-major */
  int maxneighs, rows2d_cap, cols2d_cap;
  int n_fill_passes; /* how many times the 2D fill ran (neighbor_2d.h:304-330) */
} orc_neighbor;
void orc_neighbor_init(orc_neighbor *n, int kind, double neigh_cut);
void orc_neighbor_destroy(orc_neighbor *n);
void orc_create_neigh_list(orc_neighbor *n, const orc_system *s, const orc_binning *b, int half_neigh);
static inline const int *orc_neigh_row(const orc_neighbor *n, int i, int *count) {
  if (n->kind == ORC_NEIGH_2D) { *count = n->num_neighs[i]; return n->neighs2d + (long)i * n->maxneighs; }
  *count = n->row_map[i + 1] - n->row_map[i];
  return n->entries + n->row_map[i];
}

/* src/force_types/force_lj_neigh{.h,_impl.h} */
#define ORC_MAX_TYPES_STACKPARAMS 12
typedef struct {
  int ntypes, half_neigh, comm_newton;
  double *lj1, *lj2, *cutsq; /* [ntypes][ntypes] */
  double *intensity;         /* [ntypes][ntypes], ForceLJIDialNeigh only (idial != 0) */
  int idial;
} orc_force_lj;
void orc_force_lj_init(orc_force_lj *f, int ntypes, int half_neigh);
void orc_force_lj_destroy(orc_force_lj *f);
void orc_force_lj_init_coeff(orc_force_lj *f, int nargs, char args[][ORC_WORD]); /* _impl.h:57-98 */
void orc_force_lj_compute(const orc_force_lj *f, orc_system *s, const orc_neighbor *n); /* _impl.h:100-126,161-254 */
double orc_force_lj_energy(const orc_force_lj *f, const orc_system *s, const orc_neighbor *n); /* _impl.h:128-156,256-343 */
/* ForceLJIDialNeigh (pair_style lj/cut/idial), src/force_types/force_lj_idial_neigh_impl.h: init_coeff :50-88, functors :113-213.
 * No compute_energy (Force::compute_energy default: PE = 0). */
void orc_force_lj_idial_init_coeff(orc_force_lj *f, int nargs, char args[][ORC_WORD]);
void orc_force_lj_idial_compute(const orc_force_lj *f, orc_system *s, const orc_neighbor *n);

/* src/integrator_nve.cpp:41-121 */
void orc_initial_integrate(orc_system *s);
void orc_final_integrate(orc_system *s);

/* src/property_temperature.cpp:43-62, property_kine.cpp:43-61 */
double orc_temperature(const orc_system *s);
double orc_kine(const orc_system *s);

/* SNAP: src/force_types/force_snap_neigh_impl.h + sna_impl.hpp (oracle_snap.c) */
typedef struct orc_force_snap orc_force_snap;
orc_force_snap *orc_force_snap_create(int ntypes);
void orc_force_snap_destroy(orc_force_snap *f);
int orc_force_snap_init_coeff(orc_force_snap *f, int nargs, char args[][ORC_WORD], const char *dir);
void orc_force_snap_compute(orc_force_snap *f, orc_system *s, const orc_neighbor *n);
int orc_force_snap_ncoeff(const orc_force_snap *f);
double orc_force_snap_rcutmax(const orc_force_snap *f);
/* exposed for function-level parity tests */
int orc_snap_tables(const orc_force_snap *f, int *twojmax, int *idxj_max, int *idxj_full_max,
                    const double **cgarray, const double **rootpq, const double **coeffelem);
/* bispectrum pieces for one atom: rij[n][3], returns U_tot (r,i), and dB/dr per neighbour */
void orc_snap_atom(orc_force_snap *f, int ninside, const double *rij, const double *wj,
                   const double *rcutij, double *utot_r, double *utot_i, double *dbvec /* [n][ncoeff][3] */,
                   double *fij /* [n][3] */);

/* the same on the live system: atom i with its neighbor row; returns the number of in-cutoff neighbors */
int orc_force_snap_probe(orc_force_snap *f, const orc_system *s, const orc_neighbor *n, int i, double *utot_r, double *utot_i,
                         int *inside, double *fij);
int orc_force_snap_jdim(const orc_force_snap *f);

/* whole application: src/examinimd.cpp:60-294 (single rank, CommSerial) */
typedef struct {
  orc_input in;
  orc_system sys;
  orc_binning bin;
  orc_comm_serial comm;
  orc_neighbor neigh;
  orc_force_lj lj;
  orc_force_snap *snap;
  int step;
  double neigh_cutoff;
  /* wall-clock seconds per phase, as the reference's PERFORMANCE line */
  double t_force, t_neigh, t_comm, t_other;
} orc_md;
int orc_md_init(orc_md *md, const char *deck, int neighbor_type, int force_iteration_type, const char *cwd_for_coeff);
void orc_md_setup(orc_md *md); /* examinimd.cpp:117-144 after lattice creation */
void orc_md_step(orc_md *md);  /* one iteration of examinimd.cpp:192-250 */
void orc_md_thermo(orc_md *md, double *T, double *PE, double *KE); /* :252-255 */
void orc_md_destroy(orc_md *md);
int orc_dump_binary(const orc_system *s, const char *path, int step, int rank); /* examinimd.cpp:296-346 */

#ifdef __cplusplus
}
#endif
#endif
