/*
 * oracle_md.c -- CPU restatement of the ExaMiniMD LJ hot path (TEST INFRASTRUCTURE ONLY,
 * see oracle.h).  Serial loops follow the reference's CPU (OpenMP back-end, 1 thread)
 * iteration order so that orders which the reference leaves to atomic arrival are the
 * deterministic ascending-index ones.  OpenMP pragmas are enabled only with
 * -DORC_OPENMP (the timed CPU baseline); parity tests use the serial build.
 */
#include "oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ------------------------------------------------------------------ System */
/* src/system.cpp:42-75 */
void orc_system_init(orc_system *s) {
  memset(s, 0, sizeof(*s));
  s->ntypes = 1;
  s->mass = (double *)calloc(1, sizeof(double));
}
void orc_system_destroy(orc_system *s) {
  free(s->x); free(s->v); free(s->f); free(s->q); free(s->type); free(s->id); free(s->mass);
  memset(s, 0, sizeof(*s));
}
static void *grow_zero(void *p, size_t old_bytes, size_t new_bytes) {
  /* Kokkos::resize preserves contents and zero-initialises the new tail */
  char *q = (char *)realloc(p, new_bytes);
  if (new_bytes > old_bytes) memset(q + old_bytes, 0, new_bytes - old_bytes);
  return q;
}
/* src/system.cpp:95-109 */
void orc_system_grow(orc_system *s, int n_new) {
  if (n_new > s->N_max) {
    size_t o = (size_t)s->N_max, n = (size_t)n_new;
    s->x = (double *)grow_zero(s->x, o * 24, n * 24);
    s->v = (double *)grow_zero(s->v, o * 24, n * 24);
    s->f = (double *)grow_zero(s->f, o * 24, n * 24);
    s->id = (int *)grow_zero(s->id, o * 4, n * 4);
    s->type = (int *)grow_zero(s->type, o * 4, n * 4);
    s->q = (double *)grow_zero(s->q, o * 8, n * 8);
    s->N_max = n_new;
  }
}

/* ------------------------------------------------------------------- Input */
/* src/input.cpp:120-149 */
void orc_input_defaults(orc_input *in) {
  memset(in, 0, sizeof(*in));
  in->neighbor_type = ORC_NEIGH_2D;
  in->force_iteration_type = ORC_ITER_NEIGH_FULL;
  in->comm_exchange_rate = 20;
  in->thermo_rate = 0;
  in->comm_newton = 0;
}

/* tokenizer: src/input.cpp:100-118 (<=32 words of <=31 chars, split on space/tab) */
static int tokenize(const char *line, char words[ORC_MAX_WORDS][ORC_WORD]) {
  const char *pos = line;
  int j = 0;
  for (int w = 0; w < ORC_MAX_WORDS; w++) words[w][0] = 0;
  while (*pos && j < ORC_MAX_WORDS) {
    while ((*pos == ' ' || *pos == '\t') && *pos) pos++;
    int k = 0;
    while (*pos != ' ' && *pos != '\t' && *pos && k < ORC_WORD) {
      /* the reference writes up to k==32 then the NUL at [32] (one past); we stop at 31 */
      if (k < ORC_WORD - 1) words[j][k] = *pos;
      k++; pos++;
    }
    words[j][k < ORC_WORD - 1 ? k : ORC_WORD - 1] = 0;
    j++;
  }
  int count = 0;
  for (int w = 0; w < ORC_MAX_WORDS; w++) if (words[w][0]) count++;
  return count;
}

/* src/input.cpp:265-458 */
int orc_input_read_deck(orc_input *in, orc_system *s, const char *file) {
  FILE *fp = fopen(file, "r");
  if (!fp) return -1;
  char line[512];
  int nlines = 0;
  while (fgets(line, 511, fp)) {
    size_t L = strlen(line);
    while (L && (line[L - 1] == '\n' || line[L - 1] == '\r')) line[--L] = 0;
    if (nlines++ >= 100) break; /* input.cpp:239 allocate_words(100) */
    char w[ORC_MAX_WORDS][ORC_WORD];
    int nw = tokenize(line, w);
    if (w[0][0] == 0 || strchr(w[0], '#')) continue;
    if (!strcmp(w[0], "units")) { /* :275-303 */
      if (!strcmp(w[1], "metal")) {
        in->units = ORC_UNITS_METAL; s->boltz = 8.617343e-5; s->mvv2e = 1.0364269e-4; s->dt = 0.001;
      } else if (!strcmp(w[1], "real")) {
        in->units = ORC_UNITS_REAL; s->boltz = 0.0019872067; s->mvv2e = 48.88821291 * 48.88821291;
        if (!in->timestepflag) s->dt = 1.0;
      } else if (!strcmp(w[1], "lj")) {
        in->units = ORC_UNITS_LJ; s->boltz = 1.0; s->mvv2e = 1.0;
        if (!in->timestepflag) s->dt = 0.005;
      }
    } else if (!strcmp(w[0], "lattice")) { /* :312-331 */
      if (!strcmp(w[1], "sc")) { in->lattice_style = ORC_LATTICE_SC; in->lattice_constant = atof(w[2]); }
      else if (!strcmp(w[1], "fcc")) { in->lattice_style = ORC_LATTICE_FCC; in->lattice_constant = pow(4.0 / atof(w[2]), 1.0 / 3.0); }
      if (!strcmp(w[3], "origin")) { in->lattice_offset_x = atof(w[4]); in->lattice_offset_y = atof(w[5]); in->lattice_offset_z = atof(w[6]); }
    } else if (!strcmp(w[0], "region")) { /* :332-353 */
      if (!strcmp(w[2], "block")) { in->lattice_nx = atoi(w[4]); in->lattice_ny = atoi(w[6]); in->lattice_nz = atoi(w[8]); }
    } else if (!strcmp(w[0], "create_box")) { /* :354-358 */
      s->ntypes = atoi(w[1]);
      free(s->mass);
      s->mass = (double *)calloc((size_t)s->ntypes, sizeof(double));
    } else if (!strcmp(w[0], "mass")) { /* :362-368 */
      s->mass[atoi(w[1]) - 1] = atof(w[2]);
    } else if (!strcmp(w[0], "pair_style")) { /* :369-390 */
      if (!strcmp(w[1], "lj/cut/idial")) { in->force_type = ORC_FORCE_LJ_IDIAL; in->force_cutoff = atof(w[2]); }
      else if (!strcmp(w[1], "lj/cut")) { in->force_type = ORC_FORCE_LJ; in->force_cutoff = atof(w[2]); }
      if (!strcmp(w[1], "snap")) { in->force_type = ORC_FORCE_SNAP; in->force_cutoff = 4.73442; }
      memcpy(in->pair_style_words, w, sizeof(w));
    } else if (!strcmp(w[0], "pair_coeff")) { /* :391-397 */
      if (in->n_coeff_lines < ORC_MAX_COEFF_LINES) {
        memcpy(in->coeff_words[in->n_coeff_lines], w, sizeof(w));
        in->coeff_nwords[in->n_coeff_lines] = nw;
        in->n_coeff_lines++;
      }
    } else if (!strcmp(w[0], "velocity")) { /* :398-410 */
      in->temperature_target = atof(w[3]); in->temperature_seed = atoi(w[4]);
    } else if (!strcmp(w[0], "neighbor")) { in->neighbor_skin = atof(w[1]); }
    else if (!strcmp(w[0], "neigh_modify")) { /* :415-421 */
      for (int i = 1; i < ORC_MAX_WORDS - 1; i++) if (!strcmp(w[i], "every")) in->comm_exchange_rate = atoi(w[i + 1]);
    } else if (!strcmp(w[0], "run")) { in->nsteps = atoi(w[1]); }
    else if (!strcmp(w[0], "thermo")) { in->thermo_rate = atoi(w[1]); }
    else if (!strcmp(w[0], "timestep")) { s->dt = atof(w[1]); in->timestepflag = 1; }
    else if (!strcmp(w[0], "newton")) { if (!strcmp(w[1], "on")) in->comm_newton = 1; else if (!strcmp(w[1], "off")) in->comm_newton = 0; }
    /* atom_style / create_atoms / fix ... nve: accepted, no state */
  }
  fclose(fp);
  return 0;
}

/* LAMMPS_RandomVelocityGeom, src/input.h:66-133 */
#define ORC_IA 16807
#define ORC_IM 2147483647
#define ORC_AM (1.0 / ORC_IM)
#define ORC_IQ 127773
#define ORC_IR 2836
double orc_random_uniform_state(int *seed) { /* input.h:78-85 */
  int k = *seed / ORC_IQ;
  *seed = ORC_IA * (*seed - k * ORC_IQ) - ORC_IR * k;
  if (*seed < 0) *seed += ORC_IM;
  return ORC_AM * *seed;
}
void orc_random_reset(int *seed, int ibase, const double *coord) { /* input.h:100-132 */
  const char *str = (const char *)&ibase; /* plain char: signed on x86-64, as the reference */
  int n = sizeof(int);
  unsigned int hash = 0;
  for (int i = 0; i < n; i++) { hash += str[i]; hash += (hash << 10); hash ^= (hash >> 6); }
  str = (const char *)coord;
  n = 3 * sizeof(double);
  for (int i = 0; i < n; i++) { hash += str[i]; hash += (hash << 10); hash ^= (hash >> 6); }
  hash += (hash << 3); hash ^= (hash >> 11); hash += (hash << 15);
  *seed = hash & 0x7ffffff; /* 27-bit mask, input.h:126 */
  if (!*seed) *seed = 1;
  for (int i = 0; i < 5; i++) orc_random_uniform_state(seed);
}

/* Comm::create_domain_decomposition (src/comm.cpp:56-63): one brick = whole box */
static void single_rank_decomposition(orc_system *s) {
  s->sub_domain_lo_x = s->sub_domain_lo_y = s->sub_domain_lo_z = 0.0;
  s->sub_domain_x = s->sub_domain_hi_x = s->domain_x;
  s->sub_domain_y = s->sub_domain_hi_y = s->domain_y;
  s->sub_domain_z = s->sub_domain_hi_z = s->domain_z;
}

/* positions, types and ids: src/input.cpp:460-589 (sc) / :592-725 (fcc).
 * sub_domain_* must already be set when s->sub_domain_x != 0 (multi-rank callers);
 * otherwise the single-rank decomposition is applied. */
void orc_lattice_positions(const orc_input *in, orc_system *s) {
  s->domain_x = in->lattice_constant * in->lattice_nx;
  s->domain_y = in->lattice_constant * in->lattice_ny;
  s->domain_z = in->lattice_constant * in->lattice_nz;
  if (s->sub_domain_x == 0.0) single_rank_decomposition(s);

  int ix_start = (int)(s->sub_domain_lo_x / s->domain_x * in->lattice_nx - 0.5);
  int iy_start = (int)(s->sub_domain_lo_y / s->domain_y * in->lattice_ny - 0.5);
  int iz_start = (int)(s->sub_domain_lo_z / s->domain_z * in->lattice_nz - 0.5);
  int ix_end = (int)(s->sub_domain_hi_x / s->domain_x * in->lattice_nx + 0.5);
  int iy_end = (int)(s->sub_domain_hi_y / s->domain_y * in->lattice_ny + 0.5);
  int iz_end = (int)(s->sub_domain_hi_z / s->domain_z * in->lattice_nz + 0.5);
  const double a = in->lattice_constant;
  double basis[4][3] = {{0.0, 0.0, 0.0}, {0.5, 0.5, 0.0}, {0.5, 0.0, 0.5}, {0.0, 0.5, 0.5}};
  int nbasis = 4;
  if (in->lattice_style == ORC_LATTICE_SC) nbasis = 1;
  for (int k = 0; k < 4; k++) { basis[k][0] += in->lattice_offset_x; basis[k][1] += in->lattice_offset_y; basis[k][2] += in->lattice_offset_z; }

  for (int pass = 0; pass < 2; pass++) {
    int n = 0;
    for (int iz = iz_start; iz <= iz_end; iz++)
      for (int iy = iy_start; iy <= iy_end; iy++)
        for (int ix = ix_start; ix <= ix_end; ix++)
          for (int k = 0; k < nbasis; k++) {
            double xtmp, ytmp, ztmp;
            if (in->lattice_style == ORC_LATTICE_SC) { /* :486-491: a*(i+offset), int + double */
              xtmp = a * (ix + in->lattice_offset_x);
              ytmp = a * (iy + in->lattice_offset_y);
              ztmp = a * (iz + in->lattice_offset_z);
            } else { /* :626-628 */
              xtmp = a * (1.0 * ix + basis[k][0]);
              ytmp = a * (1.0 * iy + basis[k][1]);
              ztmp = a * (1.0 * iz + basis[k][2]);
            }
            if (xtmp >= s->sub_domain_lo_x && ytmp >= s->sub_domain_lo_y && ztmp >= s->sub_domain_lo_z &&
                xtmp < s->sub_domain_hi_x && ytmp < s->sub_domain_hi_y && ztmp < s->sub_domain_hi_z) {
              if (pass == 1) {
                s->x[3 * n + 0] = xtmp; s->x[3 * n + 1] = ytmp; s->x[3 * n + 2] = ztmp;
                s->type[n] = rand() % s->ntypes; /* :567,706 */
                s->id[n] = n + 1;
              }
              n++;
            }
          }
    if (pass == 0) {
      s->N_local = n; s->N = n;
      /* the reference counts twice and grows to 2n (:497-534, :638-676); capacity only */
      orc_system_grow(s, 2 * n);
    }
  }
}

/* raw velocities before momentum removal: src/input.cpp:731-757; mom4 = {px,py,pz,mass} */
void orc_lattice_velocities_raw(const orc_input *in, orc_system *s, double *mom4) {
  double total_mass = 0.0, px = 0.0, py = 0.0, pz = 0.0;
  for (int i = 0; i < s->N_local; i++) {
    int seed = 0;
    double x[3] = {s->x[3 * i], s->x[3 * i + 1], s->x[3 * i + 2]};
    orc_random_reset(&seed, in->temperature_seed, x);
    double mass_i = s->mass[s->type[i]];
    double vx = orc_random_uniform_state(&seed) - 0.5;
    double vy = orc_random_uniform_state(&seed) - 0.5;
    double vz = orc_random_uniform_state(&seed) - 0.5;
    s->v[3 * i + 0] = vx / sqrt(mass_i);
    s->v[3 * i + 1] = vy / sqrt(mass_i);
    s->v[3 * i + 2] = vz / sqrt(mass_i);
    s->q[i] = 0.0;
    total_mass += mass_i;
    px += mass_i * s->v[3 * i + 0];
    py += mass_i * s->v[3 * i + 1];
    pz += mass_i * s->v[3 * i + 2];
  }
  mom4[0] = px; mom4[1] = py; mom4[2] = pz; mom4[3] = total_mass;
}

/* src/input.cpp:460-792, single rank */
void orc_create_lattice(const orc_input *in, orc_system *s) {
  orc_lattice_positions(in, s);
  double m[4];
  orc_lattice_velocities_raw(in, s, m);
  double svx = m[0] / m[3], svy = m[1] / m[3], svz = m[2] / m[3]; /* :763-765 */
  for (int i = 0; i < s->N_local; i++) { s->v[3 * i] -= svx; s->v[3 * i + 1] -= svy; s->v[3 * i + 2] -= svz; }
  double T = orc_temperature(s);                      /* :774-775 */
  double scale = sqrt(in->temperature_target / T);    /* :777 */
  for (int i = 0; i < s->N_local; i++) { s->v[3 * i] *= scale; s->v[3 * i + 1] *= scale; s->v[3 * i + 2] *= scale; }
}

/* ----------------------------------------------------------------- Binning */
void orc_binning_init(orc_binning *b) { memset(b, 0, sizeof(*b)); }
void orc_binning_destroy(orc_binning *b) { free(b->bincount); free(b->binoffsets); free(b->permute_vector); memset(b, 0, sizeof(*b)); }

/* src/binning_types/binning_kksort.cpp:71-140 with Kokkos::BinSort/BinOp3D
 * (kokkos/kokkos >=3.0 algorithms/src/Kokkos_Sort.hpp: BinOp3D::bin,
 * BinSort::create_permute_vector, BinSort::sort) restated from the published algorithm. */
void orc_create_binning(orc_binning *b, orc_system *s, double dx_in, double dy_in, double dz_in,
                        int halo_depth, int do_local, int do_ghost, int sort) {
  if (!(do_local || do_ghost)) return;
  b->nhalo = halo_depth;
  int begin = do_local ? 0 : s->N_local;
  int end = do_ghost ? s->N_local + s->N_ghost : s->N_local;
  b->range_begin = begin; b->range_end = end;

  b->nbinx = (int)(s->sub_domain_x / dx_in);
  b->nbiny = (int)(s->sub_domain_y / dy_in);
  b->nbinz = (int)(s->sub_domain_z / dz_in);
  if (b->nbinx == 0) b->nbinx = 1;
  if (b->nbiny == 0) b->nbiny = 1;
  if (b->nbinz == 0) b->nbinz = 1;
  double dx = s->sub_domain_x / b->nbinx;
  double dy = s->sub_domain_y / b->nbiny;
  double dz = s->sub_domain_z / b->nbinz;
  b->nbinx += 2 * halo_depth; b->nbiny += 2 * halo_depth; b->nbinz += 2 * halo_depth;
  double eps = dx / 1000; /* x's eps is reused for y and z, :91 */
  b->minx = -dx * halo_depth - eps + s->sub_domain_lo_x;
  b->maxx = dx * halo_depth + eps + s->sub_domain_hi_x;
  b->miny = -dy * halo_depth - eps + s->sub_domain_lo_y;
  b->maxy = dy * halo_depth + eps + s->sub_domain_hi_y;
  b->minz = -dz * halo_depth - eps + s->sub_domain_lo_z;
  b->maxz = dz * halo_depth + eps + s->sub_domain_hi_z;

  /* BinOp3D ctor */
  int max_bins[3] = {b->nbinx, b->nbiny, b->nbinz};
  double mn[3] = {b->minx, b->miny, b->minz}, mx[3] = {b->maxx, b->maxy, b->maxz}, mul[3];
  for (int d = 0; d < 3; d++) mul[d] = (double)max_bins[d] / (mx[d] - mn[d]);
  int nbins = max_bins[0] * max_bins[1] * max_bins[2];
  int nrange = end - begin;
  if (nbins > b->nbins_cap) {
    b->bincount = (int *)realloc(b->bincount, sizeof(int) * (size_t)nbins);
    b->binoffsets = (int *)realloc(b->binoffsets, sizeof(int) * (size_t)nbins);
    b->nbins_cap = nbins;
  }
  if (nrange > b->perm_cap) { b->permute_vector = (int *)realloc(b->permute_vector, sizeof(int) * (size_t)nrange); b->perm_cap = nrange; }
  const double *x = s->x + 3 * (size_t)begin; /* subview(system->x, range, ALL) */

  /* BinSort::create_permute_vector: histogram, exclusive scan, slot claim in arrival
   * (= ascending index when serial) order */
  memset(b->bincount, 0, sizeof(int) * (size_t)nbins);
#define ORC_BIN(i) ((((int)(mul[0] * (x[3 * (i)] - mn[0]))) * max_bins[1] + (int)(mul[1] * (x[3 * (i) + 1] - mn[1]))) * max_bins[2] + (int)(mul[2] * (x[3 * (i) + 2] - mn[2])))
  for (int i = 0; i < nrange; i++) b->bincount[ORC_BIN(i)]++;
  int acc = 0;
  for (int c = 0; c < nbins; c++) { b->binoffsets[c] = acc; acc += b->bincount[c]; }
  memset(b->bincount, 0, sizeof(int) * (size_t)nbins);
  for (int i = 0; i < nrange; i++) {
    int c = ORC_BIN(i);
    b->permute_vector[b->binoffsets[c] + b->bincount[c]++] = i;
  }
#undef ORC_BIN
  /* AssignOffsets (:45-68,121-125) reshapes 1-D -> 3-D with the same flattening: no-op here */

  if (sort) { /* BinSort::sort on x,v,f,type,id,q (:126-138) */
    size_t n = (size_t)nrange;
    double *tmp3 = (double *)malloc(n * 24);
    double *arr3[3] = {s->x, s->v, s->f};
    for (int a = 0; a < 3; a++) {
      double *p = arr3[a] + 3 * (size_t)begin;
      for (size_t i = 0; i < n; i++) { const double *src = p + 3 * (size_t)b->permute_vector[i]; tmp3[3 * i] = src[0]; tmp3[3 * i + 1] = src[1]; tmp3[3 * i + 2] = src[2]; }
      memcpy(p, tmp3, n * 24);
    }
    int *tmpi = (int *)tmp3;
    int *arri[2] = {s->type, s->id};
    for (int a = 0; a < 2; a++) {
      int *p = arri[a] + begin;
      for (size_t i = 0; i < n; i++) tmpi[i] = p[b->permute_vector[i]];
      memcpy(p, tmpi, n * 4);
    }
    double *pq = s->q + begin;
    for (size_t i = 0; i < n; i++) tmp3[i] = pq[b->permute_vector[i]];
    memcpy(pq, tmp3, n * 8);
    free(tmp3);
  }
}

/* -------------------------------------------------------------- CommSerial */
void orc_comm_serial_init(orc_comm_serial *c, double comm_depth) { memset(c, 0, sizeof(*c)); c->comm_depth = comm_depth; }
void orc_comm_serial_destroy(orc_comm_serial *c) { for (int p = 0; p < 6; p++) free(c->pack_indicies[p]); memset(c, 0, sizeof(*c)); }

/* TagExchangeSelf, src/comm_types/comm_serial.h:94-107: both tests use the OLD coordinate */
void orc_comm_exchange(orc_comm_serial *c, orc_system *s) {
  (void)c;
  const double L[3] = {s->domain_x, s->domain_y, s->domain_z};
#ifdef ORC_OPENMP
#pragma omp parallel for
#endif
  for (int i = 0; i < s->N_local; i++)
    for (int d = 0; d < 3; d++) {
      const double x = s->x[3 * i + d];
      if (x > L[d]) s->x[3 * i + d] -= L[d];
      if (x < 0) s->x[3 * i + d] += L[d];
    }
}

/* System::get_particle/set_particle (src/system.h:101-118): x,v,q,id,type -- not f */
static void copy_particle_shift(orc_system *s, int dest, int src, int dim, double shift) {
  for (int d = 0; d < 3; d++) { s->x[3 * dest + d] = s->x[3 * src + d]; s->v[3 * dest + d] = s->v[3 * src + d]; }
  s->x[3 * dest + dim] = s->x[3 * src + dim] + shift; /* p.x -= domain  <=>  x + (-domain) exactly */
  s->q[dest] = s->q[src]; s->id[dest] = s->id[src]; s->type[dest] = s->type[src];
}

/* src/comm_types/comm_serial.cpp:56-97 + TagHaloSelf comm_serial.h:110-181.  The
 * grow-and-redo loop only changes capacities; with ascending-index slot claim the
 * result equals a single pass into sufficiently large arrays. */
void orc_comm_exchange_halo(orc_comm_serial *c, orc_system *s) {
  const int N_local = s->N_local;
  int N_ghost = 0;
  const double L[3] = {s->domain_x, s->domain_y, s->domain_z};
  const double lo[3] = {s->sub_domain_lo_x, s->sub_domain_lo_y, s->sub_domain_lo_z};
  const double hi[3] = {s->sub_domain_hi_x, s->sub_domain_hi_y, s->sub_domain_hi_z};
  for (int phase = 0; phase < 6; phase++) {
    const int dim = phase / 2;
    const int nparticles = N_local + N_ghost - ((phase % 2 == 1) ? c->num_ghost[phase - 1] : 0);
    int count = 0;
    for (int i = 0; i < nparticles; i++) {
      const double xi = s->x[3 * i + dim];
      const int take = (phase % 2 == 0) ? (xi >= hi[dim] - c->comm_depth) : (xi <= lo[dim] + c->comm_depth);
      count += take;
    }
    if (N_local + N_ghost + count > s->N_max) orc_system_grow(s, N_local + N_ghost + count);
    if (count > c->pack_cap[phase]) {
      c->pack_cap[phase] = (int)(count * 1.1) + 1;
      c->pack_indicies[phase] = (int *)realloc(c->pack_indicies[phase], sizeof(int) * (size_t)c->pack_cap[phase]);
    }
    int slot = 0;
    for (int i = 0; i < nparticles; i++) {
      const double xi = s->x[3 * i + dim];
      const int take = (phase % 2 == 0) ? (xi >= hi[dim] - c->comm_depth) : (xi <= lo[dim] + c->comm_depth);
      if (take) {
        c->pack_indicies[phase][slot] = i;
        copy_particle_shift(s, N_local + N_ghost + slot, i, dim, (phase % 2 == 0) ? -L[dim] : L[dim]);
        slot++;
      }
    }
    c->num_ghost[phase] = count;
    N_ghost += count;
  }
  s->N_ghost = N_ghost;
}

/* src/comm_types/comm_serial.cpp:99-110 + TagHaloUpdateSelf comm_serial.h:183-197
 * (re-copies the WHOLE particle, not just x) */
void orc_comm_update_halo(orc_comm_serial *c, orc_system *s) {
  int N_ghost = 0;
  const double L[3] = {s->domain_x, s->domain_y, s->domain_z};
  for (int phase = 0; phase < 6; phase++) {
    const int dim = phase / 2;
    const double shift = (phase % 2 == 0) ? -L[dim] : L[dim];
#ifdef ORC_OPENMP
#pragma omp parallel for
#endif
    for (int i = 0; i < c->num_ghost[phase]; i++)
      copy_particle_shift(s, s->N_local + N_ghost + i, c->pack_indicies[phase][i], dim, shift);
    N_ghost += c->num_ghost[phase];
  }
}

/* src/comm_types/comm_serial.cpp:112-127 + TagHaloForceSelf comm_serial.h:199-213 */
void orc_comm_update_force(orc_comm_serial *c, orc_system *s) {
  int ghost_offsets[6];
  ghost_offsets[0] = s->N_local;
  for (int p = 1; p < 6; p++) ghost_offsets[p] = ghost_offsets[p - 1] + c->num_ghost[p - 1];
  for (int phase = 5; phase >= 0; phase--)
    for (int ii = 0; ii < c->num_ghost[phase]; ii++) {
      const int i = c->pack_indicies[phase][ii];
      for (int d = 0; d < 3; d++) s->f[3 * i + d] += s->f[3 * (ghost_offsets[phase] + ii) + d];
    }
}

/* ---------------------------------------------------------------- Neighbor */
void orc_neighbor_init(orc_neighbor *n, int kind, double neigh_cut) {
  memset(n, 0, sizeof(*n));
  n->kind = kind; n->neigh_cut = neigh_cut;
  n->maxneighs = 16; /* neighbor_2d.h:102-104 */
}
void orc_neighbor_destroy(orc_neighbor *n) { free(n->row_map); free(n->entries); free(n->num_neighs); free(n->neighs2d); memset(n, 0, sizeof(*n)); }

/* One traversal serves count and fill for both list kinds.  Order = the reference's CPU
 * order: interior bins by league rank, atoms of the bin in permute order, 27 stencil bins
 * bx-1..bx+1 / by / bz, atoms of each in permute order.
 * Predicates: full neighbor_csr.h:199-207 (2D: neighbor_2d.h:197-205);
 *             half neighbor_csr.h:286-301 (2D: neighbor_2d.h:250-268). */
static inline int neigh_row(const orc_neighbor *n, const orc_system *s, const orc_binning *b, int half,
                            int bx, int by, int bz, int i, int *out, int out_cap) {
  const double *x = s->x;
  const int N_local = s->N_local;
  const double x_i = x[3 * i], y_i = x[3 * i + 1], z_i = x[3 * i + 2];
  const double cutsq = n->neigh_cut * n->neigh_cut;
  int count = 0;
  for (int bx_j = bx - 1; bx_j < bx + 2; bx_j++)
    for (int by_j = by - 1; by_j < by + 2; by_j++)
      for (int bz_j = bz - 1; bz_j < bz + 2; bz_j++) {
        const int c = (bx_j * b->nbiny + by_j) * b->nbinz + bz_j;
        const int j_offset = b->binoffsets[c], cnt = b->bincount[c];
        for (int bj = 0; bj < cnt; bj++) {
          const int j = b->permute_vector[j_offset + bj];
          const double x_j = x[3 * j], y_j = x[3 * j + 1], z_j = x[3 * j + 2];
          if (half) {
            if (((j == i) || (j < N_local || n->comm_newton)) &&
                !((x_j > x_i) || ((x_j == x_i) && ((y_j > y_i) || ((y_j == y_i) && (z_j > z_i))))))
              continue;
          }
          const double dx = x_i - x_j, dy = y_i - y_j, dz = z_i - z_j;
          const double rsq = dx * dx + dy * dy + dz * dz;
          if (half ? (rsq <= cutsq) : ((rsq <= cutsq) && (i != j))) {
            if (out && count < out_cap) out[count] = j;
            count++;
          }
        }
      }
  return count;
}

void orc_create_neigh_list(orc_neighbor *n, const orc_system *s, const orc_binning *b, int half) {
  const int N_local = s->N_local;
  n->N_local = N_local;
  const int nhalo = b->nhalo;
  const int nbinx = b->nbinx - 2 * nhalo, nbiny = b->nbiny - 2 * nhalo, nbinz = b->nbinz - 2 * nhalo;
  const int nbins = nbinx * nbiny * nbinz;

  if (n->kind == ORC_NEIGH_2D) { /* neighbor_2d.h:280-331 */
    if (n->rows_cap < N_local + 1) {
      n->rows_cap = N_local + 1;
      n->num_neighs = (int *)realloc(n->num_neighs, sizeof(int) * (size_t)n->rows_cap);
    }
    n->n_fill_passes = 0;
    int resize;
    do {
      /* the reference reallocates when extent(0) < N_local+1 or extent(1) < maxneighs and
       * strides rows by extent(1); we keep extent(1) == maxneighs (row stride is not observable
       * through the list accessor, neighbor_2d.h:112-114) */
      if (n->rows2d_cap < N_local + 1 || n->cols2d_cap != n->maxneighs) {
        if (n->rows2d_cap < N_local + 1) n->rows2d_cap = N_local + 1;
        n->cols2d_cap = n->maxneighs;
        n->neighs2d = (int *)realloc(n->neighs2d, sizeof(int) * (size_t)n->rows2d_cap * (size_t)n->cols2d_cap);
      }
      memset(n->num_neighs, 0, sizeof(int) * (size_t)(N_local + 1));
      resize = 0;
      int new_maxneighs = 0;
      for (int lr = 0; lr < nbins; lr++) {
        const int bx = lr / (nbiny * nbinz) + nhalo, by = (lr / nbinz) % nbiny + nhalo, bz = lr % nbinz + nhalo;
        const int c = (bx * b->nbiny + by) * b->nbinz + bz;
        for (int bi = 0; bi < b->bincount[c]; bi++) {
          const int i = b->permute_vector[b->binoffsets[c] + bi];
          if (i >= N_local) continue;
          const int cnt = neigh_row(n, s, b, half, bx, by, bz, i, n->neighs2d + (size_t)i * n->maxneighs, n->maxneighs);
          n->num_neighs[i] = cnt;
          if (cnt > n->maxneighs) { resize = 1; new_maxneighs = cnt; } /* last writer wins, :210-216 */
        }
      }
      n->n_fill_passes++;
      if (resize) n->maxneighs = (int)(new_maxneighs * 1.2);
    } while (resize);
    return;
  }

  /* CSR: neighbor_csr.h:370-435 */
  if (n->rows_cap < N_local + 1) {
    n->rows_cap = N_local + 1;
    n->row_map = (int *)realloc(n->row_map, sizeof(int) * (size_t)n->rows_cap);
  }
  int *counts = (int *)calloc((size_t)N_local + 1, sizeof(int));
#ifdef ORC_OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
  for (int lr = 0; lr < nbins; lr++) { /* count_neighbors_{half,full} */
    const int bx = lr / (nbiny * nbinz) + nhalo, by = (lr / nbinz) % nbiny + nhalo, bz = lr % nbinz + nhalo;
    const int c = (bx * b->nbiny + by) * b->nbinz + bz;
    for (int bi = 0; bi < b->bincount[c]; bi++) {
      const int i = b->permute_vector[b->binoffsets[c] + bi];
      if (i >= N_local) continue;
      counts[i] = neigh_row(n, s, b, half, bx, by, bz, i, NULL, 0);
    }
  }
  int acc = 0; /* create_offsets :359-368 */
  for (int i = 0; i < N_local; i++) { n->row_map[i] = acc; acc += counts[i]; }
  n->row_map[N_local] = acc;
  n->total = acc;
  if (n->entries_cap < acc) { n->entries_cap = acc; n->entries = (int *)realloc(n->entries, sizeof(int) * (size_t)acc); }
#ifdef ORC_OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
  for (int lr = 0; lr < nbins; lr++) { /* fill_neigh_list_{half,full} */
    const int bx = lr / (nbiny * nbinz) + nhalo, by = (lr / nbinz) % nbiny + nhalo, bz = lr % nbinz + nhalo;
    const int c = (bx * b->nbiny + by) * b->nbinz + bz;
    for (int bi = 0; bi < b->bincount[c]; bi++) {
      const int i = b->permute_vector[b->binoffsets[c] + bi];
      if (i >= N_local) continue;
      neigh_row(n, s, b, half, bx, by, bz, i, n->entries + n->row_map[i], counts[i]);
    }
  }
  free(counts);
}

/* ----------------------------------------------------------------- ForceLJ */
void orc_force_lj_init(orc_force_lj *f, int ntypes, int half_neigh) {
  memset(f, 0, sizeof(*f));
  f->ntypes = ntypes; f->half_neigh = half_neigh;
  size_t n = (size_t)ntypes * ntypes;
  f->lj1 = (double *)calloc(n, 8); f->lj2 = (double *)calloc(n, 8); f->cutsq = (double *)calloc(n, 8);
  f->intensity = (double *)calloc(n, 8);
}
void orc_force_lj_destroy(orc_force_lj *f) { free(f->lj1); free(f->lj2); free(f->cutsq); free(f->intensity); memset(f, 0, sizeof(*f)); }

/* src/force_types/force_lj_neigh_impl.h:57-98.  args = the pair_coeff line's words:
 * args[1],args[2] types (1-based), args[3] eps, args[4] sigma, args[5] cut. */
void orc_force_lj_init_coeff(orc_force_lj *f, int nargs, char args[][ORC_WORD]) {
  (void)nargs;
  int t1 = atoi(args[1]) - 1, t2 = atoi(args[2]) - 1;
  double eps = atof(args[3]), sigma = atof(args[4]), cut = atof(args[5]);
  if (f->ntypes <= ORC_MAX_TYPES_STACKPARAMS) { /* stackparams: every line overwrites ALL pairs, :66-74 */
    for (int i = 0; i < f->ntypes; i++)
      for (int j = 0; j < f->ntypes; j++) {
        f->lj1[i * f->ntypes + j] = 48.0 * eps * pow(sigma, 12.0);
        f->lj2[i * f->ntypes + j] = 24.0 * eps * pow(sigma, 6.0);
        f->cutsq[i * f->ntypes + j] = cut * cut;
      }
  } else {
    f->lj1[t1 * f->ntypes + t2] = 48.0 * eps * pow(sigma, 12.0);
    f->lj2[t1 * f->ntypes + t2] = 24.0 * eps * pow(sigma, 6.0);
    f->lj1[t2 * f->ntypes + t1] = f->lj1[t1 * f->ntypes + t2];
    f->lj2[t2 * f->ntypes + t1] = f->lj2[t1 * f->ntypes + t2];
    f->cutsq[t1 * f->ntypes + t2] = cut * cut;
    f->cutsq[t2 * f->ntypes + t1] = cut * cut;
  }
}

/* TagFullNeigh :161-206, TagHalfNeigh :208-254 */
void orc_force_lj_compute(const orc_force_lj *fl, orc_system *s, const orc_neighbor *n) {
  const double *x = s->x;
  double *f = s->f;
  const int nt = fl->ntypes;
#ifdef ORC_OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int i = 0; i < s->N_local; i++) {
    const double x_i = x[3 * i], y_i = x[3 * i + 1], z_i = x[3 * i + 2];
    const int type_i = s->type[i];
    int num_neighs;
    const int *row = orc_neigh_row(n, i, &num_neighs);
    double fxi = 0.0, fyi = 0.0, fzi = 0.0;
    for (int jj = 0; jj < num_neighs; jj++) {
      const int j = row[jj];
      const double dx = x_i - x[3 * j], dy = y_i - x[3 * j + 1], dz = z_i - x[3 * j + 2];
      const int type_j = s->type[j];
      const double rsq = dx * dx + dy * dy + dz * dz;
      const double cutsq_ij = fl->cutsq[type_i * nt + type_j];
      if (rsq < cutsq_ij) {
        const double lj1_ij = fl->lj1[type_i * nt + type_j], lj2_ij = fl->lj2[type_i * nt + type_j];
        double r2inv = 1.0 / rsq;
        double r6inv = r2inv * r2inv * r2inv;
        double fpair = (r6inv * (lj1_ij * r6inv - lj2_ij)) * r2inv;
        fxi += dx * fpair; fyi += dy * fpair; fzi += dz * fpair;
        if (fl->half_neigh) {
#ifdef ORC_OPENMP
#pragma omp atomic
          f[3 * j] -= dx * fpair;
#pragma omp atomic
          f[3 * j + 1] -= dy * fpair;
#pragma omp atomic
          f[3 * j + 2] -= dz * fpair;
#else
          f[3 * j] -= dx * fpair; f[3 * j + 1] -= dy * fpair; f[3 * j + 2] -= dz * fpair;
#endif
        }
      }
    }
#ifdef ORC_OPENMP
    if (fl->half_neigh) {
#pragma omp atomic
      f[3 * i] += fxi;
#pragma omp atomic
      f[3 * i + 1] += fyi;
#pragma omp atomic
      f[3 * i + 2] += fzi;
    } else
#endif
    { f[3 * i] += fxi; f[3 * i + 1] += fyi; f[3 * i + 2] += fzi; }
  }
}

/* ForceLJIDialNeigh::init_coeff, src/force_types/force_lj_idial_neigh_impl.h:50-88: args[6] = nrepeat; unlike ForceLJNeigh there
 * is no stack-parameter path, a line sets its own (t1,t2)/(t2,t1) entries only. */
void orc_force_lj_idial_init_coeff(orc_force_lj *f, int nargs, char args[][ORC_WORD]) {
  (void)nargs;
  const int t1 = atoi(args[1]) - 1, t2 = atoi(args[2]) - 1, nt = f->ntypes;
  const double eps = atof(args[3]), sigma = atof(args[4]), cut = atof(args[5]);
  const int nrepeat = atoi(args[6]);
  f->idial = 1;
  f->lj1[t1 * nt + t2] = 48.0 * eps * pow(sigma, 12.0);
  f->lj2[t1 * nt + t2] = 24.0 * eps * pow(sigma, 6.0);
  f->lj1[t2 * nt + t1] = f->lj1[t1 * nt + t2];
  f->lj2[t2 * nt + t1] = f->lj2[t1 * nt + t2];
  f->cutsq[t1 * nt + t2] = cut * cut;
  f->cutsq[t2 * nt + t1] = cut * cut;
  f->intensity[t1 * nt + t2] = nrepeat;
  f->intensity[t2 * nt + t1] = nrepeat;
}

/* TagFullNeigh :113-163, TagHalfNeigh :165-213: the pair force is accumulated `intensity` times, each term divided by
 * `intensity` (the loop bound compares the int counter with the double-valued view entry); half lists subtract from j only
 * when j is an owned atom (:203-207), whatever `newton` says. */
void orc_force_lj_idial_compute(const orc_force_lj *fl, orc_system *s, const orc_neighbor *n) {
  const double *x = s->x;
  double *f = s->f;
  const int nt = fl->ntypes;
  for (int i = 0; i < s->N_local; i++) {
    const double x_i = x[3 * i], y_i = x[3 * i + 1], z_i = x[3 * i + 2];
    const int type_i = s->type[i];
    int num_neighs;
    const int *row = orc_neigh_row(n, i, &num_neighs);
    double fxi = 0.0, fyi = 0.0, fzi = 0.0;
    for (int jj = 0; jj < num_neighs; jj++) {
      const int j = row[jj];
      const double dx = x_i - x[3 * j], dy = y_i - x[3 * j + 1], dz = z_i - x[3 * j + 2];
      const int type_j = s->type[j];
      const double rsq = dx * dx + dy * dy + dz * dz;
      if (rsq < fl->cutsq[type_i * nt + type_j]) {
        const double lj1_ij = fl->lj1[type_i * nt + type_j], lj2_ij = fl->lj2[type_i * nt + type_j];
        const double inten = fl->intensity[type_i * nt + type_j];
        double fpair = 0;
        for (int repeat = 0; repeat < inten; repeat++) {
          double r2inv = 1.0 / rsq;
          double r6inv = r2inv * r2inv * r2inv;
          fpair += (r6inv * (lj1_ij * r6inv - lj2_ij)) * r2inv / inten;
        }
        fxi += dx * fpair; fyi += dy * fpair; fzi += dz * fpair;
        if (fl->half_neigh && j < s->N_local) { f[3 * j] -= dx * fpair; f[3 * j + 1] -= dy * fpair; f[3 * j + 2] -= dz * fpair; }
      }
    }
    f[3 * i] += fxi; f[3 * i + 1] += fyi; f[3 * i + 2] += fzi;
  }
}

/* TagFullNeighPE :256-296, TagHalfNeighPE :298-343 (cutoff-shifted, shift_flag = true) */
double orc_force_lj_energy(const orc_force_lj *fl, const orc_system *s, const orc_neighbor *n) {
  const double *x = s->x;
  const int nt = fl->ntypes;
  double PE = 0.0;
#ifdef ORC_OPENMP
#pragma omp parallel for reduction(+ : PE) schedule(static)
#endif
  for (int i = 0; i < s->N_local; i++) {
    const double x_i = x[3 * i], y_i = x[3 * i + 1], z_i = x[3 * i + 2];
    const int type_i = s->type[i];
    int num_neighs;
    const int *row = orc_neigh_row(n, i, &num_neighs);
    for (int jj = 0; jj < num_neighs; jj++) {
      const int j = row[jj];
      const double dx = x_i - x[3 * j], dy = y_i - x[3 * j + 1], dz = z_i - x[3 * j + 2];
      const int type_j = s->type[j];
      const double rsq = dx * dx + dy * dy + dz * dz;
      const double cutsq_ij = fl->cutsq[type_i * nt + type_j];
      if (rsq < cutsq_ij) {
        const double lj1_ij = fl->lj1[type_i * nt + type_j], lj2_ij = fl->lj2[type_i * nt + type_j];
        double r2inv = 1.0 / rsq;
        double r6inv = r2inv * r2inv * r2inv;
        double r2invc = 1.0 / cutsq_ij;
        double r6invc = r2invc * r2invc * r2invc;
        if (fl->half_neigh) {
          double fac = (j < s->N_local) ? 1.0 : 0.5;
          PE += fac * r6inv * (0.5 * lj1_ij * r6inv - lj2_ij) / 6.0;
          PE -= fac * r6invc * (0.5 * lj1_ij * r6invc - lj2_ij) / 6.0;
        } else {
          PE += 0.5 * r6inv * (0.5 * lj1_ij * r6inv - lj2_ij) / 6.0;
          PE -= 0.5 * r6invc * (0.5 * lj1_ij * r6invc - lj2_ij) / 6.0;
        }
      }
    }
  }
  return PE;
}

/* -------------------------------------------------------------- Integrator */
/* src/integrator_nve.cpp:41-44 (dtv, dtf), :66-74 */
void orc_initial_integrate(orc_system *s) {
  const double dtv = s->dt, dtf = 0.5 * s->dt / s->mvv2e;
#ifdef ORC_OPENMP
#pragma omp parallel for
#endif
  for (int i = 0; i < s->N_local; i++) {
    const double dtfm = dtf / s->mass[s->type[i]];
    for (int d = 0; d < 3; d++) s->v[3 * i + d] += dtfm * s->f[3 * i + d];
    for (int d = 0; d < 3; d++) s->x[3 * i + d] += dtv * s->v[3 * i + d];
  }
}
/* :105-112 */
void orc_final_integrate(orc_system *s) {
  const double dtf = 0.5 * s->dt / s->mvv2e;
#ifdef ORC_OPENMP
#pragma omp parallel for
#endif
  for (int i = 0; i < s->N_local; i++) {
    const double dtfm = dtf / s->mass[s->type[i]];
    for (int d = 0; d < 3; d++) s->v[3 * i + d] += dtfm * s->f[3 * i + d];
  }
}

/* ------------------------------------------------------------------ Thermo */
static double sum_mv2(const orc_system *s) { /* property_temperature.h:55-57 */
  double T = 0.0;
#ifdef ORC_OPENMP
#pragma omp parallel for reduction(+ : T)
#endif
  for (int i = 0; i < s->N_local; i++)
    T += (s->v[3 * i] * s->v[3 * i] + s->v[3 * i + 1] * s->v[3 * i + 1] + s->v[3 * i + 2] * s->v[3 * i + 2]) * s->mass[s->type[i]];
  return T;
}
double orc_temperature(const orc_system *s) { /* property_temperature.cpp:43-62 */
  int dof = 3 * s->N - 3;
  double factor = s->mvv2e / (1.0 * dof * s->boltz);
  return sum_mv2(s) * factor;
}
double orc_kine(const orc_system *s) { /* property_kine.cpp:43-61 */
  return sum_mv2(s) * (0.5 * s->mvv2e);
}

/* ------------------------------------------------------------------ Driver */
/* src/examinimd.cpp:60-146 */
int orc_md_init(orc_md *md, const char *deck, int neighbor_type, int force_iteration_type, const char *coeff_dir) {
  memset(md, 0, sizeof(*md));
  orc_system_init(&md->sys);
  orc_input_defaults(&md->in);
  if (neighbor_type >= 0) md->in.neighbor_type = neighbor_type;
  if (force_iteration_type >= 0) md->in.force_iteration_type = force_iteration_type;
  if (orc_input_read_deck(&md->in, &md->sys, deck)) return -1;
  const int half = md->in.force_iteration_type == ORC_ITER_NEIGH_HALF;
  md->neigh_cutoff = md->in.force_cutoff + md->in.neighbor_skin;
  orc_binning_init(&md->bin);
  if (md->in.force_type == ORC_FORCE_LJ) {
    orc_force_lj_init(&md->lj, md->sys.ntypes, half);
    for (int l = 0; l < md->in.n_coeff_lines; l++) orc_force_lj_init_coeff(&md->lj, md->in.coeff_nwords[l], md->in.coeff_words[l]);
    md->lj.comm_newton = md->in.comm_newton;
  } else if (md->in.force_type == ORC_FORCE_LJ_IDIAL) {
    orc_force_lj_init(&md->lj, md->sys.ntypes, half);
    for (int l = 0; l < md->in.n_coeff_lines; l++) orc_force_lj_idial_init_coeff(&md->lj, md->in.coeff_nwords[l], md->in.coeff_words[l]);
    md->lj.comm_newton = md->in.comm_newton;
  } else if (md->in.force_type == ORC_FORCE_SNAP) {
    md->snap = orc_force_snap_create(md->sys.ntypes);
    for (int l = 0; l < md->in.n_coeff_lines; l++)
      if (orc_force_snap_init_coeff(md->snap, md->in.coeff_nwords[l], md->in.coeff_words[l], coeff_dir)) return -2;
  } else
    return -3;
  int kind = md->in.neighbor_type == ORC_NEIGH_2D ? ORC_NEIGH_2D : ORC_NEIGH_CSR; /* CSR_MAPCONSTR yields the same list */
  orc_neighbor_init(&md->neigh, kind, md->neigh_cutoff);
  md->neigh.comm_newton = md->in.comm_newton;
  orc_comm_serial_init(&md->comm, md->neigh_cutoff);
  orc_create_lattice(&md->in, &md->sys);
  orc_md_setup(md);
  return 0;
}

static void md_force(orc_md *md) {
  memset(md->sys.f, 0, sizeof(double) * 3 * (size_t)md->sys.N_max); /* deep_copy(f,0): whole allocation */
  if (md->snap) orc_force_snap_compute(md->snap, &md->sys, &md->neigh);
  else if (md->lj.idial) orc_force_lj_idial_compute(&md->lj, &md->sys, &md->neigh);
  else orc_force_lj_compute(&md->lj, &md->sys, &md->neigh);
}

/* src/examinimd.cpp:120-144 */
void orc_md_setup(orc_md *md) {
  const int half = md->in.force_iteration_type == ORC_ITER_NEIGH_HALF;
  const double c = md->neigh_cutoff;
  orc_comm_exchange(&md->comm, &md->sys);
  orc_create_binning(&md->bin, &md->sys, c, c, c, 1, 1, 0, 1);
  orc_comm_exchange_halo(&md->comm, &md->sys);
  orc_create_binning(&md->bin, &md->sys, c, c, c, 1, 1, 1, 0);
  orc_create_neigh_list(&md->neigh, &md->sys, &md->bin, half);
  md_force(md);
  if (md->in.comm_newton) orc_comm_update_force(&md->comm, &md->sys);
  md->step = 0;
}

/* one iteration of src/examinimd.cpp:192-250 */
void orc_md_step(orc_md *md) {
  const int half = md->in.force_iteration_type == ORC_ITER_NEIGH_HALF;
  const double c = md->neigh_cutoff;
  const int step = ++md->step;
  double t0 = now_s(), t1;
  orc_initial_integrate(&md->sys);
  t1 = now_s(); md->t_other += t1 - t0; t0 = t1;
  if (step % md->in.comm_exchange_rate == 0 && step > 0) {
    orc_comm_exchange(&md->comm, &md->sys);
    t1 = now_s(); md->t_comm += t1 - t0; t0 = t1;
    orc_create_binning(&md->bin, &md->sys, c, c, c, 1, 1, 0, 1);
    t1 = now_s(); md->t_other += t1 - t0; t0 = t1;
    orc_comm_exchange_halo(&md->comm, &md->sys);
    t1 = now_s(); md->t_comm += t1 - t0; t0 = t1;
    orc_create_binning(&md->bin, &md->sys, c, c, c, 1, 1, 1, 0);
    orc_create_neigh_list(&md->neigh, &md->sys, &md->bin, half);
    t1 = now_s(); md->t_neigh += t1 - t0; t0 = t1;
  } else {
    orc_comm_update_halo(&md->comm, &md->sys);
    t1 = now_s(); md->t_comm += t1 - t0; t0 = t1;
  }
  md_force(md);
  t1 = now_s(); md->t_force += t1 - t0; t0 = t1;
  if (md->in.comm_newton) { orc_comm_update_force(&md->comm, &md->sys); t1 = now_s(); md->t_comm += t1 - t0; t0 = t1; }
  orc_final_integrate(&md->sys);
  t1 = now_s(); md->t_other += t1 - t0;
}

/* src/examinimd.cpp:252-255: T, PE/N, KE/N */
void orc_md_thermo(orc_md *md, double *T, double *PE, double *KE) {
  *T = orc_temperature(&md->sys);
  double pe = (md->snap || md->lj.idial) ? 0.0 /* Force::compute_energy default, force.h:54 */ : orc_force_lj_energy(&md->lj, &md->sys, &md->neigh);
  *PE = pe / md->sys.N;
  *KE = orc_kine(&md->sys) / md->sys.N;
}

void orc_md_destroy(orc_md *md) {
  orc_system_destroy(&md->sys); orc_binning_destroy(&md->bin); orc_comm_serial_destroy(&md->comm);
  orc_neighbor_destroy(&md->neigh);
  if (md->snap) orc_force_snap_destroy(md->snap); else orc_force_lj_destroy(&md->lj);
}

/* src/examinimd.cpp:296-346: int n; id[n]; type[n]; q[n]; x[n][3]; v[n][3]; f[n][3] */
int orc_dump_binary(const orc_system *s, const char *path, int step, int rank) {
  char filename[1024];
  snprintf(filename, sizeof filename, "%s%s.%010d.%03d", path, "/output", step, rank);
  FILE *fp = fopen(filename, "wb");
  if (!fp) return -1;
  int n = s->N_local;
  fwrite(&n, sizeof(int), 1, fp);
  fwrite(s->id, sizeof(int), (size_t)n, fp);
  fwrite(s->type, sizeof(int), (size_t)n, fp);
  fwrite(s->q, sizeof(double), (size_t)n, fp);
  fwrite(s->x, sizeof(double), 3 * (size_t)n, fp);
  fwrite(s->v, sizeof(double), 3 * (size_t)n, fp);
  fwrite(s->f, sizeof(double), 3 * (size_t)n, fp);
  fclose(fp);
  return 0;
}
