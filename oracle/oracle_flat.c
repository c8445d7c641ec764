/*
 * oracle_flat.c -- flat, ctypes-friendly handle API over the oracle (TEST INFRASTRUCTURE ONLY).
 * tests/ load liboracle.so and drive the restated reference stage by stage through this file.
 */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void *orcf_create_deck(const char *deck, int neigh_type, int iter_type, int nx, int ny, int nz, const char *coeff_dir,
                       int do_setup) {
  orc_md *md = (orc_md *)calloc(1, sizeof(orc_md));
  orc_system_init(&md->sys);
  orc_input_defaults(&md->in);
  if (neigh_type >= 0) md->in.neighbor_type = neigh_type;
  if (iter_type >= 0) md->in.force_iteration_type = iter_type;
  if (orc_input_read_deck(&md->in, &md->sys, deck)) { free(md); return NULL; }
  if (nx > 0) { md->in.lattice_nx = nx; md->in.lattice_ny = ny; md->in.lattice_nz = nz; }
  const int half = md->in.force_iteration_type == ORC_ITER_NEIGH_HALF;
  md->neigh_cutoff = md->in.force_cutoff + md->in.neighbor_skin;
  orc_binning_init(&md->bin);
  if (md->in.force_type == ORC_FORCE_LJ) {
    orc_force_lj_init(&md->lj, md->sys.ntypes, half);
    for (int l = 0; l < md->in.n_coeff_lines; l++) orc_force_lj_init_coeff(&md->lj, md->in.coeff_nwords[l], md->in.coeff_words[l]);
    md->lj.comm_newton = md->in.comm_newton;
  } else if (md->in.force_type == ORC_FORCE_LJ_IDIAL) { /* ForceLJIDialNeigh, force_lj_idial_neigh.h:47-57 */
    orc_force_lj_init(&md->lj, md->sys.ntypes, half);
    for (int l = 0; l < md->in.n_coeff_lines; l++) orc_force_lj_idial_init_coeff(&md->lj, md->in.coeff_nwords[l], md->in.coeff_words[l]);
    md->lj.comm_newton = md->in.comm_newton;
  } else if (md->in.force_type == ORC_FORCE_SNAP) {
    md->snap = orc_force_snap_create(md->sys.ntypes);
    for (int l = 0; l < md->in.n_coeff_lines; l++)
      if (orc_force_snap_init_coeff(md->snap, md->in.coeff_nwords[l], md->in.coeff_words[l], coeff_dir)) { free(md); return NULL; }
  } else { free(md); return NULL; }
  orc_neighbor_init(&md->neigh, md->in.neighbor_type == ORC_NEIGH_2D ? ORC_NEIGH_2D : ORC_NEIGH_CSR, md->neigh_cutoff);
  md->neigh.comm_newton = md->in.comm_newton;
  orc_comm_serial_init(&md->comm, md->neigh_cutoff);
  orc_create_lattice(&md->in, &md->sys);
  if (do_setup) orc_md_setup(md);
  return md;
}

/* an LJ system from caller arrays (ragged/empty-bin/multi-type test cases); no setup performed */
void *orcf_create_raw(int n, const double *x, const double *v, const int *type, int ntypes, const double *mass,
                      const double *box, double force_cutoff, double skin, double dt, int newton, int neigh_type,
                      int iter_type, double eps, double sigma, int exchange_rate) {
  orc_md *md = (orc_md *)calloc(1, sizeof(orc_md));
  orc_system_init(&md->sys);
  orc_input_defaults(&md->in);
  orc_system *s = &md->sys;
  s->ntypes = ntypes;
  free(s->mass);
  s->mass = (double *)malloc(sizeof(double) * (size_t)ntypes);
  memcpy(s->mass, mass, sizeof(double) * (size_t)ntypes);
  s->boltz = 1.0; s->mvv2e = 1.0; s->dt = dt;
  s->domain_x = box[0]; s->domain_y = box[1]; s->domain_z = box[2];
  s->sub_domain_x = s->sub_domain_hi_x = box[0];
  s->sub_domain_y = s->sub_domain_hi_y = box[1];
  s->sub_domain_z = s->sub_domain_hi_z = box[2];
  orc_system_grow(s, 2 * n + 16);
  s->N = s->N_local = n;
  memcpy(s->x, x, sizeof(double) * 3 * (size_t)n);
  if (v) memcpy(s->v, v, sizeof(double) * 3 * (size_t)n);
  for (int i = 0; i < n; i++) { s->type[i] = type ? type[i] : 0; s->id[i] = i + 1; }
  md->in.force_type = ORC_FORCE_LJ;
  md->in.neighbor_type = neigh_type;
  md->in.force_iteration_type = iter_type;
  md->in.force_cutoff = force_cutoff;
  md->in.neighbor_skin = skin;
  md->in.comm_newton = newton;
  md->in.comm_exchange_rate = exchange_rate;
  const int half = iter_type == ORC_ITER_NEIGH_HALF;
  md->neigh_cutoff = force_cutoff + skin;
  orc_binning_init(&md->bin);
  orc_force_lj_init(&md->lj, ntypes, half);
  md->lj.comm_newton = newton;
  for (int k = 0; k < ntypes * ntypes; k++) {
    double e2 = eps, s6 = sigma * sigma * sigma * sigma * sigma * sigma;
    md->lj.lj1[k] = 48.0 * e2 * s6 * s6;
    md->lj.lj2[k] = 24.0 * e2 * s6;
    md->lj.cutsq[k] = force_cutoff * force_cutoff;
  }
  orc_neighbor_init(&md->neigh, neigh_type == ORC_NEIGH_2D ? ORC_NEIGH_2D : ORC_NEIGH_CSR, md->neigh_cutoff);
  md->neigh.comm_newton = newton;
  orc_comm_serial_init(&md->comm, md->neigh_cutoff);
  return md;
}

void orcf_destroy(void *h) { orc_md *md = (orc_md *)h; orc_md_destroy(md); free(md); }
void orcf_setup(void *h) { orc_md_setup((orc_md *)h); }
void orcf_step(void *h, int n) { for (int k = 0; k < n; k++) orc_md_step((orc_md *)h); }
void orcf_thermo(void *h, double *T, double *PE, double *KE) { orc_md_thermo((orc_md *)h, T, PE, KE); }

/* run one stage of the pipeline on the current state */
int orcf_stage(void *h, const char *what) {
  orc_md *md = (orc_md *)h;
  const double c = md->neigh_cutoff;
  const int half = md->in.force_iteration_type == ORC_ITER_NEIGH_HALF;
  if (!strcmp(what, "exchange")) orc_comm_exchange(&md->comm, &md->sys);
  else if (!strcmp(what, "bin_sort")) orc_create_binning(&md->bin, &md->sys, c, c, c, 1, 1, 0, 1);
  else if (!strcmp(what, "bin_nosort")) orc_create_binning(&md->bin, &md->sys, c, c, c, 1, 1, 0, 0);
  else if (!strcmp(what, "halo")) orc_comm_exchange_halo(&md->comm, &md->sys);
  else if (!strcmp(what, "bin_all")) orc_create_binning(&md->bin, &md->sys, c, c, c, 1, 1, 1, 0);
  else if (!strcmp(what, "neigh")) orc_create_neigh_list(&md->neigh, &md->sys, &md->bin, half);
  else if (!strcmp(what, "zero_f")) memset(md->sys.f, 0, sizeof(double) * 3 * (size_t)md->sys.N_max);
  else if (!strcmp(what, "force")) { if (md->snap) orc_force_snap_compute(md->snap, &md->sys, &md->neigh); else orc_force_lj_compute(&md->lj, &md->sys, &md->neigh); }
  else if (!strcmp(what, "update_halo")) orc_comm_update_halo(&md->comm, &md->sys);
  else if (!strcmp(what, "update_force")) orc_comm_update_force(&md->comm, &md->sys);
  else if (!strcmp(what, "initial_integrate")) orc_initial_integrate(&md->sys);
  else if (!strcmp(what, "final_integrate")) orc_final_integrate(&md->sys);
  else return -1;
  return 0;
}

long long orcf_get_int(void *h, const char *w) {
  orc_md *md = (orc_md *)h;
  if (!strcmp(w, "N")) return md->sys.N;
  if (!strcmp(w, "N_local")) return md->sys.N_local;
  if (!strcmp(w, "N_ghost")) return md->sys.N_ghost;
  if (!strcmp(w, "N_max")) return md->sys.N_max;
  if (!strcmp(w, "ntypes")) return md->sys.ntypes;
  if (!strcmp(w, "nbinx")) return md->bin.nbinx;
  if (!strcmp(w, "nbiny")) return md->bin.nbiny;
  if (!strcmp(w, "nbinz")) return md->bin.nbinz;
  if (!strcmp(w, "nhalo")) return md->bin.nhalo;
  if (!strcmp(w, "bin_range")) return md->bin.range_end - md->bin.range_begin;
  if (!strcmp(w, "total_neighs")) return md->neigh.total;
  if (!strcmp(w, "maxneighs")) return md->neigh.maxneighs;
  if (!strcmp(w, "fill_passes")) return md->neigh.n_fill_passes;
  if (!strcmp(w, "step")) return md->step;
  if (!strcmp(w, "nsteps")) return md->in.nsteps;
  if (!strcmp(w, "exchange_rate")) return md->in.comm_exchange_rate;
  if (!strcmp(w, "newton")) return md->in.comm_newton;
  if (!strncmp(w, "num_ghost", 9)) return md->comm.num_ghost[w[9] - '0'];
  return -1;
}

double orcf_get_double(void *h, const char *w) {
  orc_md *md = (orc_md *)h;
  const orc_system *s = &md->sys;
  if (!strcmp(w, "domain_x")) return s->domain_x;
  if (!strcmp(w, "domain_y")) return s->domain_y;
  if (!strcmp(w, "domain_z")) return s->domain_z;
  if (!strcmp(w, "dt")) return s->dt;
  if (!strcmp(w, "mvv2e")) return s->mvv2e;
  if (!strcmp(w, "boltz")) return s->boltz;
  if (!strcmp(w, "neigh_cutoff")) return md->neigh_cutoff;
  if (!strcmp(w, "force_cutoff")) return md->in.force_cutoff;
  if (!strcmp(w, "minx")) return md->bin.minx;
  if (!strcmp(w, "maxx")) return md->bin.maxx;
  if (!strcmp(w, "miny")) return md->bin.miny;
  if (!strcmp(w, "maxy")) return md->bin.maxy;
  if (!strcmp(w, "minz")) return md->bin.minz;
  if (!strcmp(w, "maxz")) return md->bin.maxz;
  if (!strcmp(w, "lj1")) return md->lj.lj1[0];
  if (!strcmp(w, "lj2")) return md->lj.lj2[0];
  if (!strcmp(w, "cutsq")) return md->lj.cutsq[0];
  if (!strcmp(w, "mass0")) return s->mass[0];
  return 0.0 / 0.0;
}

/* copy an array out; returns the number of BYTES copied (or -1) */
long long orcf_copy(void *h, const char *w, void *out) {
  orc_md *md = (orc_md *)h;
  const orc_system *s = &md->sys;
  const size_t na = (size_t)s->N_local + (size_t)s->N_ghost;
  const void *src = NULL;
  size_t bytes = 0;
  if (!strcmp(w, "x")) { src = s->x; bytes = 24 * na; }
  else if (!strcmp(w, "v")) { src = s->v; bytes = 24 * na; }
  else if (!strcmp(w, "f")) { src = s->f; bytes = 24 * na; }
  else if (!strcmp(w, "q")) { src = s->q; bytes = 8 * na; }
  else if (!strcmp(w, "type")) { src = s->type; bytes = 4 * na; }
  else if (!strcmp(w, "id")) { src = s->id; bytes = 4 * na; }
  else if (!strcmp(w, "mass")) { src = s->mass; bytes = 8 * (size_t)s->ntypes; }
  else if (!strcmp(w, "bincount")) { src = md->bin.bincount; bytes = 4 * (size_t)md->bin.nbinx * md->bin.nbiny * md->bin.nbinz; }
  else if (!strcmp(w, "binoffsets")) { src = md->bin.binoffsets; bytes = 4 * (size_t)md->bin.nbinx * md->bin.nbiny * md->bin.nbinz; }
  else if (!strcmp(w, "permute")) { src = md->bin.permute_vector; bytes = 4 * (size_t)(md->bin.range_end - md->bin.range_begin); }
  else if (!strcmp(w, "row_map")) { src = md->neigh.row_map; bytes = 4 * ((size_t)md->neigh.N_local + 1); }
  else if (!strcmp(w, "entries")) { src = md->neigh.entries; bytes = 4 * (size_t)md->neigh.total; }
  else if (!strcmp(w, "num_neighs")) { src = md->neigh.num_neighs; bytes = 4 * (size_t)md->neigh.N_local; }
  else if (!strcmp(w, "neighs2d")) { src = md->neigh.neighs2d; bytes = 4 * (size_t)md->neigh.N_local * (size_t)md->neigh.maxneighs; }
  else if (!strncmp(w, "pack", 4)) { int p = w[4] - '0'; src = md->comm.pack_indicies[p]; bytes = 4 * (size_t)md->comm.num_ghost[p]; }
  else return -1;
  if (out && bytes) memcpy(out, src, bytes);
  return (long long)bytes;
}

/* overwrite oracle state arrays (first n atoms) -- used to feed GPU-produced inputs back */
int orcf_set(void *h, const char *w, const void *in, int n) {
  orc_md *md = (orc_md *)h;
  orc_system *s = &md->sys;
  if (!strcmp(w, "x")) memcpy(s->x, in, 24 * (size_t)n);
  else if (!strcmp(w, "v")) memcpy(s->v, in, 24 * (size_t)n);
  else if (!strcmp(w, "f")) memcpy(s->f, in, 24 * (size_t)n);
  else return -1;
  return 0;
}

int orcf_dump(void *h, const char *path, int step) { return orc_dump_binary(&((orc_md *)h)->sys, path, step, 0); }

/* SNAP function-level probe (see orc_force_snap_probe); returns ninside, -1 without a SNAP force */
int orcf_snap_probe(void *h, int i, double *utot_r, double *utot_i, int *inside, double *fij) {
  orc_md *md = (orc_md *)h;
  if (!md->snap) return -1;
  return orc_force_snap_probe(md->snap, &md->sys, &md->neigh, i, utot_r, utot_i, inside, fij);
}
int orcf_snap_jdim(void *h) { orc_md *md = (orc_md *)h; return md->snap ? orc_force_snap_jdim(md->snap) : -1; }
