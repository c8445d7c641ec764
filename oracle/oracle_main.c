/*
 * oracle_main.c -- command-line driver for the CPU oracle (TEST INFRASTRUCTURE ONLY).
 * Mirrors the reference CLI (src/input.cpp:151-224) and its stdout tables
 * (src/examinimd.cpp:148-166,252-267,286-289) so that bench.py can time it as the
 * CPU baseline and tests can diff its --dumpbinary files.
 *
 *   oracle_md -il DECK [--neigh-type 2D|CSR|CSR_MAPCONSTR] [--force-iteration NEIGH_FULL|NEIGH_HALF]
 *             [--dumpbinary N PATH] [--nsteps K] [--region NX NY NZ]
 * (--nsteps/--region override the deck's `run`/`region` so scaled configs need no deck copy.)
 */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int main(int argc, char **argv) {
  const char *deck = NULL, *dump_path = NULL;
  int neigh = -1, iter = -1, dump_rate = 0, nsteps = -1, reg[3] = {0, 0, 0};
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-il") || !strcmp(argv[i], "--input-lammps")) deck = argv[++i];
    else if (!strcmp(argv[i], "--neigh-type")) {
      ++i;
      if (!strcmp(argv[i], "CSR")) neigh = ORC_NEIGH_CSR;
      else if (!strcmp(argv[i], "2D")) neigh = ORC_NEIGH_2D;
      else if (!strcmp(argv[i], "CSR_MAPCONSTR")) neigh = ORC_NEIGH_CSR_MAPCONSTR;
    } else if (!strcmp(argv[i], "--force-iteration")) {
      ++i;
      if (!strcmp(argv[i], "NEIGH_FULL")) iter = ORC_ITER_NEIGH_FULL;
      else if (!strcmp(argv[i], "NEIGH_HALF")) iter = ORC_ITER_NEIGH_HALF;
    } else if (!strcmp(argv[i], "--comm-type")) ++i;
    else if (!strcmp(argv[i], "--dumpbinary")) { dump_rate = atoi(argv[i + 1]); dump_path = argv[i + 2]; i += 2; }
    else if (!strcmp(argv[i], "--nsteps")) nsteps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--region")) { reg[0] = atoi(argv[i + 1]); reg[1] = atoi(argv[i + 2]); reg[2] = atoi(argv[i + 3]); i += 3; }
    else { fprintf(stderr, "ERROR: Unknown command line argument: %s\n", argv[i]); return 1; }
  }
  if (!deck) { fprintf(stderr, "usage: oracle_md -il DECK ...\n"); return 1; }

  /* init in two halves so --region can override the deck before the lattice is built */
  orc_md md;
  memset(&md, 0, sizeof md);
  orc_system_init(&md.sys);
  orc_input_defaults(&md.in);
  if (neigh >= 0) md.in.neighbor_type = neigh;
  if (iter >= 0) md.in.force_iteration_type = iter;
  if (orc_input_read_deck(&md.in, &md.sys, deck)) { fprintf(stderr, "cannot read %s\n", deck); return 1; }
  if (reg[0] > 0) { md.in.lattice_nx = reg[0]; md.in.lattice_ny = reg[1]; md.in.lattice_nz = reg[2]; }
  if (nsteps >= 0) md.in.nsteps = nsteps;
  const int half = md.in.force_iteration_type == ORC_ITER_NEIGH_HALF;
  md.neigh_cutoff = md.in.force_cutoff + md.in.neighbor_skin;
  orc_binning_init(&md.bin);
  if (md.in.force_type == ORC_FORCE_LJ) {
    orc_force_lj_init(&md.lj, md.sys.ntypes, half);
    for (int l = 0; l < md.in.n_coeff_lines; l++) orc_force_lj_init_coeff(&md.lj, md.in.coeff_nwords[l], md.in.coeff_words[l]);
    md.lj.comm_newton = md.in.comm_newton;
  } else if (md.in.force_type == ORC_FORCE_SNAP) {
    md.snap = orc_force_snap_create(md.sys.ntypes);
    for (int l = 0; l < md.in.n_coeff_lines; l++)
      if (orc_force_snap_init_coeff(md.snap, md.in.coeff_nwords[l], md.in.coeff_words[l], NULL)) { fprintf(stderr, "snap coeff error\n"); return 1; }
  } else { fprintf(stderr, "Invalid ForceType\n"); return 1; }
  orc_neighbor_init(&md.neigh, md.in.neighbor_type == ORC_NEIGH_2D ? ORC_NEIGH_2D : ORC_NEIGH_CSR, md.neigh_cutoff);
  md.neigh.comm_newton = md.in.comm_newton;
  orc_comm_serial_init(&md.comm, md.neigh_cutoff);
  orc_create_lattice(&md.in, &md.sys);
  printf("Atoms: %i %i\n", md.sys.N, md.sys.N_local);
  orc_md_setup(&md);

  int threads = 1;
#ifdef _OPENMP
  threads = omp_get_max_threads();
#endif
  printf("Using: oracle(%s,%s) threads %d\n", md.snap ? "SNAP" : (half ? "LJ-half" : "LJ-full"),
         md.neigh.kind == ORC_NEIGH_2D ? "2D" : "CSR", threads);

  double T, PE, KE;
  if (md.in.thermo_rate > 0) {
    orc_md_thermo(&md, &T, &PE, &KE);
    printf("\n#Timestep Temperature PotE ETot Time Atomsteps/s\n");
    printf("%i %lf %lf %lf %lf %e\n", 0, T, PE, PE + KE, 0.0, 0.0);
  }
  if (dump_rate) orc_dump_binary(&md.sys, dump_path, 0, 0);

  double t_start = now_s(), last_time = 0.0;
  for (int step = 1; step <= md.in.nsteps; step++) {
    orc_md_step(&md);
    if (md.in.thermo_rate > 0 && step % md.in.thermo_rate == 0) {
      orc_md_thermo(&md, &T, &PE, &KE);
      double time = now_s() - t_start;
      printf("%i %lf %lf %lf %lf %e\n", step, T, PE, PE + KE, time, 1.0 * md.sys.N * md.in.thermo_rate / (time - last_time));
      last_time = time;
    }
    if (dump_rate && step % dump_rate == 0) orc_dump_binary(&md.sys, dump_path, step, 0);
  }
  double time = now_s() - t_start;
  printf("\n#Procs Particles | Time T_Force T_Neigh T_Comm T_Other | Steps/s Atomsteps/s Atomsteps/(proc*s)\n");
  printf("%i %i | %lf %lf %lf %lf %lf | %lf %e %e PERFORMANCE\n", 1, md.sys.N, time, md.t_force, md.t_neigh, md.t_comm,
         md.t_other, 1.0 * md.in.nsteps / time, 1.0 * md.sys.N * md.in.nsteps / time, 1.0 * md.sys.N * md.in.nsteps / time);
  orc_md_destroy(&md);
  return 0;
}
