// examinimd.cpp -- module factory, time loop, thermo output, binary dump and correctness report.
// Call order and stdout formats follow the reference driver (src/examinimd.cpp:60-294,296-482);
// device work is stream-ordered and the loop only synchronises where the reference reads a
// value back (thermo steps, rebuild counts, dumps).
#include "examinimd.h"
#include "property.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

#define MAXPATHLEN 1024

static double wall_seconds() {
  using namespace std::chrono;
  return duration_cast<duration<double>>(steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------ PhaseTimers
PhaseTimers::PhaseTimers(emd_ctx *c) : enabled(true), ctx(c), next(0), cur(nullptr) {
  for (int p = 0; p < NPHASE; p++) seconds[p] = 0.0;
}
PhaseTimers::~PhaseTimers() { for (void *e : pool) emd_event_destroy(e); }
void *PhaseTimers::get() {
  if (next == pool.size()) { void *e = nullptr; emd_event_create(&e); pool.push_back(e); }
  return pool[next++];
}
void PhaseTimers::begin() {
  if (!enabled) return;
  if (!cur) { cur = get(); emd_event_record(ctx, cur); }
}
void PhaseTimers::end(int phase) {
  if (!enabled) return;
  void *b = get();
  emd_event_record(ctx, b);
  spans.push_back({cur, b, phase});
  cur = b; // phases are back to back on one stream: the end of one is the start of the next
  if (spans.size() >= 4096) {
    // fold what has accumulated; `cur` must stay a live, recorded event (the next end() measures from it), so it moves to
    // the front of the recycled pool instead of being dropped
    flush();
    std::swap(pool[0], *std::find(pool.begin(), pool.end(), b));
    next = 1;
    cur = b;
  }
}
void PhaseTimers::flush() {
  for (const Span &s : spans) {
    float ms = 0.f;
    emd_event_elapsed_ms(s.a, s.b, &ms); // synchronises on s.b
    seconds[s.phase] += 1e-3 * ms;
  }
  spans.clear();
  next = 0;
  cur = nullptr;
}

// -------------------------------------------------------------------- ExaMiniMD
ExaMiniMD::ExaMiniMD(int device, void *stream) {
  system = new System();
  if (emd_ctx_create(&system->ctx, device, stream)) {
    // no CPU fallback: the product path needs the GPU
    fprintf(stderr, "ExaMiniMD: cannot create device context: %s\n", emd_last_error());
    emd_host_exit(1);
  }
  system->init();
  input = new Input(system);
  integrator = NULL; force = NULL; neighbor = NULL; comm = NULL; binning = NULL;
  current_step = 0;
  quiet = false;
}

ExaMiniMD::~ExaMiniMD() {
  delete integrator; delete force; delete neighbor; delete comm; delete binning; delete input;
  emd_ctx *c = system ? system->ctx : nullptr;
  delete system;
  if (c) emd_ctx_destroy(c);
}

void ExaMiniMD::init(int argc, char *argv[]) {
  input->read_command_line_args(argc, argv);
  if (quiet) system->do_print = false;
  input->read_file();

  if (input->integrator_type == INTEGRATOR_NVE) integrator = new IntegratorNVE(system);
  if (input->binning_type == BINNING_KKSORT) binning = new BinningKKSort(system);

  // force / neighbor / comm factories: each module header contributes its own `else if`
  if (false) {}
#define FORCE_MODULES_INSTANTIATION
#include "modules_force.h"
#undef FORCE_MODULES_INSTANTIATION
  else { printf("Invalid ForceType\n"); emd_host_exit(1); }
  for (size_t l = 0; l < input->force_coeff_lines.size(); l++) {
    const int line = input->force_coeff_lines[l];
    force->init_coeff(input->input_data.words_in_line(line), input->input_data.words[line]);
  }

  if (false) {}
#define NEIGHBOR_MODULES_INSTANTIATION
#include "modules_neighbor.h"
#undef NEIGHBOR_MODULES_INSTANTIATION
  else { printf("Invalid NeighborType\n"); emd_host_exit(1); }

  if (false) {}
#define COMM_MODULES_INSTANTIATION
#include "modules_comm.h"
#undef COMM_MODULES_INSTANTIATION
  else { printf("Invalid CommType\n"); emd_host_exit(1); }

  force->comm_newton = input->comm_newton;
  if (neighbor) neighbor->comm_newton = input->comm_newton;

  if (system->do_print) printf("Using: %s %s %s %s\n", force->name(), neighbor->name(), comm->name(), binning->name());

  if (system->N == 0) input->create_lattice(comm);

  // initial wrap, sort, ghosts, bins over local+ghost, neighbor list, forces (examinimd.cpp:120-144)
  const T_F_FLOAT neigh_cutoff = input->force_cutoff + input->neighbor_skin;
  comm->exchange();
  binning->create_binning(neigh_cutoff, neigh_cutoff, neigh_cutoff, 1, true, false, true);
  comm->exchange_halo();
  binning->create_binning(neigh_cutoff, neigh_cutoff, neigh_cutoff, 1, true, true, false);
  if (neighbor) neighbor->create_neigh_list(system, binning, force->half_neigh, false);
  if (!force->zeroes_forces())
    emd_memset_zero(system->ctx, system->f, sizeof(T_F_FLOAT) * 3 * (size_t)system->N_max);
  force->compute(system, binning, neighbor);
  if (input->comm_newton) comm->update_force();

  const int step = 0;
  if (input->thermo_rate > 0) {
    T_FLOAT T, PE, KE;
    thermo(&T, &PE, &KE);
    if (system->do_print) {
      if (!system->print_lammps) {
        printf("\n#Timestep Temperature PotE ETot Time Atomsteps/s\n");
        printf("%i %lf %lf %lf %lf %e\n", step, T, PE, PE + KE, 0.0, 0.0);
      } else {
        printf("\nStep Temp E_pair TotEng CPU\n");
        printf("     %i %lf %lf %lf %lf\n", step, T, PE, PE + KE, 0.0);
      }
    }
  }
  if (input->dumpbinaryflag) dump_binary(step);
  if (input->correctnessflag) check_correctness(step);
  emd_ctx_sync(system->ctx);
}

void ExaMiniMD::thermo(T_FLOAT *T, T_FLOAT *PE, T_FLOAT *KE) {
  Temperature temp(comm);
  PotE pote(comm);
  KinE kine(comm);
  *T = temp.compute(system);
  *PE = pote.compute(system, binning, neighbor, force) / system->N;
  *KE = kine.compute(system) / system->N;
}

void ExaMiniMD::step_once(int step, PhaseTimers *tm, bool fuse_next, bool observed) {
  const T_F_FLOAT neigh_cutoff = input->force_cutoff + input->neighbor_skin;
  static const bool overlap_halo = !(getenv("EMD_NO_OVERLAP") && atoi(getenv("EMD_NO_OVERLAP")));
  static const bool comm_first = !(getenv("EMD_OVERLAP_ORDER") && atoi(getenv("EMD_OVERLAP_ORDER")) == 0);
  bool split = false, split_kick = false;
  T_V_FLOAT nve_factors[2] = {0.0, 0.0};
  static const bool fuse_nve = !(getenv("EMD_NO_FUSED_NVE") && atoi(getenv("EMD_NO_FUSED_NVE")));
  if (tm) tm->begin();
  system->mv2_cached = false; // (a thermo sum left by the force launch of the step before)
  if (!initial_done) integrator->initial_integrate();
  initial_done = false;
  if (tm) tm->end(PhaseTimers::OTHER);

  if (step % input->comm_exchange_rate == 0 && step > 0) {
    comm->exchange();
    if (tm) tm->end(PhaseTimers::COMM);
    binning->create_binning(neigh_cutoff, neigh_cutoff, neigh_cutoff, 1, true, false, true);
    if (tm) tm->end(PhaseTimers::OTHER);
    comm->exchange_halo();
    if (tm) tm->end(PhaseTimers::COMM);
    binning->create_binning(neigh_cutoff, neigh_cutoff, neigh_cutoff, 1, true, true, false);
    if (neighbor) neighbor->create_neigh_list(system, binning, force->half_neigh, false);
    if (tm) tm->end(PhaseTimers::NEIGH);
  } else {
    // decomposed run: the share of the force that reads no ghost atom starts now, on the side stream, and overlaps the
    // halo exchange (the reference's blocking sequence update_halo -> compute, examinimd.cpp:226-235, otherwise)
    // peer-store transport + a force kernel that waits for the ghosts itself: nothing to overlap by hand, the refresh costs
    // its pack kernels and the neighbours' stores land while the ghost-free tiles are computed
    const bool gated = comm->num_processes() > 1 && force->gates_halo(system, neighbor) && comm->update_halo_deferred();
    split = !gated && overlap_halo && !observed && comm->num_processes() > 1 && force->can_split(system, neighbor);
    // the split force can take the integrator kick along like the single launch (Force::compute_with_nve)
    split_kick = split && fuse_next && fuse_nve && !input->comm_newton && integrator->step_factors(&nve_factors[0], &nve_factors[1]) &&
                 force->can_kick(system, neighbor);
    // the exchange's kernels are enqueued first (they find the SMs free); the side stream forks from the point before them
    if (split && emd_ctx_side_mark(system->ctx)) comm->error(emd_last_error());
    if (split && !comm_first) {
      if (emd_ctx_side_begin(system->ctx)) comm->error(emd_last_error());
      force->compute_part(system, binning, neighbor, 1, split_kick ? nve_factors : nullptr);
      if (emd_ctx_side_end(system->ctx)) comm->error(emd_last_error());
    }
    if (!gated) comm->update_halo();
    if (split && comm_first) {
      if (emd_ctx_side_begin(system->ctx)) comm->error(emd_last_error());
      force->compute_part(system, binning, neighbor, 1, split_kick ? nve_factors : nullptr);
      if (emd_ctx_side_end(system->ctx)) comm->error(emd_last_error());
    }
    if (tm) tm->end(PhaseTimers::COMM);
  }

  bool kicked = false; // the force launch already applied final_integrate + the next initial_integrate
  if (split) {
    force->compute_part(system, binning, neighbor, 2, split_kick ? nve_factors : nullptr);
    if (emd_ctx_side_join(system->ctx)) comm->error(emd_last_error());
    kicked = split_kick;
  } else {
    T_V_FLOAT dtf = 0.0, dtv = 0.0;
    if (fuse_next && fuse_nve && !input->comm_newton && integrator->step_factors(&dtf, &dtv)) {
      // on a thermo step the fused launch also sums the energy and m v^2 of the velocities between its two kicks (the state
      // thermo reads in the reference), so the step keeps the fused integrator; a force that cannot do that declines
      force->expect_energy(observed);
      kicked = force->compute_with_nve(system, binning, neighbor, dtf, dtv);
    }
    if (!kicked) {
      if (!force->zeroes_forces())
        emd_memset_zero(system->ctx, system->f, sizeof(T_F_FLOAT) * 3 * (size_t)system->N_max);
      force->expect_energy(observed); // thermo follows this step (run / run_quiet): one pass over the pairs may serve both
      force->compute(system, binning, neighbor);
    }
  }
  if (tm) tm->end(PhaseTimers::FORCE);

  if (input->comm_newton) {
    comm->update_force();
    if (tm) tm->end(PhaseTimers::COMM);
  }

  if (kicked) initial_done = true;
  else if (fuse_next && fuse_nve && !observed) { integrator->final_initial_integrate(); initial_done = true; }
  else integrator->final_integrate();
  if (tm) tm->end(PhaseTimers::OTHER);
}

void ExaMiniMD::advance(int nsteps) {
  for (int k = 0; k < nsteps; k++) step_once(++current_step, nullptr, k + 1 < nsteps);
}

// run() without output: what bench.py times as the reference's own metric (thermo passes included)
void ExaMiniMD::run_quiet(int nsteps, T_FLOAT *last_thermo3) {
  for (int s = 1; s <= nsteps; s++) {
    const int step = ++current_step;
    const bool observed = input->thermo_rate > 0 && step % input->thermo_rate == 0;
    step_once(step, nullptr, s < nsteps, observed); // (an observed step fuses only if the force launch can serve thermo itself)
    if (observed) {
      T_FLOAT T, PE, KE;
      thermo(&T, &PE, &KE);
      if (last_thermo3) { last_thermo3[0] = T; last_thermo3[1] = PE; last_thermo3[2] = KE; }
    }
  }
}

void ExaMiniMD::run(int nsteps) {
  PhaseTimers tm(system->ctx);
  emd_ctx_sync(system->ctx);
  const double t_start = wall_seconds();
  double last_time = 0.0;

  for (int s = 1; s <= nsteps; s++) {
    const int step = ++current_step;
    // dumps and the correctness check read x, v, f between two steps: no fusion across them; thermo output alone can be served
    // by the force launch (step_once)
    const bool thermo_step = input->thermo_rate > 0 && step % input->thermo_rate == 0;
    const bool state_read = input->dumpbinaryflag || input->correctnessflag; // these read x, v, f themselves
    step_once(step, &tm, s < nsteps && !state_read, thermo_step || state_read);

    if (input->thermo_rate > 0 && step % input->thermo_rate == 0) {
      T_FLOAT T, PE, KE;
      thermo(&T, &PE, &KE); // synchronises (host read-back), like the reference's reductions
      if (system->do_print) {
        const double time = wall_seconds() - t_start;
        if (!system->print_lammps)
          printf("%i %lf %lf %lf %lf %e\n", step, T, PE, PE + KE, time, 1.0 * system->N * input->thermo_rate / (time - last_time));
        else
          printf("     %i %lf %lf %lf %lf\n", step, T, PE, PE + KE, time);
        last_time = time;
      }
    }
    if (input->dumpbinaryflag) dump_binary(step);
    if (input->correctnessflag) check_correctness(step);
  }

  emd_ctx_sync(system->ctx);
  const double time = wall_seconds() - t_start;
  tm.flush();
  T_FLOAT T, PE, KE;
  thermo(&T, &PE, &KE);

  if (system->do_print) {
    if (!system->print_lammps) {
      printf("\n#Procs Particles | Time T_Force T_Neigh T_Comm T_Other | Steps/s Atomsteps/s Atomsteps/(proc*s)\n");
      printf("%i %i | %lf %lf %lf %lf %lf | %lf %e %e PERFORMANCE\n", comm->num_processes(), system->N, time,
             tm.seconds[PhaseTimers::FORCE], tm.seconds[PhaseTimers::NEIGH], tm.seconds[PhaseTimers::COMM],
             tm.seconds[PhaseTimers::OTHER], 1.0 * nsteps / time, 1.0 * system->N * nsteps / time,
             1.0 * system->N * nsteps / time / comm->num_processes());
    } else {
      printf("Loop time of %f on %i procs for %i steps with %i atoms\n", time, comm->num_processes(), nsteps, system->N);
    }
  }
}

// Binary dump: int n; int id[n]; int type[n]; double q[n]; double x[n][3]; v[n][3]; f[n][3]
// in PATH/output.<step:%010d>.<rank:%03d> (src/examinimd.cpp:296-346).
void ExaMiniMD::dump_binary(int step) {
  if (step % input->dumpbinary_rate) return;
  char filename[MAXPATHLEN];
  snprintf(filename, sizeof filename, "%s%s.%010d.%03d", input->dumpbinary_path, "/output", step, comm->process_rank());
  FILE *fp = fopen(filename, "wb");
  if (fp == NULL) {
    char str[MAXPATHLEN + 64];
    snprintf(str, sizeof str, "Cannot open dump file %s", filename);
    comm->error(str);
  }
  const T_INT n = system->N_local;
  HostAtoms h;
  system->download(h, n);
  fwrite(&n, sizeof(T_INT), 1, fp);
  fwrite(h.id.data(), sizeof(T_INT), n, fp);
  fwrite(h.type.data(), sizeof(T_INT), n, fp);
  fwrite(h.q.data(), sizeof(T_FLOAT), n, fp);
  fwrite(h.x.data(), sizeof(T_X_FLOAT), 3 * (size_t)n, fp);
  fwrite(h.v.data(), sizeof(T_V_FLOAT), 3 * (size_t)n, fp);
  fwrite(h.f.data(), sizeof(T_F_FLOAT), 3 * (size_t)n, fp);
  fclose(fp);
}

// Correctness report against a reference dump (src/examinimd.cpp:355-482): atoms are matched
// by id, the report line is `step |dr|_2 max|dr| |dv|_2 max|dv| |df|_2 max|df|`.  The reference
// matches ids with an O(n^2) scan; a sorted id index gives the same pairing in O(n log n).
void ExaMiniMD::check_correctness(int step) {
  if (step % input->correctness_rate) return;
  char filename[MAXPATHLEN];
  snprintf(filename, sizeof filename, "%s%s.%010d.%03d", input->reference_path, "/output", step, comm->process_rank());
  FILE *fpref = fopen(filename, "rb");
  if (fpref == NULL) {
    char str[MAXPATHLEN + 64];
    snprintf(str, sizeof str, "Cannot open input file %s", filename);
    comm->error(str);
  }
  const T_INT n = system->N_local;
  T_INT ntmp = 0;
  if (fread(&ntmp, sizeof(T_INT), 1, fpref) != 1 || ntmp != n) comm->error("Mismatch in current and reference atom counts");
  HostAtoms ref;
  ref.resize(n);
  size_t got = 0;
  got += fread(ref.id.data(), sizeof(T_INT), n, fpref);
  got += fread(ref.type.data(), sizeof(T_INT), n, fpref);
  got += fread(ref.q.data(), sizeof(T_FLOAT), n, fpref);
  got += fread(ref.x.data(), sizeof(T_X_FLOAT), 3 * (size_t)n, fpref);
  got += fread(ref.v.data(), sizeof(T_V_FLOAT), 3 * (size_t)n, fpref);
  got += fread(ref.f.data(), sizeof(T_F_FLOAT), 3 * (size_t)n, fpref);
  fclose(fpref);
  if (got != 12 * (size_t)n) comm->error("Short read of reference dump");

  HostAtoms cur;
  system->download(cur, n);
  std::vector<T_INT> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](T_INT a, T_INT b) { return cur.id[a] < cur.id[b]; });

  T_FLOAT sumsq[3] = {0.0, 0.0, 0.0}, maxd[3] = {0.0, 0.0, 0.0};
  for (T_INT i = 0; i < n; i++) {
    T_INT ii = -1;
    if (cur.id[i] == ref.id[i]) ii = i;
    else {
      auto it = std::lower_bound(order.begin(), order.end(), ref.id[i], [&](T_INT a, T_INT idv) { return cur.id[a] < idv; });
      if (it != order.end() && cur.id[*it] == ref.id[i]) ii = *it;
    }
    if (ii == -1) { printf("Unable to find current id matchinf reference id %d \n", ref.id[i]); continue; }
    const double *c[3] = {&cur.x[3 * (size_t)ii], &cur.v[3 * (size_t)ii], &cur.f[3 * (size_t)ii]};
    const double *r[3] = {&ref.x[3 * (size_t)i], &ref.v[3 * (size_t)i], &ref.f[3 * (size_t)i]};
    for (int q = 0; q < 3; q++) {
      const T_FLOAT dx = c[q][0] - r[q][0], dy = c[q][1] - r[q][1], dz = c[q][2] - r[q][2];
      sumsq[q] += dx * dx + dy * dy + dz * dz;
      maxd[q] = std::max(maxd[q], std::max(fabs(dx), std::max(fabs(dy), fabs(dz))));
    }
  }
  for (int q = 0; q < 3; q++) { comm->reduce_float(&sumsq[q], 1); comm->reduce_max_float(&maxd[q], 1); }

  if (comm->process_rank() == 0) { // one writer (quiet library sessions included)
    FILE *fpout = fopen(input->correctness_file, step == 0 ? "w" : "a");
    if (fpout) {
      if (step == 0) fprintf(fpout, "# timestep deltarnorm maxdelr deltavnorm maxdelv deltafnorm maxdelf\n");
      fprintf(fpout, "%d %g %g %g %g %g %g\n", step, sqrt(sumsq[0]), maxd[0], sqrt(sumsq[1]), maxd[1], sqrt(sumsq[2]), maxd[2]);
      fclose(fpout);
    }
  }
}

void ExaMiniMD::print_performance() {}

void ExaMiniMD::shutdown() { system->destroy(); }
