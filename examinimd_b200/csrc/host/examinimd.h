// examinimd.h -- the application object: owns the modules and runs the time loop
// (role of src/examinimd.h:50-71).  Can be used as a library (see capi.cpp) or through main.cpp.
#pragma once
#include "types.h"
#include "system.h"
#include "integrator.h"
#include "force.h"
#include "neighbor.h"
#include "comm.h"
#include "input.h"
#include "binning.h"
#include <vector>

// accumulates device time per phase from event pairs recorded on the module stream;
// events are resolved lazily so the time loop never blocks on them
class PhaseTimers {
public:
  enum { FORCE, NEIGH, COMM, OTHER, NPHASE };
  PhaseTimers(emd_ctx *ctx);
  ~PhaseTimers();
  void begin();
  void end(int phase);
  void flush();
  double seconds[NPHASE];
  bool enabled;

private:
  emd_ctx *ctx;
  std::vector<void *> pool;
  struct Span { void *a, *b; int phase; };
  std::vector<Span> spans;
  size_t next;
  void *cur;
  void *get();
};

class ExaMiniMD {
public:
  System *system;
  Integrator *integrator;
  Force *force;
  Neighbor *neighbor;
  Comm *comm;
  Input *input;
  Binning *binning;

  ExaMiniMD(int device = 0, void *stream = nullptr);
  ~ExaMiniMD();

  void init(int argc, char *argv[]);
  void run(int nsteps);

  // one timestep of the loop body of run() (examinimd.cpp:192-250), without output
  // fuse_next: the next step follows with nothing observing the state in between, so this step ends with
  // Integrator::final_initial_integrate() and the next one skips its initial_integrate()
  // observed: thermo output follows this step (the force module may evaluate the energy in the same pass, Force::expect_energy)
  void step_once(int step, PhaseTimers *timers, bool fuse_next = false, bool observed = false);
  bool initial_done = false; // the pending step's initial_integrate has already been applied
  // advance `nsteps` steps continuing the global step counter (rebuild cadence preserved)
  void advance(int nsteps);
  // run() without its output: steps + the thermo reductions at the deck's cadence (the work the reference's Atomsteps/s spans)
  void run_quiet(int nsteps, T_FLOAT *last_thermo3);
  void thermo(T_FLOAT *T, T_FLOAT *PE, T_FLOAT *KE);

  void dump_binary(int);
  void check_correctness(int);
  void print_performance();
  void shutdown();

  int current_step;
  bool quiet; // suppress stdout tables (library use)
};
