// input.h -- command line + restricted LAMMPS deck + lattice/velocity creation
// (role of src/input.{h,cpp}).  Host code only; must yield bit-identical initial x, v, type, id.
#pragma once
#include <string>
#include <vector>
#include "comm.h"
#include "system.h"
#include "types.h"

// the deck split into words; force modules receive rows of it as `char** args`
// (same shape limits as src/input.cpp:44-118: 100 lines x 32 words x 31 chars)
class ItemizedFile {
public:
  char ***words;
  int max_nlines, nlines, words_per_line, max_word_size;
  ItemizedFile();
  ~ItemizedFile();
  void allocate_words(int num_lines);
  void free_words();
  void print_line(int line);
  int words_in_line(int line);
  void print();
  void add_line(const char *const line);
};

// LAMMPS "velocity ... loop geom" generator: per-atom Park-Miller stream seeded by a hash of
// (seed, position) -- src/input.h:66-133
class LAMMPS_RandomVelocityGeom {
  int seed;
public:
  LAMMPS_RandomVelocityGeom() : seed(0) {}
  double uniform();
  void reset(int ibase, double *coord);
};

class Input {
  bool timestepflag;

public:
  System *system;
  char *input_file;
  int input_file_type;
  ItemizedFile input_data;

  int units;
  int lattice_style;
  double lattice_constant, lattice_offset_x, lattice_offset_y, lattice_offset_z;
  int lattice_nx, lattice_ny, lattice_nz;
  double temperature_target;
  int temperature_seed;
  int integrator_type, nsteps;
  int binning_type;
  int comm_type, comm_exchange_rate, comm_newton;
  int force_type, force_iteration_type, force_line;
  T_F_FLOAT force_cutoff;
  std::vector<int> force_coeff_lines;
  T_F_FLOAT neighbor_skin;
  int neighbor_type;
  int thermo_rate, dumpbinary_rate, correctness_rate;
  bool dumpbinaryflag, correctnessflag;
  char *dumpbinary_path, *reference_path, *correctness_file;
  // extensions (not in the reference): override the deck's region / run from the command line
  int override_region[3], override_nsteps;

  Input(System *s);
  void read_command_line_args(int argc, char *argv[]);
  void read_file(const char *filename = NULL);
  void read_lammps_file(const char *filename);
  void check_lammps_command(int line);
  void create_lattice(Comm *comm);
  void create_lattice_device(Comm *comm); // the same on the device (one atom type), kernels/lattice.cu
};
