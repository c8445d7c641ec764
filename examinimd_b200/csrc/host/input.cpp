// input.cpp -- CLI, deck parser and initial-condition generator of the host layer.
// Behavioural contract (what must match the reference, src/input.cpp):
//   * flags  -il/--input-lammps, --force-iteration, --comm-type, --neigh-type, --dumpbinary,
//     --correctness, -h/--help; "--kokkos-*" ignored; anything else is fatal (:151-224)
//   * deck keywords and their side effects (:265-458), incl. `units metal` always resetting dt
//   * sc/fcc lattice sites, types, ids (:460-725) and "loop geom" velocities, momentum removal
//     and temperature rescale (:731-785), all evaluated with the same double expressions.
#include "input.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>

// ---------------------------------------------------------------- ItemizedFile
ItemizedFile::ItemizedFile() : words(NULL), max_nlines(0), nlines(0), words_per_line(32), max_word_size(32) {}
ItemizedFile::~ItemizedFile() { free_words(); }

void ItemizedFile::allocate_words(int num_lines) {
  free_words();
  max_nlines = num_lines;
  nlines = 0;
  words = new char **[max_nlines];
  for (int l = 0; l < max_nlines; l++) {
    words[l] = new char *[words_per_line];
    for (int w = 0; w < words_per_line; w++) {
      words[l][w] = new char[max_word_size + 1];
      words[l][w][0] = 0;
    }
  }
}

void ItemizedFile::free_words() {
  if (!words) return;
  for (int l = 0; l < max_nlines; l++) {
    for (int w = 0; w < words_per_line; w++) delete[] words[l][w];
    delete[] words[l];
  }
  delete[] words;
  words = NULL;
  max_nlines = 0;
}

int ItemizedFile::words_in_line(int l) {
  int n = 0;
  for (int w = 0; w < words_per_line; w++) n += words[l][w][0] != 0;
  return n;
}

void ItemizedFile::print_line(int l) {
  for (int w = 0; w < words_per_line; w++)
    if (words[l][w][0]) printf("%s ", words[l][w]);
  printf("\n");
}

void ItemizedFile::print() {
  for (int l = 0; l < nlines; l++) print_line(l);
}

// split on blanks/tabs; words longer than max_word_size-1 are truncated
void ItemizedFile::add_line(const char *const line) {
  if (nlines < max_nlines) {
    const char *p = line;
    int w = 0;
    while (*p && w < words_per_line) {
      while (*p == ' ' || *p == '\t') p++;
      int k = 0;
      while (*p && *p != ' ' && *p != '\t') {
        if (k < max_word_size - 1) words[nlines][w][k++] = *p;
        p++;
      }
      words[nlines][w][k] = 0;
      w++;
    }
  }
  nlines++;
}

// ---------------------------------------------------- LAMMPS_RandomVelocityGeom
// Park-Miller minimal standard generator with Schrage's factorisation (src/input.h:78-85)
double LAMMPS_RandomVelocityGeom::uniform() {
  const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836;
  const int k = seed / IQ;
  seed = IA * (seed - k * IQ) - IR * k;
  if (seed < 0) seed += IM;
  return (1.0 / IM) * seed;
}

// Jenkins one-at-a-time hash over the bytes of (ibase, coord[3]); bytes are read as plain
// (signed) char like the reference; 27-bit mask and five warm-up draws (src/input.h:100-132)
void LAMMPS_RandomVelocityGeom::reset(int ibase, double *coord) {
  unsigned int hash = 0;
  auto mix = [&hash](const char *bytes, int n) {
    for (int i = 0; i < n; i++) {
      hash += bytes[i];
      hash += (hash << 10);
      hash ^= (hash >> 6);
    }
  };
  mix(reinterpret_cast<const char *>(&ibase), (int)sizeof(int));
  mix(reinterpret_cast<const char *>(coord), (int)(3 * sizeof(double)));
  hash += (hash << 3);
  hash ^= (hash >> 11);
  hash += (hash << 15);
  seed = hash & 0x7ffffff;
  if (!seed) seed = 1;
  for (int i = 0; i < 5; i++) uniform();
}

// ------------------------------------------------------------------------ Input
Input::Input(System *p) : system(p), integrator_type(INTEGRATOR_NVE) {
  timestepflag = false;
  input_file = NULL;
  input_file_type = -1;
  units = UNITS_LJ;
  lattice_style = LATTICE_FCC;
  lattice_constant = 0.0;
  lattice_offset_x = lattice_offset_y = lattice_offset_z = 0.0;
  lattice_nx = lattice_ny = lattice_nz = 0;
  temperature_target = 0.0;
  temperature_seed = 0;
  nsteps = 0;
  binning_type = BINNING_KKSORT;
  comm_type = COMM_SERIAL; // the reference defaults to MPI only when compiled with it (:126-130)
  comm_exchange_rate = 20;
  comm_newton = 0;
  force_type = FORCE_LJ;
  force_iteration_type = FORCE_ITER_NEIGH_FULL;
  force_line = 0;
  force_cutoff = 0.0;
  neighbor_skin = 0.0;
  neighbor_type = NEIGH_2D;
  thermo_rate = dumpbinary_rate = correctness_rate = 0;
  dumpbinaryflag = correctnessflag = false;
  dumpbinary_path = reference_path = correctness_file = NULL;
  override_region[0] = override_region[1] = override_region[2] = 0;
  override_nsteps = -1;
}

static void print_help() {
  printf("ExaMiniMD-B200 1.0 (sm_100a CUDA version)\n\n");
  printf("Options:\n");
  printf("  -il [file] / --input-lammps [FILE]: Provide LAMMPS input file\n");
  printf("  --force-iteration [TYPE]:   Specify which iteration style to use\n");
  printf("                              for force calculations (CELL_FULL, NEIGH_FULL, NEIGH_HALF)\n");
  printf("  --comm-type [TYPE]:         Specify Communication Routines implementation \n");
  printf("                              (MPI, SERIAL)\n");
  printf("  --dumpbinary [N] [PATH]:    Request that binary output files PATH/output* be generated every N steps\n");
  printf("  --correctness [N] [PATH] [FILE]:   Request that correctness check against files PATH/output* be performed every N steps, correctness data written to FILE\n");
  printf("  --neigh-type [TYPE]:        Specify Neighbor Routines implementation \n");
  printf("                              (2D, CSR, CSR_MAPCONSTR)\n");
  printf("  --region NX NY NZ / --nsteps K: override the deck's region / run (extension)\n");
}

void Input::read_command_line_args(int argc, char *argv[]) {
#define MODULES_OPTION_CHECK
  for (int i = 1; i < argc; i++) {
    const char *a = argv[i];
    if (!strcmp(a, "-h") || !strcmp(a, "--help")) {
      if (system->do_print) print_help();
    } else if (!strcmp(a, "-il") || !strcmp(a, "--input-lammps")) {
      input_file = argv[++i];
      input_file_type = INPUT_LAMMPS;
    } else if (!strcmp(a, "--force-iteration") || !strcmp(a, "--force-type")) {
#include "modules_force.h"
      ++i;
    } else if (!strcmp(a, "--comm-type")) {
#include "modules_comm.h"
      ++i;
    } else if (!strcmp(a, "--neigh-type") || !strcmp(a, "--neighbor-type")) {
#include "modules_neighbor.h"
      ++i;
    } else if (!strcmp(a, "--dumpbinary")) {
      dumpbinary_rate = atoi(argv[i + 1]);
      dumpbinary_path = argv[i + 2];
      dumpbinaryflag = true;
      i += 2;
    } else if (!strcmp(a, "--correctness")) {
      correctness_rate = atoi(argv[i + 1]);
      reference_path = argv[i + 2];
      correctness_file = argv[i + 3];
      correctnessflag = true;
      i += 3;
    } else if (!strcmp(a, "--region")) {
      for (int d = 0; d < 3; d++) override_region[d] = atoi(argv[i + 1 + d]);
      i += 3;
    } else if (!strcmp(a, "--nsteps")) {
      override_nsteps = atoi(argv[++i]);
    } else if (strstr(a, "--kokkos-") == NULL) {
      if (system->do_print) printf("ERROR: Unknown command line argument: %s\n", a);
      emd_host_exit(1);
    }
  }
#undef MODULES_OPTION_CHECK
}

void Input::read_file(const char *filename) {
  if (filename == NULL) filename = input_file;
  if (input_file_type == INPUT_LAMMPS) {
    read_lammps_file(filename);
    return;
  }
  if (system->do_print) printf("ERROR: Unknown input file type\n");
  emd_host_exit(1);
}

void Input::read_lammps_file(const char *filename) {
  {
    const std::string path(filename);
    const size_t slash = path.find_last_of('/');
    system->input_dir = slash == std::string::npos ? std::string(".") : path.substr(0, slash);
  }
  input_data.allocate_words(100);
  std::ifstream file(filename);
  if (!file.good()) {
    if (system->do_print) printf("ERROR: cannot open input file %s\n", filename);
    emd_host_exit(1);
  }
  std::string line;
  while (std::getline(file, line)) {
    if (line.size() > 510) line.resize(510);
    input_data.add_line(line.c_str());
  }
  if (system->do_print) {
    printf("\n#InputFile:\n#=========================================================\n");
    input_data.print();
    printf("#=========================================================\n\n");
  }
  const int n = input_data.nlines < input_data.max_nlines ? input_data.nlines : input_data.max_nlines;
  for (int l = 0; l < n; l++) check_lammps_command(l);
  if (override_region[0] > 0) { lattice_nx = override_region[0]; lattice_ny = override_region[1]; lattice_nz = override_region[2]; }
  if (override_nsteps >= 0) nsteps = override_nsteps;
}

void Input::check_lammps_command(int line) {
  char **w = input_data.words[line];
  const bool print = system->do_print;
  if (w[0][0] == 0 || strchr(w[0], '#')) return;
  auto is = [&](int k, const char *s) { return strcmp(w[k], s) == 0; };
  const std::string key = w[0];
  bool known = true;

  if (key == "units") {
    if (is(1, "metal")) { units = UNITS_METAL; system->boltz = 8.617343e-5; system->mvv2e = 1.0364269e-4; system->dt = 0.001; }
    else if (is(1, "real")) { units = UNITS_REAL; system->boltz = 0.0019872067; system->mvv2e = 48.88821291 * 48.88821291; if (!timestepflag) system->dt = 1.0; }
    else if (is(1, "lj")) { units = UNITS_LJ; system->boltz = 1.0; system->mvv2e = 1.0; if (!timestepflag) system->dt = 0.005; }
    else { known = false; if (print) printf("LAMMPS-Command: 'units' command only supports 'real' and 'lj' in ExaMiniMD\n"); }
  } else if (key == "atom_style") {
    if (!is(1, "atomic")) { known = false; if (print) printf("LAMMPS-Command: 'atom_style' command only supports 'atomic' in ExaMiniMD\n"); }
  } else if (key == "lattice") {
    if (is(1, "sc")) { lattice_style = LATTICE_SC; lattice_constant = atof(w[2]); }
    else if (is(1, "fcc")) { lattice_style = LATTICE_FCC; lattice_constant = std::pow((4.0 / atof(w[2])), (1.0 / 3.0)); }
    else { known = false; if (print) printf("LAMMPS-Command: 'lattice' command only supports 'sc' and 'fcc' in ExaMiniMD\n"); }
    if (is(3, "origin")) { lattice_offset_x = atof(w[4]); lattice_offset_y = atof(w[5]); lattice_offset_z = atof(w[6]); }
  } else if (key == "region") {
    if (is(2, "block")) {
      if ((atoi(w[3]) != 0 || atoi(w[5]) != 0 || atoi(w[7]) != 0) && print)
        printf("Error: LAMMPS-Command: region only allows for boxes with 0,0,0 offset\n");
      lattice_nx = atoi(w[4]); lattice_ny = atoi(w[6]); lattice_nz = atoi(w[8]);
    } else { known = false; if (print) printf("LAMMPS-Command: 'region' command only supports 'block' option in ExaMiniMD\n"); }
  } else if (key == "create_box") {
    system->ntypes = atoi(w[1]);
    system->h_mass.assign(system->ntypes, 0.0);
  } else if (key == "create_atoms") {
  } else if (key == "mass") {
    const int t = atoi(w[1]) - 1;
    if (t >= 0 && t < (int)system->h_mass.size()) system->h_mass[t] = atof(w[2]);
  } else if (key == "pair_style") {
    known = false;
    if (is(1, "lj/cut/idial")) { known = true; force_type = FORCE_LJ_IDIAL; force_cutoff = atof(w[2]); force_line = line; }
    else if (is(1, "lj/cut")) { known = true; force_type = FORCE_LJ; force_cutoff = atof(w[2]); force_line = line; }
    if (is(1, "snap")) { known = true; force_type = FORCE_SNAP; force_cutoff = 4.73442; /* hard-wired, :381 */ force_line = line; }
    if (print && !known) printf("LAMMPS-Command: 'pair_style' command only supports 'lj/cut', 'lj/cut/idial', and 'snap' style in ExaMiniMD\n");
  } else if (key == "pair_coeff") {
    force_coeff_lines.push_back(line);
  } else if (key == "velocity") {
    if (!is(1, "all") && print) printf("Error: LAMMPS-Command: 'velocity' command can only be applied to 'all'\n");
    if (!is(2, "create") && print) printf("Error: LAMMPS-Command: 'velocity' command can only be used with option 'create'\n");
    temperature_target = atof(w[3]);
    temperature_seed = atoi(w[4]);
  } else if (key == "neighbor") {
    neighbor_skin = atof(w[1]);
  } else if (key == "neigh_modify") {
    for (int i = 1; i < input_data.words_per_line - 1; i++)
      if (is(i, "every")) comm_exchange_rate = atoi(w[i + 1]);
  } else if (key == "fix") {
    if (is(3, "nve")) integrator_type = INTEGRATOR_NVE;
    else { known = false; if (print) printf("LAMMPS-Command: 'fix' command only supports 'nve' style in ExaMiniMD\n"); }
  } else if (key == "run") {
    nsteps = atoi(w[1]);
  } else if (key == "thermo") {
    thermo_rate = atoi(w[1]);
  } else if (key == "timestep") {
    system->dt = atof(w[1]);
    timestepflag = true;
  } else if (key == "newton") {
    if (is(1, "on")) comm_newton = 1;
    else if (is(1, "off")) comm_newton = 0;
    else if (print) printf("LAMMPS-Command: 'newton' must be followed by 'on' or 'off'\n");
  } else if (key == "variable") {
    known = false;
    if (print) printf("LAMMPS-Command: 'variable' keyword is not supported in ExaMiniMD\n");
  } else
    known = false;

  if (!known && print) {
    printf("ERROR: unknown keyword\n");
    input_data.print_line(line);
  }
}

// Lattice sites inside this rank's brick, in the reference's loop order (z, y, x, basis).
// One generator serves sc (1-atom basis, a*(i+offset)) and fcc (4-atom basis, a*(1.0*i+basis)).
namespace {
struct Site { double x, y, z; };

template <class Visit>
void for_each_site(const Input &in, const System &s, Visit visit) {
  const double a = in.lattice_constant;
  const T_INT ix0 = s.sub_domain_lo_x / s.domain_x * in.lattice_nx - 0.5, ix1 = s.sub_domain_hi_x / s.domain_x * in.lattice_nx + 0.5;
  const T_INT iy0 = s.sub_domain_lo_y / s.domain_y * in.lattice_ny - 0.5, iy1 = s.sub_domain_hi_y / s.domain_y * in.lattice_ny + 0.5;
  const T_INT iz0 = s.sub_domain_lo_z / s.domain_z * in.lattice_nz - 0.5, iz1 = s.sub_domain_hi_z / s.domain_z * in.lattice_nz + 0.5;
  const bool fcc = in.lattice_style == LATTICE_FCC;
  double basis[4][3] = {{0.0, 0.0, 0.0}, {0.5, 0.5, 0.0}, {0.5, 0.0, 0.5}, {0.0, 0.5, 0.5}};
  for (int k = 0; k < 4; k++) { basis[k][0] += in.lattice_offset_x; basis[k][1] += in.lattice_offset_y; basis[k][2] += in.lattice_offset_z; }
  const int nbasis = fcc ? 4 : 1;
  for (T_INT iz = iz0; iz <= iz1; iz++)
    for (T_INT iy = iy0; iy <= iy1; iy++)
      for (T_INT ix = ix0; ix <= ix1; ix++)
        for (int k = 0; k < nbasis; k++) {
          Site p;
          if (fcc) { p.x = a * (1.0 * ix + basis[k][0]); p.y = a * (1.0 * iy + basis[k][1]); p.z = a * (1.0 * iz + basis[k][2]); }
          else { p.x = a * (ix + in.lattice_offset_x); p.y = a * (iy + in.lattice_offset_y); p.z = a * (iz + in.lattice_offset_z); }
          if (p.x >= s.sub_domain_lo_x && p.y >= s.sub_domain_lo_y && p.z >= s.sub_domain_lo_z && p.x < s.sub_domain_hi_x &&
              p.y < s.sub_domain_hi_y && p.z < s.sub_domain_hi_z)
            visit(p);
        }
}
} // namespace

void Input::create_lattice_device(Comm *comm) {
  System &s = *system;
  emd_lattice lat;
  // index ranges exactly as the host loops compute them (truncation of the double expressions, :485-490 / :604-609)
  const T_INT ix0 = s.sub_domain_lo_x / s.domain_x * lattice_nx - 0.5, ix1 = s.sub_domain_hi_x / s.domain_x * lattice_nx + 0.5;
  const T_INT iy0 = s.sub_domain_lo_y / s.domain_y * lattice_ny - 0.5, iy1 = s.sub_domain_hi_y / s.domain_y * lattice_ny + 0.5;
  const T_INT iz0 = s.sub_domain_lo_z / s.domain_z * lattice_nz - 0.5, iz1 = s.sub_domain_hi_z / s.domain_z * lattice_nz + 0.5;
  lat.i0[0] = ix0; lat.i0[1] = iy0; lat.i0[2] = iz0;
  lat.n[0] = ix1 - ix0 + 1; lat.n[1] = iy1 - iy0 + 1; lat.n[2] = iz1 - iz0 + 1;
  lat.fcc = lattice_style == LATTICE_FCC ? 1 : 0;
  lat.a = lattice_constant;
  lat.offset[0] = lattice_offset_x; lat.offset[1] = lattice_offset_y; lat.offset[2] = lattice_offset_z;
  lat.lo[0] = s.sub_domain_lo_x; lat.lo[1] = s.sub_domain_lo_y; lat.lo[2] = s.sub_domain_lo_z;
  lat.hi[0] = s.sub_domain_hi_x; lat.hi[1] = s.sub_domain_hi_y; lat.hi[2] = s.sub_domain_hi_z;
  auto die = [&](const char *what) { fprintf(stderr, "Input::create_lattice: %s: %s\n", what, emd_last_error()); emd_host_exit(1); };
  int n = 0;
  if (emd_lattice_count(s.ctx, &lat, &n)) die("count");
  s.N_local = n;
  s.N = n;
  s.grow(n + n / 4 + 16); // head-room for ghosts (the reference ends up with 2n, :497-534)
  // global atom count and globally unique ids (:578-584 / :714-721)
  T_INT N_local_offset = n;
  comm->scan_int(&N_local_offset, 1);
  comm->reduce_int(&s.N, 1);
  if (s.do_print) printf("Atoms: %i %i\n", s.N, s.N_local);
  if (emd_lattice_fill(s.ctx, &lat, temperature_seed, N_local_offset - n, s.mass, s.x, s.v, s.q, s.type, s.id)) die("fill");
  if (emd_memset_zero(s.ctx, s.f, sizeof(T_F_FLOAT) * 3 * (size_t)n)) die("zero f");
  // centre-of-mass velocity out, then rescale to the target temperature (:731-785); sums in atom order
  double m4[4];
  if (emd_velocity_sums(s.ctx, s.v, s.type, s.mass, n, 0, m4)) die("momentum sums");
  T_FLOAT total_mass = m4[0], px = m4[1], py = m4[2], pz = m4[3];
  comm->reduce_float(&px, 1);
  comm->reduce_float(&py, 1);
  comm->reduce_float(&pz, 1);
  comm->reduce_float(&total_mass, 1);
  if (emd_velocity_shift(s.ctx, s.v, n, px / total_mass, py / total_mass, pz / total_mass)) die("shift");
  if (emd_velocity_sums(s.ctx, s.v, s.type, s.mass, n, 1, m4)) die("temperature sum");
  T_V_FLOAT T = m4[0];
  comm->reduce_float(&T, 1);
  const T_INT dof = 3 * s.N - 3;
  T *= s.mvv2e / (1.0 * dof * s.boltz);
  if (emd_velocity_scale(s.ctx, s.v, n, sqrt(temperature_target / T))) die("scale");
}

void Input::create_lattice(Comm *comm) {
  system->set_mass(system->h_mass);
  system->domain_x = lattice_constant * lattice_nx;
  system->domain_y = lattice_constant * lattice_ny;
  system->domain_z = lattice_constant * lattice_nz;
  comm->create_domain_decomposition();

  // one atom type: everything below runs on the device, bit-identical (kernels/lattice.cu); more types need the
  // sequential libc rand() stream of the reference for the type draw and keep the host loops
  const bool host_only = getenv("EMD_HOST_LATTICE") && atoi(getenv("EMD_HOST_LATTICE")); // (read at every call: the tests compare both paths)
  if (system->ntypes == 1 && !host_only) { create_lattice_device(comm); return; }

  T_INT n = 0;
  for_each_site(*this, *system, [&](const Site &) { n++; });
  system->N_local = n;
  system->N = n;
  system->grow(n + n / 4 + 16); // head-room for ghosts (the reference ends up with 2n, :497-534)

  HostAtoms h;
  h.resize(n);
  T_INT k = 0;
  for_each_site(*this, *system, [&](const Site &p) {
    h.x[3 * k] = p.x; h.x[3 * k + 1] = p.y; h.x[3 * k + 2] = p.z;
    h.type[k] = rand() % system->ntypes;
    h.id[k] = k + 1;
    k++;
  });

  // global atom count and globally unique ids (:578-584 / :714-721)
  T_INT N_local_offset = n;
  comm->scan_int(&N_local_offset, 1);
  for (T_INT i = 0; i < n; i++) h.id[i] += N_local_offset - n;
  comm->reduce_int(&system->N, 1);
  if (system->do_print) printf("Atoms: %i %i\n", system->N, system->N_local);

  // velocities: uniform in [-0.5,0.5)/sqrt(m) from the per-position stream, then remove the
  // centre-of-mass velocity and rescale to the target temperature (:731-785)
  T_FLOAT total_mass = 0.0, px = 0.0, py = 0.0, pz = 0.0;
  for (T_INT i = 0; i < n; i++) {
    LAMMPS_RandomVelocityGeom random;
    double x[3] = {h.x[3 * i], h.x[3 * i + 1], h.x[3 * i + 2]};
    random.reset(temperature_seed, x);
    const T_FLOAT mass_i = system->h_mass[h.type[i]];
    const T_FLOAT vx = random.uniform() - 0.5, vy = random.uniform() - 0.5, vz = random.uniform() - 0.5;
    h.v[3 * i] = vx / sqrt(mass_i);
    h.v[3 * i + 1] = vy / sqrt(mass_i);
    h.v[3 * i + 2] = vz / sqrt(mass_i);
    h.q[i] = 0.0;
    total_mass += mass_i;
    px += mass_i * h.v[3 * i];
    py += mass_i * h.v[3 * i + 1];
    pz += mass_i * h.v[3 * i + 2];
  }
  comm->reduce_float(&px, 1);
  comm->reduce_float(&py, 1);
  comm->reduce_float(&pz, 1);
  comm->reduce_float(&total_mass, 1);
  const T_FLOAT svx = px / total_mass, svy = py / total_mass, svz = pz / total_mass;
  for (T_INT i = 0; i < n; i++) { h.v[3 * i] -= svx; h.v[3 * i + 1] -= svy; h.v[3 * i + 2] -= svz; }

  // temperature of the un-scaled velocities: summed on the host in atom order, which is what
  // the reference's 1-thread parallel_reduce does (property_temperature.cpp:49)
  T_V_FLOAT T = 0.0;
  for (T_INT i = 0; i < n; i++)
    T += (h.v[3 * i] * h.v[3 * i] + h.v[3 * i + 1] * h.v[3 * i + 1] + h.v[3 * i + 2] * h.v[3 * i + 2]) * system->h_mass[h.type[i]];
  comm->reduce_float(&T, 1);
  const T_INT dof = 3 * system->N - 3;
  T *= system->mvv2e / (1.0 * dof * system->boltz);
  const T_V_FLOAT T_init_scale = sqrt(temperature_target / T);
  for (T_INT i = 0; i < 3 * n; i++) h.v[i] *= T_init_scale;

  std::fill(h.f.begin(), h.f.end(), 0.0);
  system->upload(h, n, true);
}
