// ForceSNAP host class: file parsing of src/force_types/force_snap_neigh_impl.h:227-336 (init_coeff)
// and :340-587 (read_files), then the CUDA path.  No arithmetic of the force lives here.
#include "force_snap_neigh.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static void snap_abort(const char *msg) { // the reference uses Kokkos::abort(msg)
  fprintf(stderr, "%s\n", msg);
  emd_host_exit(1);
}

ForceSNAP::ForceSNAP(char **args, System *system, bool half_neigh_)
    : Force(args, system, half_neigh_), sys(system), snap(nullptr), nelements(0), ncoeffall(0), ncoeff(0), rcutfac(0.0), rfac0(0.99363),
      rmin0(0.0), rcutmax(0.0), twojmax(0), diagonalstyle(3), switchflag(1), bzeroflag(1), quadraticflag(0) {}

ForceSNAP::~ForceSNAP() { if (snap) emd_snap_destroy(snap); }

// coefficient files are opened relative to the working directory like the reference (:348,:519); as a
// convenience the directory of the input deck is tried next
static FILE *open_coeff_file(const System *sys, const char *name) {
  FILE *fp = fopen(name, "r");
  if (!fp && !sys->input_dir.empty()) fp = fopen((sys->input_dir + "/" + name).c_str(), "r");
  return fp;
}

void ForceSNAP::read_files(const char *coefffilename, const char *paramfilename) {
  const int MAXLINE = 1024;
  const char *sep = "' \t\n\r\f";
  char line[MAXLINE], *ptr;
  FILE *fpcoeff = open_coeff_file(sys, coefffilename);
  if (fpcoeff == NULL) { snprintf(line, sizeof line, "Cannot open SNAP coefficient file %s", coefffilename); snap_abort(line); }
  // first line that has no '#' anywhere and does not start with a newline: "nelemfile ncoeffall" (:356-385)
  int nwords = 0;
  while (nwords == 0) {
    if (fgets(line, MAXLINE, fpcoeff) == NULL) { fclose(fpcoeff); break; }
    if ((ptr = strchr(line, '#'))) *ptr = '\0';
    else if (line[0] != 10) nwords = 2;
  }
  if (nwords != 2) snap_abort("Incorrect format in SNAP coefficient file");
  char *w0 = strtok(line, sep), *w1 = strtok(NULL, sep);
  if (!w0 || !w1) snap_abort("Incorrect format in SNAP coefficient file");
  const int nelemfile = atoi(w0);
  ncoeffall = atoi(w1);
  radelem.assign(nelements, 0.0);
  wjelem.assign(nelements, 0.0);
  coeffelem.assign((size_t)nelements * ncoeffall, 0.0);
  std::vector<int> found(nelements, 0);
  for (int ielemfile = 0; ielemfile < nelemfile; ielemfile++) {
    if (fgets(line, MAXLINE, fpcoeff) == NULL) snap_abort("Incorrect format in SNAP coefficient file");
    char *elemtmp = strtok(line, sep), *rad = strtok(NULL, sep), *wjw = strtok(NULL, sep);
    if (!elemtmp || !rad || !wjw) snap_abort("Incorrect format in SNAP coefficient file");
    const double radtmp = atof(rad), wjtmp = atof(wjw);
    int ielem;
    for (ielem = 0; ielem < nelements; ielem++)
      if (elements[ielem] == elemtmp) break;
    if (ielem == nelements || found[ielem]) { // not in the element list, or seen before: skip its block (:441-456)
      for (int icoeff = 0; icoeff < ncoeffall; icoeff++) ptr = fgets(line, MAXLINE, fpcoeff);
      continue;
    }
    found[ielem] = 1;
    // the reference deep_copies the scalar into the WHOLE radelem / wjelem views (:459-462)
    for (int e = 0; e < nelements; e++) { radelem[e] = radtmp; wjelem[e] = wjtmp; }
    for (int icoeff = 0; icoeff < ncoeffall; icoeff++) {
      if (fgets(line, MAXLINE, fpcoeff) == NULL) snap_abort("Incorrect format in SNAP coefficient file");
      char *w = strtok(line, sep);
      coeffelem[(size_t)ielem * ncoeffall + icoeff] = w ? atof(w) : 0.0;
    }
  }
  fclose(fpcoeff);

  int rcutfacflag = 0, twojmaxflag = 0; // :505-515
  rfac0 = 0.99363; rmin0 = 0.0; diagonalstyle = 3; switchflag = 1; bzeroflag = 1; quadraticflag = 0;
  FILE *fpparam = open_coeff_file(sys, paramfilename);
  if (fpparam == NULL) { snprintf(line, sizeof line, "Cannot open SNAP parameter file %s", paramfilename); snap_abort(line); }
  while (fgets(line, MAXLINE, fpparam) != NULL) {
    if ((ptr = strchr(line, '#'))) { *ptr = '\0'; continue; }
    if (line[0] == 10) continue;
    char *keywd = strtok(line, sep), *keyval = strtok(NULL, sep);
    if (!keywd || !keyval) snap_abort("Incorrect format in SNAP parameter file");
    if (strcmp(keywd, "rcutfac") == 0) { rcutfac = atof(keyval); rcutfacflag = 1; }
    else if (strcmp(keywd, "twojmax") == 0) { twojmax = atoi(keyval); twojmaxflag = 1; }
    else if (strcmp(keywd, "rfac0") == 0) rfac0 = atof(keyval);
    else if (strcmp(keywd, "rmin0") == 0) rmin0 = atof(keyval);
    else if (strcmp(keywd, "diagonalstyle") == 0) diagonalstyle = atoi(keyval);
    else if (strcmp(keywd, "switchflag") == 0) switchflag = atoi(keyval);
    else if (strcmp(keywd, "bzeroflag") == 0) bzeroflag = atoi(keyval);
    else if (strcmp(keywd, "quadraticflag") == 0) quadraticflag = atoi(keyval);
    else snap_abort("Incorrect SNAP parameter file");
  }
  fclose(fpparam);
  if (rcutfacflag == 0 || twojmaxflag == 0) snap_abort("Incorrect SNAP parameter file");
}

// args = words of the pair_coeff line: pair_coeff * * <coeff file> <elements...> <param file> <element per type...>
void ForceSNAP::init_coeff(int narg, char **arg) {
  if (narg < 7) snap_abort("SNAP 1: Incorrect args for pair coefficients");
  nelements = narg - 5 - sys->ntypes;
  if (nelements < 1) snap_abort("SNAP 2: Incorrect args for pair coefficients");
  if (strcmp(arg[1], "*") != 0 || strcmp(arg[2], "*") != 0) snap_abort("A Incorrect args for pair coefficients");
  elements.clear();
  for (int i = 0; i < nelements; i++) elements.push_back(arg[4 + i]);
  read_files(arg[3], arg[4 + nelements]);
  if (!quadraticflag) ncoeff = ncoeffall - 1;
  else snap_abort("ForceSNAP: quadratic SNAP is parsed but not evaluated by ExaMiniMD's compute; refusing to run it");
  if (diagonalstyle != 3) snap_abort("ForceSNAP: only diagonalstyle 3 has index lists in ExaMiniMD (sna_impl.hpp:86-132)");
  map.assign(sys->ntypes + 1, 0);
  for (int i = 1; i <= sys->ntypes; i++) {
    const char *elemname = arg[5 + nelements + i - 1];
    int jelem;
    for (jelem = 0; jelem < nelements; jelem++)
      if (elements[jelem] == elemname) break;
    if (jelem < nelements) map[i] = jelem;
    else if (strcmp(elemname, "NULL") == 0) map[i] = -1;
    else snap_abort("Incorrect args for pair coefficients");
  }
  rcutmax = 0.0; // :321-327
  for (int ielem = 0; ielem < nelements; ielem++) rcutmax = std::fmax(2.0 * radelem[ielem] * rcutfac, rcutmax);

  emd_snap_params p;
  memset(&p, 0, sizeof p);
  p.twojmax = twojmax; p.switchflag = switchflag; p.ntypes = sys->ntypes; p.nelements = nelements; p.ncoeffall = ncoeffall;
  p.rcutfac = rcutfac; p.rfac0 = rfac0; p.rmin0 = rmin0; p.wself = 1.0; // sna_impl.hpp:30
  if (sys->ntypes > 12) snap_abort("ForceSNAP: more than 12 atom types");
  // the kernel reads map[type] with the 0-based atom type (:597) although map was filled from index 1
  for (int t = 0; t < sys->ntypes; t++) {
    if (map[t] < 0) snap_abort("ForceSNAP: NULL element mapping is not usable (the reference would index out of bounds)");
    p.elem_of_type[t] = map[t];
  }
  p.radelem = radelem.data(); p.wjelem = wjelem.data(); p.coeffelem = coeffelem.data();
  if (snap) { emd_snap_destroy(snap); snap = nullptr; }
  if (emd_snap_create(&snap, &p)) { // includes the reference's "Incorrect SNAP parameter file" check (ncoeff vs twojmax, :315-318)
    fprintf(stderr, "ForceSNAP: %s\n", emd_last_error());
    emd_host_exit(1);
  }
}

// src/force_types/force_snap_neigh_impl.h:159-208
void ForceSNAP::compute(System *system, Binning *, Neighbor *neighbor) {
  if (comm_newton == false) snap_abort("ForceSNAP requires 'newton on'");
  if (!snap) snap_abort("ForceSNAP: pair_coeff missing");
  const emd_neigh_list l = neighbor->list_view();
  if (emd_force_snap_compute(system->ctx, snap, system->x, system->type, system->f, system->N_local,
                             system->N_local + system->N_ghost, &l)) {
    fprintf(stderr, "ForceSNAP: compute: %s\n", emd_last_error());
    emd_host_exit(1);
  }
}

T_F_FLOAT ForceSNAP::compute_energy(System *system, Binning *, Neighbor *neighbor) {
  const char *e = getenv("EMD_SNAP_ENERGY");
  if (!e || !atoi(e) || !snap) return 0.0; // src/force.h:54
  const emd_neigh_list l = neighbor->list_view();
  double pe = 0.0;
  if (emd_force_snap_energy(system->ctx, snap, system->x, system->type, system->N_local, &l, bzeroflag, &pe)) {
    fprintf(stderr, "ForceSNAP: compute_energy: %s\n", emd_last_error());
    emd_host_exit(1);
  }
  return pe;
}

const char *ForceSNAP::name() { return "ForceSNAP"; }
