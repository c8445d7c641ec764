/*
 * oracle_snap.c -- CPU restatement of the reference's SNAP force (TEST INFRASTRUCTURE ONLY, see
 * oracle.h).  Follows, operation for operation, the reference's CPU path (team size 1, vector
 * length collapsed to 1: every Kokkos nested loop is a plain serial loop):
 *   ForceSNAP::init_coeff / read_files   src/force_types/force_snap_neigh_impl.h:227-336, 340-587
 *   ForceSNAP::operator() (one atom)      src/force_types/force_snap_neigh_impl.h:589-725
 *   SNA::build_indexlist                  src/force_types/sna_impl.hpp:86-132
 *   SNA::compute_ui / compute_uarray / add_uarraytot    sna_impl.hpp:152-190, 641-720, 612-635
 *   SNA::compute_zi                       sna_impl.hpp:196-283
 *   SNA::compute_duidrj / compute_duarray sna_impl.hpp:290-321, 728-893
 *   SNA::compute_dbidrj / copy_dbi2dbvec  sna_impl.hpp:329-539, 545-570
 *   SNA::init_clebsch_gordan / init_rootpqarray / compute_ncoeff / compute_sfac / compute_dsfac
 *                                         sna_impl.hpp:991-1061, 1065-1096, 1098-1128
 * Pinned against the unmodified reference compiled over the Kokkos stand-in (oracle/_ref): the
 * forces of input/snap/in.snap.W and in.snap.Ta06A agree bit for bit (tests/test_oracle_vs_reference.py).
 * Arrays keep the reference's index order: u(j,ma,mb), du(j,mb,ma,k), z(j1,j2,j,mb,ma), all of
 * extent jdim = twojmax+1 per index.
 */
#include "oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const double ORC_PI = 3.14159265358979323846; /* sna_impl.hpp:23 */

typedef struct { int j1, j2, j; } snap_triple;

struct orc_force_snap {
  int ntypes;
  int nelements, ncoeffall, ncoeff;
  char elements[8][ORC_WORD];
  int map[16];          /* atom type (1-based) -> element, force_snap_neigh_impl.h:293-306 */
  double *radelem, *wjelem, *coeffelem; /* [nelements], [nelements][ncoeffall] */
  double rcutfac, rfac0, rmin0, rcutmax;
  int twojmax, diagonalstyle, switchflag, bzeroflag, quadraticflag;
  double wself;
  /* SNA tables */
  int jdim;
  snap_triple *idxj, *idxj_full;
  int idxj_max, idxj_full_max;
  double *cgarray;  /* [jdim]^5 */
  double *rootpq;   /* [jdim+1][jdim+1] */
  /* per-atom work arrays */
  double *utot_r, *utot_i;       /* [jdim]^3 */
  double *z_r, *z_i;             /* [jdim]^5 */
  double *u_r, *u_i;             /* [jdim]^3 */
  double *du_r, *du_i;           /* [jdim]^3 x 3 */
  double *dbarray;               /* [jdim]^3 x 3 */
  double *dbvec;                 /* [ncoeff][3] */
  /* per-atom neighbor scratch */
  int nmax;
  double *rij, *wj, *rcutij;
  int *inside;
};

#define U3(f, j, a, b) ((((size_t)(j)) * (f)->jdim + (a)) * (f)->jdim + (b))
#define Z5(f, j1, j2, j, a, b) ((((((size_t)(j1)) * (f)->jdim + (j2)) * (f)->jdim + (j)) * (f)->jdim + (a)) * (f)->jdim + (b))

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* sna_impl.hpp:962-968 */
static double factorial(int n) {
  double result = 1.0;
  for (int i = 1; i <= n; i++) result *= 1.0 * i;
  return result;
}

/* sna_impl.hpp:975-981 */
static double deltacg(int j1, int j2, int j) {
  double sfaccg = factorial((j1 + j2 + j) / 2 + 1);
  return sqrt(factorial((j1 + j2 - j) / 2) * factorial((j1 - j2 + j) / 2) * factorial((-j1 + j2 + j) / 2) / sfaccg);
}

/* sna_impl.hpp:1065-1096 (diagonalstyle 0..3) */
static int compute_ncoeff(int twojmax, int diagonalstyle) {
  int ncount = 0;
  for (int j1 = 0; j1 <= twojmax; j1++) {
    if (diagonalstyle == 0) {
      for (int j2 = 0; j2 <= j1; j2++)
        for (int j = abs(j1 - j2); j <= imin(twojmax, j1 + j2); j += 2) ncount++;
    } else if (diagonalstyle == 1) {
      int j2 = j1;
      for (int j = abs(j1 - j2); j <= imin(twojmax, j1 + j2); j += 2) ncount++;
    } else if (diagonalstyle == 2) {
      ncount++;
    } else if (diagonalstyle == 3) {
      for (int j2 = 0; j2 <= j1; j2++)
        for (int j = abs(j1 - j2); j <= imin(twojmax, j1 + j2); j += 2)
          if (j >= j1) ncount++;
    }
  }
  return ncount;
}

/* sna_impl.hpp:86-132: only diagonalstyle 3 builds lists (others leave them empty) */
static void build_indexlist(orc_force_snap *f) {
  f->idxj_max = f->idxj_full_max = 0;
  if (f->diagonalstyle != 3) return;
  int c = 0, cf = 0;
  for (int j1 = 0; j1 <= f->twojmax; j1++)
    for (int j2 = 0; j2 <= j1; j2++)
      for (int j = abs(j1 - j2); j <= imin(f->twojmax, j1 + j2); j += 2) { if (j >= j1) c++; cf++; }
  f->idxj = (snap_triple *)malloc(sizeof(snap_triple) * (size_t)imax(c, 1));
  f->idxj_full = (snap_triple *)malloc(sizeof(snap_triple) * (size_t)imax(cf, 1));
  f->idxj_max = c; f->idxj_full_max = cf;
  c = cf = 0;
  for (int j1 = 0; j1 <= f->twojmax; j1++)
    for (int j2 = 0; j2 <= j1; j2++)
      for (int j = abs(j1 - j2); j <= imin(f->twojmax, j1 + j2); j += 2) {
        if (j >= j1) { f->idxj[c].j1 = j1; f->idxj[c].j2 = j2; f->idxj[c].j = j; c++; }
        f->idxj_full[cf].j1 = j1; f->idxj_full[cf].j2 = j2; f->idxj_full[cf].j = j; cf++;
      }
}

/* sna_impl.hpp:991-1046 */
static void init_clebsch_gordan(orc_force_snap *f) {
  const int twojmax = f->twojmax;
  for (int j1 = 0; j1 <= twojmax; j1++)
    for (int j2 = 0; j2 <= twojmax; j2++)
      for (int j = abs(j1 - j2); j <= imin(twojmax, j1 + j2); j += 2)
        for (int m1 = 0; m1 <= j1; m1 += 1) {
          const int aa2 = 2 * m1 - j1;
          for (int m2 = 0; m2 <= j2; m2 += 1) {
            const int bb2 = 2 * m2 - j2;
            const int m = (aa2 + bb2 + j) / 2;
            if (m < 0 || m > j) continue;
            double sum = 0.0;
            for (int z = imax(0, imax(-(j - j2 + aa2) / 2, -(j - j1 - bb2) / 2));
                 z <= imin((j1 + j2 - j) / 2, imin((j1 - aa2) / 2, (j2 + bb2) / 2)); z++) {
              const int ifac = z % 2 ? -1 : 1;
              sum += ifac / (factorial(z) * factorial((j1 + j2 - j) / 2 - z) * factorial((j1 - aa2) / 2 - z) *
                             factorial((j2 + bb2) / 2 - z) * factorial((j - j2 + aa2) / 2 + z) *
                             factorial((j - j1 - bb2) / 2 + z));
            }
            const int cc2 = 2 * m - j;
            const double dcg = deltacg(j1, j2, j);
            const double sfaccg = sqrt(factorial((j1 + aa2) / 2) * factorial((j1 - aa2) / 2) * factorial((j2 + bb2) / 2) *
                                       factorial((j2 - bb2) / 2) * factorial((j + cc2) / 2) * factorial((j - cc2) / 2) * (j + 1));
            f->cgarray[Z5(f, j1, j2, j, m1, m2)] = sum * dcg * sfaccg;
          }
        }
}

/* sna_impl.hpp:1053-1061 */
static void init_rootpqarray(orc_force_snap *f) {
  for (int p = 1; p <= f->twojmax; p++)
    for (int q = 1; q <= f->twojmax; q++) f->rootpq[p * (f->jdim + 1) + q] = sqrt((double)p / q);
}
#define ROOTPQ(f, p, q) ((f)->rootpq[(p) * ((f)->jdim + 1) + (q)])

/* sna_impl.hpp:1098-1111 */
static double compute_sfac(const orc_force_snap *f, double r, double rcut) {
  if (f->switchflag == 0) return 1.0;
  if (f->switchflag == 1) {
    if (r <= f->rmin0) return 1.0;
    else if (r > rcut) return 0.0;
    else {
      double rcutfac = ORC_PI / (rcut - f->rmin0);
      return 0.5 * (cos((r - f->rmin0) * rcutfac) + 1.0);
    }
  }
  return 0.0;
}

/* sna_impl.hpp:1115-1128 */
static double compute_dsfac(const orc_force_snap *f, double r, double rcut) {
  if (f->switchflag == 0) return 0.0;
  if (f->switchflag == 1) {
    if (r <= f->rmin0) return 0.0;
    else if (r > rcut) return 0.0;
    else {
      double rcutfac = ORC_PI / (rcut - f->rmin0);
      return -0.5 * sin((r - f->rmin0) * rcutfac) * rcutfac;
    }
  }
  return 0.0;
}

orc_force_snap *orc_force_snap_create(int ntypes) {
  orc_force_snap *f = (orc_force_snap *)calloc(1, sizeof(orc_force_snap));
  f->ntypes = ntypes;
  return f;
}

void orc_force_snap_destroy(orc_force_snap *f) {
  if (!f) return;
  free(f->radelem); free(f->wjelem); free(f->coeffelem); free(f->idxj); free(f->idxj_full); free(f->cgarray); free(f->rootpq);
  free(f->utot_r); free(f->utot_i); free(f->z_r); free(f->z_i); free(f->u_r); free(f->u_i); free(f->du_r); free(f->du_i);
  free(f->dbarray); free(f->dbvec); free(f->rij); free(f->wj); free(f->rcutij); free(f->inside);
  free(f);
}

static FILE *open_in_dir(const char *dir, const char *name) {
  char path[1200];
  if (dir && dir[0]) snprintf(path, sizeof path, "%s/%s", dir, name);
  else snprintf(path, sizeof path, "%s", name);
  return fopen(path, "r");
}

/* force_snap_neigh_impl.h:340-587 */
static int read_files(orc_force_snap *f, const char *dir, const char *coefffilename, const char *paramfilename) {
  FILE *fpcoeff = open_in_dir(dir, coefffilename);
  if (!fpcoeff) { fprintf(stderr, "oracle: cannot open SNAP coefficient file %s\n", coefffilename); return -1; }
  char line[1024], *ptr;
  int nwords = 0;
  while (nwords == 0) { /* :358-377: a line holding '#' anywhere is dropped whole; a line starting with \n is blank */
    if (!fgets(line, sizeof line, fpcoeff)) { fclose(fpcoeff); return -2; }
    if ((ptr = strchr(line, '#'))) *ptr = '\0';
    else if (line[0] != 10) nwords = 2;
  }
  const char *sep = "' \t\n\r\f";
  char *w0 = strtok(line, sep), *w1 = strtok(NULL, sep);
  if (!w0 || !w1) { fclose(fpcoeff); return -2; }
  const int nelemfile = atoi(w0);
  f->ncoeffall = atoi(w1);
  f->radelem = (double *)calloc((size_t)f->nelements, sizeof(double));
  f->wjelem = (double *)calloc((size_t)f->nelements, sizeof(double));
  f->coeffelem = (double *)calloc((size_t)f->nelements * f->ncoeffall, sizeof(double));
  int found[8] = {0};
  for (int ielemfile = 0; ielemfile < nelemfile; ielemfile++) {
    if (!fgets(line, sizeof line, fpcoeff)) { fclose(fpcoeff); return -2; }
    char *elemtmp = strtok(line, sep), *rad = strtok(NULL, sep), *wjw = strtok(NULL, sep);
    const double radtmp = atof(rad), wjtmp = atof(wjw);
    int ielem;
    for (ielem = 0; ielem < f->nelements; ielem++)
      if (strcmp(elemtmp, f->elements[ielem]) == 0) break;
    if (ielem == f->nelements || found[ielem]) { /* :441-456 */
      for (int icoeff = 0; icoeff < f->ncoeffall; icoeff++) ptr = fgets(line, sizeof line, fpcoeff);
      continue;
    }
    found[ielem] = 1;
    /* :459-462: deep_copy(radelem, radtmp) assigns EVERY element's radius and weight */
    for (int e = 0; e < f->nelements; e++) { f->radelem[e] = radtmp; f->wjelem[e] = wjtmp; }
    for (int icoeff = 0; icoeff < f->ncoeffall; icoeff++) {
      if (!fgets(line, sizeof line, fpcoeff)) { fclose(fpcoeff); return -2; }
      char *w = strtok(line, sep);
      f->coeffelem[(size_t)ielem * f->ncoeffall + icoeff] = atof(w);
    }
  }
  fclose(fpcoeff);

  int rcutfacflag = 0, twojmaxflag = 0; /* :505-515 */
  f->rfac0 = 0.99363; f->rmin0 = 0.0; f->diagonalstyle = 3; f->switchflag = 1; f->bzeroflag = 1; f->quadraticflag = 0;
  FILE *fpparam = open_in_dir(dir, paramfilename);
  if (!fpparam) { fprintf(stderr, "oracle: cannot open SNAP parameter file %s\n", paramfilename); return -1; }
  while (fgets(line, sizeof line, fpparam)) {
    if ((ptr = strchr(line, '#'))) { *ptr = '\0'; continue; }
    if (line[0] == 10) continue;
    char *keywd = strtok(line, sep), *keyval = strtok(NULL, sep);
    if (!keywd || !keyval) continue;
    if (strcmp(keywd, "rcutfac") == 0) { f->rcutfac = atof(keyval); rcutfacflag = 1; }
    else if (strcmp(keywd, "twojmax") == 0) { f->twojmax = atoi(keyval); twojmaxflag = 1; }
    else if (strcmp(keywd, "rfac0") == 0) f->rfac0 = atof(keyval);
    else if (strcmp(keywd, "rmin0") == 0) f->rmin0 = atof(keyval);
    else if (strcmp(keywd, "diagonalstyle") == 0) f->diagonalstyle = atoi(keyval);
    else if (strcmp(keywd, "switchflag") == 0) f->switchflag = atoi(keyval);
    else if (strcmp(keywd, "bzeroflag") == 0) f->bzeroflag = atoi(keyval);
    else if (strcmp(keywd, "quadraticflag") == 0) f->quadraticflag = atoi(keyval);
    else { fclose(fpparam); return -3; }
  }
  fclose(fpparam);
  if (!rcutfacflag || !twojmaxflag) return -3;
  return 0;
}

/* force_snap_neigh_impl.h:227-336; args = the words of the pair_coeff line, args[0] = "pair_coeff" */
int orc_force_snap_init_coeff(orc_force_snap *f, int narg, char args[][ORC_WORD], const char *dir) {
  if (narg < 7) return -1;
  f->nelements = narg - 5 - f->ntypes;
  if (f->nelements < 1 || f->nelements > 8) return -1;
  if (strcmp(args[1], "*") != 0 || strcmp(args[2], "*") != 0) return -1;
  for (int i = 0; i < f->nelements; i++) { strncpy(f->elements[i], args[4 + i], ORC_WORD - 1); f->elements[i][ORC_WORD - 1] = 0; }
  if (read_files(f, dir, args[3], args[4 + f->nelements])) return -2;
  if (!f->quadraticflag) f->ncoeff = f->ncoeffall - 1;
  else return -4; /* quadratic terms are parsed but never evaluated by the reference's compute */
  for (int i = 1; i <= f->ntypes; i++) {
    const char *elemname = args[5 + f->nelements + i - 1];
    int jelem;
    for (jelem = 0; jelem < f->nelements; jelem++)
      if (strcmp(elemname, f->elements[jelem]) == 0) break;
    if (jelem < f->nelements) f->map[i] = jelem;
    else if (strcmp(elemname, "NULL") == 0) f->map[i] = -1;
    else return -1;
  }
  /* SNA::SNA + init, sna_impl.hpp:25-53,136-140 */
  f->wself = 1.0;
  f->jdim = f->twojmax + 1;
  const size_t j3 = (size_t)f->jdim * f->jdim * f->jdim, j5 = j3 * f->jdim * f->jdim;
  if (compute_ncoeff(f->twojmax, f->diagonalstyle) != f->ncoeff) return -5; /* :315-318 */
  build_indexlist(f);
  f->cgarray = (double *)calloc(j5, sizeof(double));
  f->rootpq = (double *)calloc((size_t)(f->jdim + 1) * (f->jdim + 1), sizeof(double));
  init_clebsch_gordan(f);
  init_rootpqarray(f);
  f->utot_r = (double *)calloc(j3, sizeof(double)); f->utot_i = (double *)calloc(j3, sizeof(double));
  f->z_r = (double *)calloc(j5, sizeof(double)); f->z_i = (double *)calloc(j5, sizeof(double));
  f->u_r = (double *)calloc(j3, sizeof(double)); f->u_i = (double *)calloc(j3, sizeof(double));
  f->du_r = (double *)calloc(3 * j3, sizeof(double)); f->du_i = (double *)calloc(3 * j3, sizeof(double));
  f->dbarray = (double *)calloc(3 * j3, sizeof(double));
  f->dbvec = (double *)calloc(3 * (size_t)imax(f->ncoeff, 1), sizeof(double));
  f->rcutmax = 0.0; /* :321-327 */
  for (int ielem = 0; ielem < f->nelements; ielem++) {
    const double c = 2.0 * f->radelem[ielem] * f->rcutfac;
    if (c > f->rcutmax) f->rcutmax = c;
  }
  return 0;
}

int orc_force_snap_ncoeff(const orc_force_snap *f) { return f->ncoeff; }
double orc_force_snap_rcutmax(const orc_force_snap *f) { return f->rcutmax; }

int orc_snap_tables(const orc_force_snap *f, int *twojmax, int *idxj_max, int *idxj_full_max, const double **cgarray,
                    const double **rootpq, const double **coeffelem) {
  if (twojmax) *twojmax = f->twojmax;
  if (idxj_max) *idxj_max = f->idxj_max;
  if (idxj_full_max) *idxj_full_max = f->idxj_full_max;
  if (cgarray) *cgarray = f->cgarray;
  if (rootpq) *rootpq = f->rootpq;
  if (coeffelem) *coeffelem = f->coeffelem;
  return 0;
}

static void grow_nmax(orc_force_snap *f, int n) {
  if (n <= f->nmax) return;
  f->nmax = n + 16;
  f->rij = (double *)realloc(f->rij, sizeof(double) * 3 * (size_t)f->nmax);
  f->wj = (double *)realloc(f->wj, sizeof(double) * (size_t)f->nmax);
  f->rcutij = (double *)realloc(f->rcutij, sizeof(double) * (size_t)f->nmax);
  f->inside = (int *)realloc(f->inside, sizeof(int) * (size_t)f->nmax);
}

/* sna_impl.hpp:641-720 */
static void compute_uarray(orc_force_snap *f, double x, double y, double z, double z0, double r) {
  const double r0inv = 1.0 / sqrt(r * r + z0 * z0);
  const double a_r = r0inv * z0, a_i = -r0inv * z, b_r = r0inv * y, b_i = -r0inv * x;
  double *u_r = f->u_r, *u_i = f->u_i;
  u_r[U3(f, 0, 0, 0)] = 1.0;
  u_i[U3(f, 0, 0, 0)] = 0.0;
  for (int j = 1; j <= f->twojmax; j++) {
    for (int mb = 0; mb < (j + 2) / 2; mb++) {
      u_r[U3(f, j, 0, mb)] = 0.0;
      u_i[U3(f, j, 0, mb)] = 0.0;
      for (int ma = 0; ma < j; ma++) {
        double rootpq = ROOTPQ(f, j - ma, j - mb);
        u_r[U3(f, j, ma, mb)] += rootpq * (a_r * u_r[U3(f, j - 1, ma, mb)] + a_i * u_i[U3(f, j - 1, ma, mb)]);
        u_i[U3(f, j, ma, mb)] += rootpq * (a_r * u_i[U3(f, j - 1, ma, mb)] - a_i * u_r[U3(f, j - 1, ma, mb)]);
        rootpq = ROOTPQ(f, ma + 1, j - mb);
        u_r[U3(f, j, ma + 1, mb)] = -rootpq * (b_r * u_r[U3(f, j - 1, ma, mb)] + b_i * u_i[U3(f, j - 1, ma, mb)]);
        u_i[U3(f, j, ma + 1, mb)] = -rootpq * (b_r * u_i[U3(f, j - 1, ma, mb)] - b_i * u_r[U3(f, j - 1, ma, mb)]);
      }
    }
    /* inversion symmetry VMK 4.4(2): u[j-ma][j-mb] = (-1)^(ma-mb) conj(u[ma][mb]) */
    for (int mb = 0; mb < (j + 2) / 2; mb++) {
      int mbpar = (mb) % 2 == 0 ? 1 : -1;
      int mapar = -mbpar;
      for (int ma = 0; ma <= j; ma++) {
        mapar = -mapar;
        if (mapar == 1) {
          u_r[U3(f, j, j - ma, j - mb)] = u_r[U3(f, j, ma, mb)];
          u_i[U3(f, j, j - ma, j - mb)] = -u_i[U3(f, j, ma, mb)];
        } else {
          u_r[U3(f, j, j - ma, j - mb)] = -u_r[U3(f, j, ma, mb)];
          u_i[U3(f, j, j - ma, j - mb)] = u_i[U3(f, j, ma, mb)];
        }
      }
    }
  }
}

/* sna_impl.hpp:152-190 (+ zero :572-590, addself :594-605, add :612-635) */
static void compute_ui(orc_force_snap *f, int jnum) {
  const size_t j3 = (size_t)f->jdim * f->jdim * f->jdim;
  memset(f->utot_r, 0, sizeof(double) * j3);
  memset(f->utot_i, 0, sizeof(double) * j3);
  for (int j = 0; j <= f->twojmax; j++)
    for (int ma = 0; ma <= j; ma++) { f->utot_r[U3(f, j, ma, ma)] = f->wself; f->utot_i[U3(f, j, ma, ma)] = 0.0; }
  for (int j = 0; j < jnum; j++) {
    const double x = f->rij[3 * j], y = f->rij[3 * j + 1], z = f->rij[3 * j + 2];
    const double rsq = x * x + y * y + z * z;
    const double r = sqrt(rsq);
    const double theta0 = (r - f->rmin0) * f->rfac0 * ORC_PI / (f->rcutij[j] - f->rmin0);
    const double z0 = r / tan(theta0);
    compute_uarray(f, x, y, z, z0, r);
    const double sfac = compute_sfac(f, r, f->rcutij[j]) * f->wj[j];
    /* the reference walks the whole [jdim]^3 span; slots with ma or mb > j are never read */
    for (int jj = 0; jj <= f->twojmax; jj++)
      for (int ma = 0; ma <= jj; ma++)
        for (int mb = 0; mb <= jj; mb++) {
          f->utot_r[U3(f, jj, ma, mb)] += sfac * f->u_r[U3(f, jj, ma, mb)];
          f->utot_i[U3(f, jj, ma, mb)] += sfac * f->u_i[U3(f, jj, ma, mb)];
        }
  }
}

/* sna_impl.hpp:196-283 */
static void compute_zi(orc_force_snap *f) {
  for (int idx = 0; idx < f->idxj_full_max; idx++) {
    const int j1 = f->idxj_full[idx].j1, j2 = f->idxj_full[idx].j2, j = f->idxj_full[idx].j;
    const int bound = (j + 2) / 2;
    for (int mbma = 0; mbma < (j + 1) * bound; mbma++) {
      const int ma = mbma % (j + 1);
      const int mb = mbma / (j + 1);
      double z_r = 0.0, z_i = 0.0;
      for (int ma1 = imax(0, (2 * ma - j - j2 + j1) / 2); ma1 <= imin(j1, (2 * ma - j + j2 + j1) / 2); ma1++) {
        double sumb1_r = 0.0, sumb1_i = 0.0;
        const int ma2 = (2 * ma - j - (2 * ma1 - j1) + j2) / 2;
        for (int mb1 = imax(0, (2 * mb - j - j2 + j1) / 2); mb1 <= imin(j1, (2 * mb - j + j2 + j1) / 2); mb1++) {
          const int mb2 = (2 * mb - j - (2 * mb1 - j1) + j2) / 2;
          const double cga = f->cgarray[Z5(f, j1, j2, j, mb1, mb2)];
          const double uat1_r = f->utot_r[U3(f, j1, ma1, mb1)], uat1_i = f->utot_i[U3(f, j1, ma1, mb1)];
          const double uat2_r = f->utot_r[U3(f, j2, ma2, mb2)], uat2_i = f->utot_i[U3(f, j2, ma2, mb2)];
          sumb1_r += cga * (uat1_r * uat2_r - uat1_i * uat2_i);
          sumb1_i += cga * (uat1_r * uat2_i + uat1_i * uat2_r);
        }
        const double cga = f->cgarray[Z5(f, j1, j2, j, ma1, ma2)];
        z_r += sumb1_r * cga;
        z_i += sumb1_i * cga;
      }
      f->z_r[Z5(f, j1, j2, j, mb, ma)] = z_r;
      f->z_i[Z5(f, j1, j2, j, mb, ma)] = z_i;
    }
  }
}

#define DU4(f, j, mb, ma, k) (3 * U3(f, j, mb, ma) + (k))

/* sna_impl.hpp:728-893 */
static void compute_duarray(orc_force_snap *f, double x, double y, double z, double z0, double r, double dz0dr, double wj,
                            double rcut) {
  double da_r[3], da_i[3], db_r[3], db_i[3], dz0[3], dr0inv[3];
  const double rinv = 1.0 / r;
  const double ux = x * rinv, uy = y * rinv, uz = z * rinv;
  const double r0inv = 1.0 / sqrt(r * r + z0 * z0);
  const double a_r = z0 * r0inv, a_i = -z * r0inv, b_r = y * r0inv, b_i = -x * r0inv;
  const double dr0invdr = -pow(r0inv, 3.0) * (r + z0 * dz0dr);
  dr0inv[0] = dr0invdr * ux; dr0inv[1] = dr0invdr * uy; dr0inv[2] = dr0invdr * uz;
  dz0[0] = dz0dr * ux; dz0[1] = dz0dr * uy; dz0[2] = dz0dr * uz;
  for (int k = 0; k < 3; k++) { da_r[k] = dz0[k] * r0inv + z0 * dr0inv[k]; da_i[k] = -z * dr0inv[k]; }
  da_i[2] += -r0inv;
  for (int k = 0; k < 3; k++) { db_r[k] = y * dr0inv[k]; db_i[k] = -x * dr0inv[k]; }
  db_i[0] += -r0inv;
  db_r[1] += r0inv;
  double *u_r = f->u_r, *u_i = f->u_i, *du_r = f->du_r, *du_i = f->du_i;
  u_r[U3(f, 0, 0, 0)] = 1.0; u_i[U3(f, 0, 0, 0)] = 0.0;
  for (int k = 0; k < 3; k++) { du_r[DU4(f, 0, 0, 0, k)] = 0.0; du_i[DU4(f, 0, 0, 0, k)] = 0.0; }
  for (int j = 1; j <= f->twojmax; j++) {
    for (int mb = 0; mb < (j + 2) / 2; mb++) {
      u_r[U3(f, j, 0, mb)] = 0.0; u_i[U3(f, j, 0, mb)] = 0.0;
      for (int k = 0; k < 3; k++) { du_r[DU4(f, j, mb, 0, k)] = 0.0; du_i[DU4(f, j, mb, 0, k)] = 0.0; }
      for (int ma = 0; ma < j; ma++) {
        double rootpq = ROOTPQ(f, j - ma, j - mb);
        u_r[U3(f, j, ma, mb)] += rootpq * (a_r * u_r[U3(f, j - 1, ma, mb)] + a_i * u_i[U3(f, j - 1, ma, mb)]);
        u_i[U3(f, j, ma, mb)] += rootpq * (a_r * u_i[U3(f, j - 1, ma, mb)] - a_i * u_r[U3(f, j - 1, ma, mb)]);
        for (int k = 0; k < 3; k++) {
          du_r[DU4(f, j, mb, ma, k)] += rootpq * (da_r[k] * u_r[U3(f, j - 1, ma, mb)] + da_i[k] * u_i[U3(f, j - 1, ma, mb)] +
                                                  a_r * du_r[DU4(f, j - 1, mb, ma, k)] + a_i * du_i[DU4(f, j - 1, mb, ma, k)]);
          du_i[DU4(f, j, mb, ma, k)] += rootpq * (da_r[k] * u_i[U3(f, j - 1, ma, mb)] - da_i[k] * u_r[U3(f, j - 1, ma, mb)] +
                                                  a_r * du_i[DU4(f, j - 1, mb, ma, k)] - a_i * du_r[DU4(f, j - 1, mb, ma, k)]);
        }
        rootpq = ROOTPQ(f, ma + 1, j - mb);
        u_r[U3(f, j, ma + 1, mb)] = -rootpq * (b_r * u_r[U3(f, j - 1, ma, mb)] + b_i * u_i[U3(f, j - 1, ma, mb)]);
        u_i[U3(f, j, ma + 1, mb)] = -rootpq * (b_r * u_i[U3(f, j - 1, ma, mb)] - b_i * u_r[U3(f, j - 1, ma, mb)]);
        for (int k = 0; k < 3; k++) {
          du_r[DU4(f, j, mb, ma + 1, k)] = -rootpq * (db_r[k] * u_r[U3(f, j - 1, ma, mb)] + db_i[k] * u_i[U3(f, j - 1, ma, mb)] +
                                                      b_r * du_r[DU4(f, j - 1, mb, ma, k)] + b_i * du_i[DU4(f, j - 1, mb, ma, k)]);
          du_i[DU4(f, j, mb, ma + 1, k)] = -rootpq * (db_r[k] * u_i[U3(f, j - 1, ma, mb)] - db_i[k] * u_r[U3(f, j - 1, ma, mb)] +
                                                      b_r * du_i[DU4(f, j - 1, mb, ma, k)] - b_i * du_r[DU4(f, j - 1, mb, ma, k)]);
        }
      }
    }
    for (int mb = 0; mb < (j + 2) / 2; mb++) {
      int mbpar = (mb) % 2 == 0 ? 1 : -1;
      int mapar = -mbpar;
      for (int ma = 0; ma <= j; ma++) {
        mapar = -mapar;
        if (mapar == 1) {
          u_r[U3(f, j, j - ma, j - mb)] = u_r[U3(f, j, ma, mb)];
          u_i[U3(f, j, j - ma, j - mb)] = -u_i[U3(f, j, ma, mb)];
          for (int k = 0; k < 3; k++) {
            du_r[DU4(f, j, j - mb, j - ma, k)] = du_r[DU4(f, j, mb, ma, k)];
            du_i[DU4(f, j, j - mb, j - ma, k)] = -du_i[DU4(f, j, mb, ma, k)];
          }
        } else {
          u_r[U3(f, j, j - ma, j - mb)] = -u_r[U3(f, j, ma, mb)];
          u_i[U3(f, j, j - ma, j - mb)] = u_i[U3(f, j, ma, mb)];
          for (int k = 0; k < 3; k++) {
            du_r[DU4(f, j, j - mb, j - ma, k)] = -du_r[DU4(f, j, mb, ma, k)];
            du_i[DU4(f, j, j - mb, j - ma, k)] = du_i[DU4(f, j, mb, ma, k)];
          }
        }
      }
    }
  }
  double sfac = compute_sfac(f, r, rcut);
  double dsfac = compute_dsfac(f, r, rcut);
  sfac *= wj;
  dsfac *= wj;
  const double uhat[3] = {ux, uy, uz};
  for (int j = 0; j <= f->twojmax; j++)
    for (int mb = 0; mb <= j; mb++)
      for (int ma = 0; ma <= j; ma++)
        for (int k = 0; k < 3; k++) {
          du_r[DU4(f, j, mb, ma, k)] = dsfac * u_r[U3(f, j, ma, mb)] * uhat[k] + sfac * du_r[DU4(f, j, mb, ma, k)];
          du_i[DU4(f, j, mb, ma, k)] = dsfac * u_i[U3(f, j, ma, mb)] * uhat[k] + sfac * du_i[DU4(f, j, mb, ma, k)];
        }
}

/* sna_impl.hpp:290-321 */
static void compute_duidrj(orc_force_snap *f, const double *rij, double wj, double rcut) {
  const double x = rij[0], y = rij[1], z = rij[2];
  const double rsq = x * x + y * y + z * z;
  const double r = sqrt(rsq);
  const double rscale0 = f->rfac0 * ORC_PI / (rcut - f->rmin0);
  const double theta0 = (r - f->rmin0) * rscale0;
  const double cs = cos(theta0), sn = sin(theta0);
  const double z0 = r * cs / sn;
  const double dz0dr = z0 / r - (r * rscale0) * (rsq + z0 * z0) / rsq;
  compute_duarray(f, x, y, z, z0, r, dz0dr, wj, rcut);
}

/* one of the three conj(dU) . Z sums of compute_dbidrj: over the half mb < jd/2 plus, for even jd,
 * the middle column up to the diagonal (whose last element counts half), sna_impl.hpp:393-424 */
static void sum_zdu(const orc_force_snap *f, int jd, int za, int zb, int zc, double s[3]) {
  s[0] = s[1] = s[2] = 0.0;
  for (int mb = 0; 2 * mb < jd; mb++)
    for (int ma = 0; ma <= jd; ma++) {
      const double *dudr_r = &f->du_r[DU4(f, jd, mb, ma, 0)], *dudr_i = &f->du_i[DU4(f, jd, mb, ma, 0)];
      const double zr = f->z_r[Z5(f, za, zb, zc, mb, ma)], zi = f->z_i[Z5(f, za, zb, zc, mb, ma)];
      s[0] += (dudr_r[0] * zr + dudr_i[0] * zi);
      s[1] += (dudr_r[1] * zr + dudr_i[1] * zi);
      s[2] += (dudr_r[2] * zr + dudr_i[2] * zi);
    }
  if (jd % 2 == 0) {
    const int mb = jd / 2;
    for (int ma = 0; ma <= mb; ma++) {
      const double *dudr_r = &f->du_r[DU4(f, jd, mb, ma, 0)], *dudr_i = &f->du_i[DU4(f, jd, mb, ma, 0)];
      const double factor = ma == mb ? 0.5 : 1.0;
      const double zr = f->z_r[Z5(f, za, zb, zc, mb, ma)] * factor, zi = f->z_i[Z5(f, za, zb, zc, mb, ma)] * factor;
      s[0] += (dudr_r[0] * zr + dudr_i[0] * zi);
      s[1] += (dudr_r[1] * zr + dudr_i[1] * zi);
      s[2] += (dudr_r[2] * zr + dudr_i[2] * zi);
    }
  }
}

/* sna_impl.hpp:329-539 + copy_dbi2dbvec :545-570 */
static void compute_dbidrj(orc_force_snap *f) {
  for (int JJ = 0; JJ < f->idxj_max; JJ++) {
    const int j1 = f->idxj[JJ].j1, j2 = f->idxj[JJ].j2, j = f->idxj[JJ].j;
    double dbdr[3] = {0.0, 0.0, 0.0}, s[3];
    /* conj(dU(j)) . Z(j1,j2,j), using Z's j1<->j2 symmetry */
    if (j1 >= j2) sum_zdu(f, j, j1, j2, j, s); else sum_zdu(f, j, j2, j1, j, s);
    for (int k = 0; k < 3; k++) dbdr[k] += 2.0 * s[k];
    /* conj(dU(j1)) . Z(j,j2,j1) */
    const double j1fac = (j + 1) / (j1 + 1.0);
    if (j >= j2) sum_zdu(f, j1, j, j2, j1, s); else sum_zdu(f, j1, j2, j, j1, s);
    for (int k = 0; k < 3; k++) dbdr[k] += 2.0 * s[k] * j1fac;
    /* conj(dU(j2)) . Z(j1,j,j2) */
    const double j2fac = (j + 1) / (j2 + 1.0);
    if (j1 >= j) sum_zdu(f, j2, j1, j, j2, s); else sum_zdu(f, j2, j, j1, j2, s);
    for (int k = 0; k < 3; k++) dbdr[k] += 2.0 * s[k] * j2fac;
    for (int k = 0; k < 3; k++) f->dbarray[DU4(f, j1, j2, j, k)] = dbdr[k];
  }
  for (int JJ = 0; JJ < f->idxj_max; JJ++) {
    const int j1 = f->idxj[JJ].j1, j2 = f->idxj[JJ].j2, j = f->idxj[JJ].j;
    for (int k = 0; k < 3; k++) f->dbvec[3 * JJ + k] = f->dbarray[DU4(f, j1, j2, j, k)];
  }
}

/* the per-neighbor force of force_snap_neigh_impl.h:694-711 */
static void neighbor_force(orc_force_snap *f, int jj, const double *coeffi, double fij[3]) {
  compute_duidrj(f, &f->rij[3 * jj], f->wj[jj], f->rcutij[jj]);
  compute_dbidrj(f);
  fij[0] = fij[1] = fij[2] = 0.0;
  for (int k = 1; k <= f->ncoeff; k++) {
    const double bgb = coeffi[k];
    fij[0] += bgb * f->dbvec[3 * (k - 1)];
    fij[1] += bgb * f->dbvec[3 * (k - 1) + 1];
    fij[2] += bgb * f->dbvec[3 * (k - 1) + 2];
  }
  const double dx = f->rij[3 * jj], dy = f->rij[3 * jj + 1], dz = f->rij[3 * jj + 2];
  const double fdivr = -1.5e6 / pow(dx * dx + dy * dy + dz * dz, 7.0); /* :708, the hard-wired ZBL stand-in */
  fij[0] += dx * fdivr; fij[1] += dy * fdivr; fij[2] += dz * fdivr;
}

/* ForceSNAP::compute + operator(), force_snap_neigh_impl.h:159-208, 589-725.  Requires newton on
 * (:163-164) and a full list; accumulates onto s->f (zeroed by the caller like examinimd.cpp:232). */
void orc_force_snap_compute(orc_force_snap *f, orc_system *s, const orc_neighbor *n) {
  const double cutsq = f->rcutmax * f->rcutmax; /* :327-329: one value for every type pair */
  for (int i = 0; i < s->N_local; i++) {
    const double x_i = s->x[3 * i], y_i = s->x[3 * i + 1], z_i = s->x[3 * i + 2];
    const int type_i = s->type[i];
    /* NB the reference indexes map[] with the 0-based atom type although map is filled from 1
     * (:293-306 vs :597); with map zero-initialised every type lands on element 0 */
    const int elem_i = f->map[type_i];
    const double radi = f->radelem[elem_i];
    int num_neighs;
    const int *row = orc_neigh_row(n, i, &num_neighs);
    grow_nmax(f, num_neighs);
    int ninside = 0;
    for (int jj = 0; jj < num_neighs; jj++) {
      const int j = row[jj];
      const double dx = s->x[3 * j] - x_i, dy = s->x[3 * j + 1] - y_i, dz = s->x[3 * j + 2] - z_i;
      const int type_j = s->type[j];
      const double rsq = dx * dx + dy * dy + dz * dz;
      const int elem_j = f->map[type_j];
      if (rsq < cutsq) {
        f->rij[3 * ninside] = dx; f->rij[3 * ninside + 1] = dy; f->rij[3 * ninside + 2] = dz;
        f->inside[ninside] = j;
        f->wj[ninside] = f->wjelem[elem_j];
        f->rcutij[ninside] = (radi + f->radelem[elem_j]) * f->rcutfac;
        ninside++;
      }
    }
    compute_ui(f, ninside);
    compute_zi(f);
    const double *coeffi = f->coeffelem + (size_t)elem_i * f->ncoeffall;
    for (int jj = 0; jj < ninside; jj++) {
      const int j = f->inside[jj];
      double fij[3];
      neighbor_force(f, jj, coeffi, fij);
      s->f[3 * i] += fij[0]; s->f[3 * i + 1] += fij[1]; s->f[3 * i + 2] += fij[2];
      s->f[3 * j] -= fij[0]; s->f[3 * j + 1] -= fij[1]; s->f[3 * j + 2] -= fij[2];
    }
  }
}

/* function-level probe: the bispectrum pieces of ONE atom with the given in-cutoff neighbors
 * (element 0 coefficients).  utot_*: [jdim]^3, dbvec: [n][ncoeff][3], fij: [n][3]; any may be NULL */
void orc_snap_atom(orc_force_snap *f, int ninside, const double *rij, const double *wj, const double *rcutij, double *utot_r,
                   double *utot_i, double *dbvec, double *fij) {
  grow_nmax(f, ninside);
  memcpy(f->rij, rij, sizeof(double) * 3 * (size_t)ninside);
  memcpy(f->wj, wj, sizeof(double) * (size_t)ninside);
  memcpy(f->rcutij, rcutij, sizeof(double) * (size_t)ninside);
  compute_ui(f, ninside);
  const size_t j3 = (size_t)f->jdim * f->jdim * f->jdim;
  if (utot_r) memcpy(utot_r, f->utot_r, sizeof(double) * j3);
  if (utot_i) memcpy(utot_i, f->utot_i, sizeof(double) * j3);
  if (!dbvec && !fij) return;
  compute_zi(f);
  for (int jj = 0; jj < ninside; jj++) {
    double fj[3];
    neighbor_force(f, jj, f->coeffelem, fj);
    if (dbvec) memcpy(dbvec + (size_t)jj * f->ncoeff * 3, f->dbvec, sizeof(double) * 3 * (size_t)f->ncoeff);
    if (fij) memcpy(fij + 3 * jj, fj, sizeof(double) * 3);
  }
}

/* function-level probe on the live system: atom i of s with its row of n -> U_tot (r,i: [jdim]^3, index
 * (j*jdim+ma)*jdim+mb), the in-cutoff neighbor indices and F_ij of each (inside[k], fij[3k..]); returns ninside */
int orc_force_snap_probe(orc_force_snap *f, const orc_system *s, const orc_neighbor *n, int i, double *utot_r, double *utot_i,
                         int *inside, double *fij) {
  const double cutsq = f->rcutmax * f->rcutmax;
  const double x_i = s->x[3 * i], y_i = s->x[3 * i + 1], z_i = s->x[3 * i + 2];
  const int elem_i = f->map[s->type[i]];
  const double radi = f->radelem[elem_i];
  int num_neighs;
  const int *row = orc_neigh_row(n, i, &num_neighs);
  grow_nmax(f, num_neighs);
  int ninside = 0;
  for (int jj = 0; jj < num_neighs; jj++) {
    const int j = row[jj];
    const double dx = s->x[3 * j] - x_i, dy = s->x[3 * j + 1] - y_i, dz = s->x[3 * j + 2] - z_i;
    const double rsq = dx * dx + dy * dy + dz * dz;
    const int elem_j = f->map[s->type[j]];
    if (rsq < cutsq) {
      f->rij[3 * ninside] = dx; f->rij[3 * ninside + 1] = dy; f->rij[3 * ninside + 2] = dz;
      f->inside[ninside] = j;
      f->wj[ninside] = f->wjelem[elem_j];
      f->rcutij[ninside] = (radi + f->radelem[elem_j]) * f->rcutfac;
      ninside++;
    }
  }
  compute_ui(f, ninside);
  const size_t j3 = (size_t)f->jdim * f->jdim * f->jdim;
  if (utot_r) memcpy(utot_r, f->utot_r, sizeof(double) * j3);
  if (utot_i) memcpy(utot_i, f->utot_i, sizeof(double) * j3);
  if (inside) memcpy(inside, f->inside, sizeof(int) * (size_t)ninside);
  if (fij) {
    compute_zi(f);
    const double *coeffi = f->coeffelem + (size_t)elem_i * f->ncoeffall;
    for (int jj = 0; jj < ninside; jj++) neighbor_force(f, jj, coeffi, fij + 3 * jj);
  }
  return ninside;
}
int orc_force_snap_jdim(const orc_force_snap *f) { return f->jdim; }
