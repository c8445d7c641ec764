// capi.cpp -- C session API over the application object, for hosts that are not C++
// (bench.py / tests drive it through ctypes).  Declared in include/emd_b200_app.h.
#include "emd_b200_app.h"
#include "examinimd.h"
#include <cstring>
#include <exception>
#include <string>
#include <vector>

struct emd_app {
  ExaMiniMD *md = nullptr;
  std::vector<std::string> args;
  std::vector<char *> argv;
};

namespace emd { void set_error(const char *fmt, ...); }

// Host-layer fatal errors (emd_host_exit, types.h) become return codes here instead of ending the embedding process.
// The message the call site printed is also kept in emd_last_error() when the failing C-ABI call set one.
template <class F>
static int guarded(const char *where, F &&f) {
  emd_host_throw_on_exit(true);
  try {
    f();
  } catch (const EmdFatal &e) {
    const std::string prev = emd_last_error();
    emd::set_error("%s: fatal error in the host layer (exit code %d)%s%s", where, e.code, prev.empty() ? "" : ": ", prev.c_str());
    return 1;
  } catch (const std::exception &e) {
    emd::set_error("%s: %s", where, e.what());
    return 1;
  }
  return 0;
}

extern "C" {

int emd_app_create(emd_app **out, int argc, const char *const *argv, int device, void *stream) {
  emd_app *a = new emd_app();
  a->args.push_back("ExaMiniMD");
  for (int i = 0; i < argc; i++) a->args.push_back(argv[i]);
  for (auto &s : a->args) a->argv.push_back(const_cast<char *>(s.c_str()));
  const int rc = guarded("emd_app_create", [&] {
    a->md = new ExaMiniMD(device, stream);
    a->md->quiet = true;
    a->md->init((int)a->argv.size(), a->argv.data());
  });
  if (rc) { *out = nullptr; return rc; } // the half-built application is leaked on purpose: its destructors may touch a failed device context
  *out = a;
  return 0;
}

void emd_app_destroy(emd_app *a) {
  if (!a) return;
  delete a->md;
  delete a;
}

emd_ctx *emd_app_ctx(emd_app *a) { return a->md->system->ctx; }

int emd_app_advance(emd_app *a, int nsteps) {
  return guarded("emd_app_advance", [&] { a->md->advance(nsteps); });
}

int emd_app_run(emd_app *a, int nsteps, double *h_last_thermo3) {
  return guarded("emd_app_run", [&] { a->md->run_quiet(nsteps, h_last_thermo3); });
}

int emd_app_advance_timed(emd_app *a, int nsteps, double *h_seconds4) {
  return guarded("emd_app_advance_timed", [&] {
    PhaseTimers tm(a->md->system->ctx);
    for (int k = 0; k < nsteps; k++) a->md->step_once(++a->md->current_step, &tm, k + 1 < nsteps);
    tm.flush();
    h_seconds4[0] = tm.seconds[PhaseTimers::FORCE]; h_seconds4[1] = tm.seconds[PhaseTimers::NEIGH];
    h_seconds4[2] = tm.seconds[PhaseTimers::COMM]; h_seconds4[3] = tm.seconds[PhaseTimers::OTHER];
  });
}

int emd_app_thermo(emd_app *a, double *T, double *PE, double *KE) {
  return guarded("emd_app_thermo", [&] { a->md->thermo(T, PE, KE); });
}

long long emd_app_get(emd_app *a, const char *what) {
  System *s = a->md->system;
  if (!strcmp(what, "N")) return s->N;
  if (!strcmp(what, "N_local")) return s->N_local;
  if (!strcmp(what, "N_ghost")) return s->N_ghost;
  if (!strcmp(what, "N_max")) return s->N_max;
  if (!strcmp(what, "step")) return a->md->current_step;
  if (!strcmp(what, "total_neighs")) return a->md->neighbor ? a->md->neighbor->total_neighs() : 0;
  if (!strcmp(what, "nsteps")) return a->md->input->nsteps;
  if (!strcmp(what, "exchange_rate")) return a->md->input->comm_exchange_rate;
  if (!strcmp(what, "half_neigh")) return a->md->force->half_neigh;
  if (!strcmp(what, "nbinx")) return a->md->binning->nbinx;
  if (!strcmp(what, "nbiny")) return a->md->binning->nbiny;
  if (!strcmp(what, "nbinz")) return a->md->binning->nbinz;
  if (!strcmp(what, "rank")) return a->md->comm->process_rank();
  if (!strcmp(what, "nranks")) return a->md->comm->num_processes();
  return -1;
}

int emd_app_download(emd_app *a, int *id, int *type, double *q, double *x, double *v, double *f) {
  System *s = a->md->system;
  const size_t n = (size_t)s->N_local;
  emd_ctx *c = s->ctx;
  int rc = 0;
  if (id) rc |= emd_memcpy_d2h(c, id, s->id, sizeof(int) * n);
  if (type) rc |= emd_memcpy_d2h(c, type, s->type, sizeof(int) * n);
  if (q) rc |= emd_memcpy_d2h(c, q, s->q, sizeof(double) * n);
  if (x) rc |= emd_memcpy_d2h(c, x, s->x, sizeof(double) * 3 * n);
  if (v) rc |= emd_memcpy_d2h(c, v, s->v, sizeof(double) * 3 * n);
  if (f) rc |= emd_memcpy_d2h(c, f, s->f, sizeof(double) * 3 * n);
  return rc;
}

int emd_app_upload(emd_app *a, const double *x, const double *v, const double *f) {
  System *s = a->md->system;
  const size_t n = (size_t)s->N_local;
  emd_ctx *c = s->ctx;
  int rc = 0;
  if (x) rc |= emd_memcpy_h2d(c, s->x, x, sizeof(double) * 3 * n);
  if (v) rc |= emd_memcpy_h2d(c, s->v, v, sizeof(double) * 3 * n);
  if (f) rc |= emd_memcpy_h2d(c, s->f, f, sizeof(double) * 3 * n);
  return rc;
}

void *emd_app_device_ptr(emd_app *a, const char *what) {
  System *s = a->md->system;
  if (!strcmp(what, "x")) return s->x;
  if (!strcmp(what, "v")) return s->v;
  if (!strcmp(what, "f")) return s->f;
  if (!strcmp(what, "type")) return s->type;
  if (!strcmp(what, "id")) return s->id;
  if (!strcmp(what, "q")) return s->q;
  if (!strcmp(what, "bincount")) return a->md->binning->bincount;
  if (!strcmp(what, "binoffsets")) return a->md->binning->binoffsets;
  if (!strcmp(what, "permute")) return a->md->binning->permute_vector;
  if (!strcmp(what, "tiles")) return a->md->neighbor->tiles();
  if (!strcmp(what, "snap")) { // the emd_snap* of a ForceSNAP (function-level tests), else NULL
    ForceSNAP *fs = dynamic_cast<ForceSNAP *>(a->md->force);
    return fs ? fs->handle() : nullptr;
  }
  const emd_neigh_list l = a->md->neighbor->list_view();
  if (!strcmp(what, "row_map")) return const_cast<int *>(l.d_row_map);
  if (!strcmp(what, "num_neighs")) return const_cast<int *>(l.d_num_neighs);
  if (!strcmp(what, "neighs")) return const_cast<int *>(l.d_neighs);
  return nullptr;
}

int emd_app_neigh_stride(emd_app *a) { return a->md->neighbor->list_view().stride; }

int emd_app_dump_binary(emd_app *a, const char *path, int step) {
  Input *in = a->md->input;
  char *old_path = in->dumpbinary_path;
  int old_rate = in->dumpbinary_rate;
  in->dumpbinary_path = const_cast<char *>(path);
  in->dumpbinary_rate = 1;
  const int rc = guarded("emd_app_dump_binary", [&] { a->md->dump_binary(step); });
  in->dumpbinary_path = old_path;
  in->dumpbinary_rate = old_rate;
  return rc;
}

} // extern "C"
