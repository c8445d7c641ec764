#include "system.h"
#include <cstdio>
#include <cstdlib>
#include <utility>

static void die(const char *what) {
  fprintf(stderr, "System: %s: %s\n", what, emd_last_error());
  emd_host_exit(1);
}

System::System() {
  N = N_max = N_local = N_ghost = 0;
  ntypes = 1;
  x = v = f = nullptr; type = nullptr; id = nullptr; q = nullptr; mass = nullptr;
  x_alt = v_alt = f_alt = nullptr; type_alt = nullptr; id_alt = nullptr; q_alt = nullptr;
  domain_x = domain_y = domain_z = 0.0;
  sub_domain_x = sub_domain_y = sub_domain_z = 0.0;
  sub_domain_lo_x = sub_domain_lo_y = sub_domain_lo_z = 0.0;
  sub_domain_hi_x = sub_domain_hi_y = sub_domain_hi_z = 0.0;
  mvv2e = boltz = dt = 0.0;
  do_print = true;
  print_lammps = false;
  ctx = nullptr;
}

System::~System() { release(); }

void System::init() { h_mass.assign(ntypes, 0.0); }

void System::release() {
  void *all[] = {x, v, f, type, id, q, mass, x_alt, v_alt, f_alt, type_alt, id_alt, q_alt};
  for (void *p : all) if (p) emd_free(p);
  x = v = f = nullptr; type = nullptr; id = nullptr; q = nullptr; mass = nullptr;
  x_alt = v_alt = f_alt = nullptr; type_alt = nullptr; id_alt = nullptr; q_alt = nullptr;
}

void System::destroy() {
  release();
  N_max = N_local = N_ghost = 0;
  ntypes = 1;
}

template <class T>
static void regrow(emd_ctx *ctx, T *&p, size_t old_n, size_t new_n, size_t width, bool keep) {
  void *np = nullptr;
  if (emd_malloc(&np, sizeof(T) * width * new_n)) die("grow/malloc");
  if (emd_memset_zero(ctx, np, sizeof(T) * width * new_n)) die("grow/memset"); // resize zero-fills the tail
  if (keep && p && old_n) {
    if (emd_memcpy_d2d(ctx, np, p, sizeof(T) * width * old_n)) die("grow/copy");
  }
  if (emd_ctx_sync(ctx)) die("grow/sync");
  if (p) emd_free(p);
  p = static_cast<T *>(np);
}

void System::grow(T_INT N_new) {
  if (N_new <= N_max) return;
  // Head-room of 1/16 of the atoms on top of what the caller asks for: the ghost counts of a melt fluctuate by a few atoms per
  // face from one re-neighboring to the next, and every growth is twelve allocations, six copies and a synchronisation (the
  // first re-neighboring after a lattice start used to grow once per halo phase: 8.5 ms at 2 M atoms, which a 20-step
  // measurement window sees as +0.4 ms per step).  The resize semantics of the reference (system.cpp:76-103) are unchanged.
  N_new += N_new / 16;
  const size_t o = N_max, n = N_new;
  regrow(ctx, x, o, n, 3, true); regrow(ctx, v, o, n, 3, true); regrow(ctx, f, o, n, 3, true);
  regrow(ctx, id, o, n, 1, true); regrow(ctx, type, o, n, 1, true); regrow(ctx, q, o, n, 1, true);
  regrow(ctx, x_alt, o, n, 3, false); regrow(ctx, v_alt, o, n, 3, false); regrow(ctx, f_alt, o, n, 3, false);
  regrow(ctx, id_alt, o, n, 1, false); regrow(ctx, type_alt, o, n, 1, false); regrow(ctx, q_alt, o, n, 1, false);
  N_max = N_new;
}

void System::swap_sorted() {
  std::swap(x, x_alt); std::swap(v, v_alt); std::swap(f, f_alt);
  std::swap(type, type_alt); std::swap(id, id_alt); std::swap(q, q_alt);
}

void System::upload(const HostAtoms &h, T_INT n, bool with_f) {
  if (n > N_max) grow(n);
  int rc = 0;
  rc |= emd_memcpy_h2d(ctx, x, h.x.data(), sizeof(double) * 3 * (size_t)n);
  rc |= emd_memcpy_h2d(ctx, v, h.v.data(), sizeof(double) * 3 * (size_t)n);
  if (with_f) rc |= emd_memcpy_h2d(ctx, f, h.f.data(), sizeof(double) * 3 * (size_t)n);
  rc |= emd_memcpy_h2d(ctx, q, h.q.data(), sizeof(double) * (size_t)n);
  rc |= emd_memcpy_h2d(ctx, type, h.type.data(), sizeof(int) * (size_t)n);
  rc |= emd_memcpy_h2d(ctx, id, h.id.data(), sizeof(int) * (size_t)n);
  rc |= emd_ctx_sync(ctx);
  if (rc) die("upload");
}

void System::download(HostAtoms &h, T_INT n) const {
  h.resize(n);
  int rc = 0;
  rc |= emd_memcpy_d2h(ctx, h.x.data(), x, sizeof(double) * 3 * (size_t)n);
  rc |= emd_memcpy_d2h(ctx, h.v.data(), v, sizeof(double) * 3 * (size_t)n);
  rc |= emd_memcpy_d2h(ctx, h.f.data(), f, sizeof(double) * 3 * (size_t)n);
  rc |= emd_memcpy_d2h(ctx, h.q.data(), q, sizeof(double) * (size_t)n);
  rc |= emd_memcpy_d2h(ctx, h.type.data(), type, sizeof(int) * (size_t)n);
  rc |= emd_memcpy_d2h(ctx, h.id.data(), id, sizeof(int) * (size_t)n);
  if (rc) die("download");
}

void System::set_mass(const std::vector<double> &m) {
  h_mass = m;
  if (mass) emd_free(mass);
  void *p = nullptr;
  if (emd_malloc(&p, sizeof(double) * m.size())) die("mass/malloc");
  mass = static_cast<double *>(p);
  if (emd_memcpy_h2d(ctx, mass, m.data(), sizeof(double) * m.size()) || emd_ctx_sync(ctx)) die("mass/copy");
}
