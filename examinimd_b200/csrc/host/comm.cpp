#include "comm.h"
#include <cstdio>
#include <cstdlib>
void Comm::error(const char *errormsg) { // src/comm.cpp:65-68
  printf("%s\n", errormsg);
  emd_host_exit(1);
}

// see types.h
static bool g_throw_on_exit = false;
void emd_host_throw_on_exit(bool on) { g_throw_on_exit = on; }
void emd_host_exit(int code) {
  fflush(stdout);
  fflush(stderr);
  if (g_throw_on_exit) throw EmdFatal{code};
  exit(code);
}
