#include "comm.h"
#include <cstdio>
#include <cstdlib>
void Comm::error(const char *errormsg) { // src/comm.cpp:65-68
  printf("%s\n", errormsg);
  exit(1);
}
