/* temporary stub until oracle_snap.c lands */
#include "oracle.h"
#include <stddef.h>
orc_force_snap *orc_force_snap_create(int ntypes) { (void)ntypes; return NULL; }
void orc_force_snap_destroy(orc_force_snap *f) { (void)f; }
int orc_force_snap_init_coeff(orc_force_snap *f, int nargs, char args[][ORC_WORD], const char *dir) { (void)f; (void)nargs; (void)args; (void)dir; return -1; }
void orc_force_snap_compute(orc_force_snap *f, orc_system *s, const orc_neighbor *n) { (void)f; (void)s; (void)n; }
