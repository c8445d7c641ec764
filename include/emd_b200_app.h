/*
 * emd_b200_app.h -- C session API over the whole application (libemd_b200.so).
 * Replaces, for non-C++ hosts, the reference's library use of the ExaMiniMD class
 * (src/examinimd.h:50-71, "ExaMiniMD can be used as a library", src/main.cpp:41-42):
 *   emd_app_create   = ExaMiniMD() + init(argc, argv)            (examinimd.cpp:47-174)
 *   emd_app_advance  = the body of run() for n steps, no output  (examinimd.cpp:192-250)
 *   emd_app_thermo   = Temperature/PotE/KinE                     (examinimd.cpp:252-255)
 *   emd_app_download = the arrays dump_binary writes             (examinimd.cpp:337-343)
 * argv uses the reference's own flags (-il, --neigh-type, --force-iteration, --comm-type, ...).
 */
#ifndef EMD_B200_APP_H
#define EMD_B200_APP_H
#include "emd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct emd_app emd_app;

int emd_app_create(emd_app **out, int argc, const char *const *argv, int device, void *stream);
void emd_app_destroy(emd_app *app);
emd_ctx *emd_app_ctx(emd_app *app);
int emd_app_advance(emd_app *app, int nsteps);
/* the loop of ExaMiniMD::run as the reference times it (src/examinimd.cpp:192-267): nsteps steps WITH the thermo
 * reductions (Temperature, PotE = a second pass over the list, KinE) every `thermo` steps of the deck, nothing printed;
 * h_last_thermo3 (may be NULL) = {T, PE/atom, KE/atom} of the last thermo step executed */
int emd_app_run(emd_app *app, int nsteps, double *h_last_thermo3);
/* same, with the reference's four phase timers (src/examinimd.cpp:183-189) taken as device events:
 * h_seconds4 = {T_Force, T_Neigh, T_Comm, T_Other} accumulated over the nsteps */
int emd_app_advance_timed(emd_app *app, int nsteps, double *h_seconds4);
int emd_app_thermo(emd_app *app, double *T, double *PE_per_atom, double *KE_per_atom);
/* integer properties: "N","N_local","N_ghost","N_max","step","total_neighs","nsteps",
 * "exchange_rate","half_neigh","nbinx","nbiny","nbinz","rank","nranks"; -1 if unknown */
long long emd_app_get(emd_app *app, const char *what);
/* owned atoms [0,N_local) to HOST arrays (any pointer may be NULL) */
int emd_app_download(emd_app *app, int *h_id, int *h_type, double *h_q, double *h_x, double *h_v, double *h_f);
/* HOST x,v,f of the owned atoms to the device (any pointer may be NULL) */
int emd_app_upload(emd_app *app, const double *h_x, const double *h_v, const double *h_f);
/* device pointers of the live arrays: "x","v","f","type","id","q","bincount","binoffsets",
 * "permute","row_map","num_neighs","neighs" (valid until the next rebuild/grow); "snap" = the emd_snap* of a
 * ForceSNAP (NULL for other forces); "tiles" = the
 * emd_tiles* of the last neighbor build, or NULL when the fast path is not in use */
void *emd_app_device_ptr(emd_app *app, const char *what);
int emd_app_neigh_stride(emd_app *app);
int emd_app_dump_binary(emd_app *app, const char *path, int step);

#ifdef __cplusplus
}
#endif
#endif
