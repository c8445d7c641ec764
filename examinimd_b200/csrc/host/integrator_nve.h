// IntegratorNVE -- velocity Verlet on the device (plugin surface of src/integrator_nve.h; arithmetic of
// src/integrator_nve.cpp:41-121 in kernels/integrator.cu).  The host object only keeps the two step factors and
// forwards to the C ABI; final_initial_integrate() is the fused variant the driver uses between unobserved steps.
#ifndef INTEGRATOR_NVE_H
#define INTEGRATOR_NVE_H
#include "integrator.h"

class IntegratorNVE : public Integrator {
public:
  explicit IntegratorNVE(System *s);

  void initial_integrate() override;        // emd_nve_initial_integrate:       v += dtf/m f ; x += dtv v
  void final_integrate() override;          // emd_nve_final_integrate:         v += dtf/m f
  void final_initial_integrate() override;  // emd_nve_final_initial_integrate: both, one pass over the atoms
  bool step_factors(T_V_FLOAT *dtf_, T_V_FLOAT *dtv_) override { *dtf_ = dtf; *dtv_ = dtv; return true; }
  const char *name() override;

private:
  T_V_FLOAT dtv; // dt
  T_V_FLOAT dtf; // dt / (2 mvv2e)
};
#endif
