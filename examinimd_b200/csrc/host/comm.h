// comm.h -- abstract communication module (plugin surface of src/comm.h:47-109).
#pragma once
#include "types.h"
#include "system.h"
#include "binning.h"

class Comm {
protected:
  System *system;
  T_X_FLOAT comm_depth;

public:
  Comm(System *s, T_X_FLOAT comm_depth_) : system(s), comm_depth(comm_depth_) {}
  virtual ~Comm() {}
  virtual void init() {}
  virtual void exchange() {}       // move atoms that left the sub-domain / wrap periodic images
  virtual void exchange_halo() {}  // (re)create ghost atoms
  virtual void update_halo() {}    // refresh ghost positions
  // Optional (not in the reference): the same refresh, but the stream is not made to wait for the neighbours' data; the
  // consumer named by Force::gates_halo waits itself.  false: not available, nothing was done (call update_halo()).
  virtual bool update_halo_deferred() { return false; }
  virtual void update_force() {}   // reverse: fold ghost forces back (newton on)
  virtual void reduce_float(T_FLOAT *values, T_INT N) {}
  virtual void reduce_int(T_INT *values, T_INT N) {}
  virtual void reduce_max_float(T_FLOAT *values, T_INT N) {}
  virtual void reduce_max_int(T_INT *values, T_INT N) {}
  virtual void reduce_min_float(T_FLOAT *values, T_INT N) {}
  virtual void reduce_min_int(T_INT *values, T_INT N) {}
  virtual void scan_int(T_INT *values, T_INT N) {}
  virtual void weighted_reduce_float(T_FLOAT *values, T_INT *weight, T_INT N) {}
  // default = one brick covering the whole box (src/comm.cpp:56-63)
  virtual void create_domain_decomposition() {
    system->sub_domain_lo_x = system->sub_domain_lo_y = system->sub_domain_lo_z = 0.0;
    system->sub_domain_x = system->sub_domain_hi_x = system->domain_x;
    system->sub_domain_y = system->sub_domain_hi_y = system->domain_y;
    system->sub_domain_z = system->sub_domain_hi_z = system->domain_z;
  }
  virtual int process_rank() { return 0; }
  virtual int num_processes() { return 1; }
  virtual void error(const char *msg);
  virtual const char *name() { return "InvalidComm"; }
};

#include "modules_comm.h"
