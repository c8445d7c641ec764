// property.h -- thermo output quantities (src/property_temperature.*, property_kine.*, property_pote.*).
#pragma once
#include "comm.h"
#include "force.h"

class Temperature {
  Comm *comm;
public:
  Temperature(Comm *comm_) : comm(comm_) {}
  T_V_FLOAT compute(System *system);
};

class KinE {
  Comm *comm;
public:
  KinE(Comm *comm_) : comm(comm_) {}
  T_V_FLOAT compute(System *system);
};

class PotE {
  Comm *comm;
public:
  PotE(Comm *comm_) : comm(comm_) {}
  T_F_FLOAT compute(System *system, Binning *binning, Neighbor *neighbor, Force *force);
};
