// Stub of LAMMPS' pair.h: just enough for pair_reaxc_sunway.h to declare its class (never instantiated here).
#pragma once
#include "lmptype.h"
namespace LAMMPS_NS {
class LAMMPS;
class NeighList;
class Pair {
 public:
  Pair(LAMMPS*) {}
  virtual ~Pair() {}
  virtual void init_style() {}
  virtual void init_list(int, NeighList*) {}
};
}
