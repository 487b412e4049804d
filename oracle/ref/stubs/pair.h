// Stub of LAMMPS' pair.h (written for this repo): just enough for pair_reaxc_sunway.h to declare its class and for the
// reference's serial energy routines to compile.  Tally hooks are no-ops and every e/v flag is 0, i.e. the routines run
// exactly as under `eflag = vflag = 0` in LAMMPS.
#pragma once
#include "lmptype.h"
#ifndef MIN
#define MIN(A, B) ((A) < (B) ? (A) : (B))
#endif
#ifndef MAX
#define MAX(A, B) ((A) > (B) ? (A) : (B))
#endif
namespace LAMMPS_NS {
class LAMMPS;
class NeighList;
class Pair {
 public:
  Pair(LAMMPS*) {}
  virtual ~Pair() {}
  virtual void init_style() {}
  virtual void init_list(int, NeighList*) {}
  virtual void* extract(const char*, int&) { return nullptr; }
  NeighList* list = nullptr;       // Pair::list of the LAMMPS core
  NeighList* listfull = nullptr;   // the Sunway-patched LAMMPS core adds this member
  int evflag = 0, eflag_either = 0, eflag_global = 0, eflag_atom = 0, vflag_either = 0, vflag_global = 0, vflag_atom = 0;
  double eng_vdwl = 0, eng_coul = 0;
  double virial[6] = {0, 0, 0, 0, 0, 0};
  void ev_tally(int, int, int, int, double, double, double, double, double, double) {}
  void ev_tally_full(int, double, double, double, double, double, double) {}
  void ev_tally_xyz(int, int, int, int, double, double, double, double, double, double, double, double) {}
  void ev_tally3(int, int, int, double, double, double*, double*, double*, double*) {}
  void v_tally(int, double*, double*) {}
  void v_tally3(int, int, int, double*, double*, double*, double*) {}
  void v_tally4(int, int, int, int, double*, double*, double*, double*, double*, double*) {}
};
}
