// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under sw_reaxff_b200/ may include, link or call this.
//
// Minimal single-rank stand-in for the LAMMPS core that SW_REAXFF plugs into (the core is ABSENT from
// /root/reference).  It restates only what the hot path observes:
//   * triclinic box, periodic remap, periodic-image ghost atoms out to cutghost (LAMMPS Comm::borders semantics),
//   * binned FULL neighbour list incl. rows for ghost atoms (npair_full_bin_ghost_sw5.c:80-228,
//     serial form npair_full_bin_atomonly_sunway.cpp:97-128),
//   * velocity-Verlet fix nve (fix_nve_sw64.c:43-99), units real,
//   * Verlet step order: initial_integrate -> [reneighbour every N | forward_comm] ->
//     qeq pre_force -> pair compute -> reverse_comm -> final_integrate.
#pragma once
#include <vector>

#include "orc_qeq.h"
#include "orc_system.h"

namespace orc {

struct Box {
  double lo[3] = {0, 0, 0};
  double h[6];      // xprd, yprd, zprd, yz, xz, xy
  double h_inv[6];
  void set(double xprd, double yprd, double zprd, double xy, double xz, double yz);
  void x2lamda(const double* x, double* l) const;
  void shift(int sx, int sy, int sz, double* d) const;  // sx*a + sy*b + sz*c
  void cutghost_lamda(double cut, double* cg) const;
};

struct MD {
  System sys;
  QEq qeq;
  Box box;
  int nlocal = 0;
  std::vector<double> v, f, mass;  // v,f: [nlocal][3] (f accumulates ghosts after reverse); mass per LAMMPS type (1-based)
  std::vector<int> ltype;          // LAMMPS types of local atoms
  std::vector<int> ghost_owner, ghost_shift;  // per ghost: owner index, 3 ints shift
  double dt = 0.0625, skin = 2.5, cutneigh = 12.5;
  int every = 5, ago = 0;
  long ntimestep = 0;
  bool qeq_on = true;

  void remap();
  void make_ghosts();
  void forward_x();
  void build_neighbors();
  void force();        // qeq pre_force + pair compute + reverse comm
  void setup();
  void run(int nsteps);
  double kinetic() const;
  double potential() const;
};

void build_full_neighbor_list(System& s, double cutneigh);

}  // namespace orc
