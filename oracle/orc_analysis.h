// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under sw_reaxff_b200/ may include, link or call this.
// fix reax/c/bonds and fix reax/c/species restated on the oracle's MD state (see orc_analysis.cpp).
#pragma once
#include <string>
#include <vector>

#include "orc_md.h"

namespace orc {

constexpr int MAXSPECBOND = 12;  // reaxc_defs_sunway.h:124

std::string bonds_text(const MD& md, long ntimestep);  // the block fix reax/c/bonds appends to its file

struct SpeciesFix {
  int nevery = 1, nrepeat = 1, nfreq = 1, ntypes = 0;
  long nvalid = -1, ave_nvalid = -1;
  int irepeat = 0;
  std::vector<double> BOCut;          // (ntypes+1)^2
  std::vector<int> tmpid;             // [N][MAXSPECBOND] local index of the bonded partner, 0 = unused
  std::vector<double> tmpbo, array;   // current and averaged bond orders
  std::vector<double> qxyz;           // [nlocal][4]: the fix ave/atom result for compute SPEC/ATOM's q, x, y, z columns
  std::vector<double> clusterID;
  int Nmole = 0, Nspec = 0;
  std::vector<int> MolName, NMol, composition;
  std::string error;

  void init(const MD& md, int nevery, int nrepeat, int nfreq, const std::vector<double>& bocut);
  bool post_integrate(const MD& md, long step);
  void pair_find_bond(const MD& md);
  void find_molecule(const MD& md);
  void sort_molecule(const MD& md);
  void find_species(const MD& md);
  std::string formulas_text(long ntimestep) const;
  // WritePos (fix_reaxc_species_sunway.cpp:814-925) for the single-file form of `position`; box6 = boxlo[3], boxhi[3]
  std::string pos_text(const MD& md, long ntimestep, const double* box6);
};

}  // namespace orc
