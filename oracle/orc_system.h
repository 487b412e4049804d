// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under sw_reaxff_b200/ may include, link or call this.
//
// Data model of the CPU restatement (single rank, local atoms [0,n) + ghost atoms [n,N), exactly the
// index space LAMMPS hands to pair reax/c: pair_reaxc_sunway.cpp:568-570).
#pragma once
#include <vector>

#include "orc_params.h"

namespace orc {

struct Bond {  // bond_data + bond_order_data + SoA side arrays (reaxc_ctypes_sunway.h:590-640,759-782)
  int nbr, sym;
  double d, dvec[3];
  double BO, BO_s, BO_pi, BO_pi2;
  double dBOp[3], dln_BOp_s[3], dln_BOp_pi[3], dln_BOp_pi2[3];
  double C1dbo, C2dbo, C3dbo;
  double C1dbopi, C2dbopi, C3dbopi, C4dbopi;
  double C1dbopi2, C2dbopi2, C3dbopi2, C4dbopi2;
  double Cdbo, Cdbopi, Cdbopi2;
};

struct HBond { int nbr; double d, dvec[3]; };

struct Energies {  // simulation_data::my_en
  double e_bond = 0, e_ov = 0, e_un = 0, e_lp = 0, e_ang = 0, e_pen = 0, e_coa = 0, e_hb = 0, e_tor = 0, e_con = 0,
         e_vdW = 0, e_ele = 0, e_pol = 0;
};

struct System {
  Params prm;
  int n = 0, N = 0;
  std::vector<double> x;   // [N][3]
  std::vector<int> type;   // ff element index (map applied), -1 = NULL
  std::vector<int> tag;
  std::vector<double> q;
  // full neighbour list (a1): CSR over all N rows, r <= cutneigh
  std::vector<long> nb_off;
  std::vector<int> nb;
  // bond list (rows sorted by neighbour index), CSR
  std::vector<int> b_start, b_end;
  std::vector<Bond> bonds;
  // hbond list for local H atoms
  std::vector<int> Hindex, hb_start, hb_end;
  std::vector<HBond> hbonds;
  // workspace
  std::vector<double> total_bo, Delta_boc, Deltap, Deltap_boc, Delta, Delta_e, Delta_val, vlpex, nlp, Delta_lp, Clp,
      dDelta_lp, nlp_temp, Delta_lp_temp, dDelta_lp_temp;
  std::vector<double> dDeltap_self;  // [N][3]
  std::vector<double> fCd;           // [N][4]  (-force xyz, CdDelta)
  Energies en;
  double virial[6];
  double eng_vdwl = 0, eng_coul = 0;
};

void build_bond_list(System& s);       // a4  Init_Forces_noQEq_Full / BOp_single
void build_hbond_list(System& s);      // a5
void build_lookup_tables(Params& P);    // a9' (dead code in the reference, restated)
void nonbonded(System& s);             // a9
void bond_orders(System& s);           // a6
void bonds_atom_energy(System& s);     // a7
void hydrogen_bonds(System& s);        // a8
void valence_torsion(System& s);       // a10
void add_dbond_forces(System& s);      // a11
void compute_forces(System& s);        // Compute_Forces order, reaxc_forces_sunway.cpp:1297-1365

}  // namespace orc
