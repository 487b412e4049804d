// Parameter model of the B200 ReaxFF path: flat POD tables that are copied verbatim to HBM.
//
// Replaces the reference's pointer-chasing reax_interaction (sbp[], tbp[][], thbp[][][], fbp[][][][], hbp[][][];
// /root/reference/reaxc_ctypes_sunway.h:240-400) and the per-kernel "param pack" re-packing that every
// *_sw64.c wrapper does before athread_spawn (e.g. reaxc_forces_sw64.c:640-668).
#pragma once
#include <string>
#include <vector>

namespace rxb {

constexpr int kMaxAngleSets = 5;  // MAX_3BODY_PARAM, reaxc_defs_sunway.h:113

constexpr double kCeleConst = 332.06371;   // C_ele, reaxc_defs_sunway.h:62

struct AtomPar {  // single_body_parameters
  double r_s, r_pi, r_pi_pi;
  double valency, valency_e, valency_boc, valency_val, nlp_opt, mass;
  double p_lp2, p_ovun2, p_ovun5, p_val3, p_val5;
  double chi, eta, gamma;
  double r_vdw, epsilon, alpha, gamma_w, rcore2, ecore2, acore2, lgcij, lgre;
  double b_o_131, b_o_132, b_o_133;
  int p_hbond;
  int is_carbon;  // name == "C" (the C2-correction test in reaxc_multi_body_sw64.c:127 is a strcmp)
};

struct PairPar {  // two_body_parameters
  double p_bo1, p_bo2, p_bo3, p_bo4, p_bo5, p_bo6;
  double r_s, r_p, r_pp;
  double p_boc3, p_boc4, p_boc5;
  double p_be1, p_be2, De_s, De_p, De_pp, p_ovun1;
  double D, alpha, r_vdW, gamma_w, gamma;
  double rcore, ecore, acore, lgcij, lgre;
  double v13cor, ovc;
  double powgi_vdW1;   // derived: (1/gamma_w)^p_vdW1, hoisted out of the pair loop (reaxc_nonbonded_sw64.c:137)
  double inv_r_vdW, alpha_over_r_vdW; // derived: 1/r_vdW, alpha/r_vdW
  double log_r_s, log_r_p, log_r_pp;  // derived: logs of the bond radii, so that (d/r)^p = exp(p (log d - log r))
  double d_bond_max;   // derived: beyond this distance BO' < bo_cut for certain (BOp_single would reject), <= bond_cut
};

struct AnglePar { double theta_00, p_val1, p_val2, p_coa1, p_val7, p_pen1, p_val4; };
struct AngleSet { int cnt; int pad_; AnglePar prm[kMaxAngleSets]; };
struct TorsPar { double V1, V2, V3, p_tor1, p_cot1; int cnt; int pad_; };
struct HbPar { double r0_hb, p_hb1, p_hb2, p_hb3; };

struct Control {
  double bo_cut, nonb_low, nonb_cut, bond_cut, hbond_cut, bg_cut, thb_cut, thb_cutsq;
  int tabulate, energy_update_freq, lgflag, enobondsflag;
  int vdw_type, pad_;
  double Tap[8];
};

struct ForceField {
  int nt = 0;  // element types in the force field
  std::vector<double> gp;
  Control ctl{};
  std::vector<std::string> names;
  std::vector<AtomPar> atom;
  std::vector<PairPar> pair;    // nt*nt
  std::vector<AngleSet> angle;  // nt^3, [k][j][h] with j the centre
  std::vector<TorsPar> tors;    // nt^4
  std::vector<HbPar> hb;        // nt^3, [acceptor-bonded i][H j][k]
  std::vector<int> map;         // LAMMPS type (1-based) -> element index or -1

  // pair_style reax/c <control|NULL> keywords; pair_coeff * * <ffield> <elements...>
  std::string load_control(const char* path);                     // "" on success
  std::string load_ffield(const char* path);                      // "" on success
  std::string set_elements(int ntypes, const char* const* names); // "" on success
  void derive();                                                  // taper + hoisted constants
  std::vector<double> dump() const;                               // canonical flat dump (parity tests)
  // tabulated long-range mode (ctl.tabulate > 0): [nt*nt][n = tabulate+2][4 sets CEvd,CEclmb,e_vdW,e_ele][a,b,c,d]
  std::vector<double> lookup_tables(int* n_out, double* dx_out) const;
};

}  // namespace rxb
