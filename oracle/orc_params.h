// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under sw_reaxff_b200/ may include, link or call this.
//
// CPU restatement of the ReaxFF parameter model of run-towards-the-future/SW_REAXFF.
// Follows (file:line in /root/reference):
//   reaxc_ffield_sunway.cpp:35-714   Read_Force_Field   (sections, combination rules, vdw_type)
//   reaxc_control_sunway.cpp:34-391  Read_Control_File  (defaults + keys)
//   reaxc_tool_box_sunway.cpp:42-57  Tokenize           (separators "\t \n\r\f!=")
//   reaxc_init_md_sunway.cpp:100-136 Init_Taper
//   pair_reaxc_sunway.cpp:202-290    settings() defaults (enobondsflag=1, lgflag=0 ...)
//   pair_reaxc_sunway.cpp:294-362    coeff() element map
// Parity status: pinned against the reference's own parsers compiled from /root/reference into
// oracle/_ref (see oracle/Makefile target `ref`, tests/test_oracle_vs_ref.py).
#pragma once
#include <string>
#include <vector>

namespace orc {

struct Sbp {
  char name[16];
  double r_s, valency, mass, r_vdw, epsilon, gamma, r_pi, valency_e, nlp_opt;
  double alpha, gamma_w, valency_boc, p_ovun5, chi, eta;
  int p_hbond;
  double r_pi_pi, p_lp2, b_o_131, b_o_132, b_o_133;
  double p_ovun2, p_val3, valency_val, p_val5, rcore2, ecore2, acore2;
  double lgcij, lgre;
};

struct Tbp {
  double p_bo1, p_bo2, p_bo3, p_bo4, p_bo5, p_bo6;
  double r_s, r_p, r_pp;
  double p_boc3, p_boc4, p_boc5;
  double p_be1, p_be2, De_s, De_p, De_pp;
  double p_ovun1;
  double D, alpha, r_vdW, gamma_w, rcore, ecore, acore, lgcij, lgre;
  double gamma;
  double v13cor, ovc;
};

// cubic-spline tables of the tabulated long-range mode (LR_lookup_table, reaxc_ctypes_sunway.h:846-860):
// per type pair (i <= j) five tables H, vdW, CEvd, ele, CEclmb of n = tabulate + 2 coefficient sets
struct SplineCoef { double a, b, c, d; };
struct Lookup {
  int n = 0;
  double dx = 0, inv_dx = 0;
  std::vector<SplineCoef> tables;
  SplineCoef* at(int nt, int i, int j, int which) { return &tables[(((size_t)i * nt + j) * 5 + which) * n]; }
  const SplineCoef* at(int nt, int i, int j, int which) const { return &tables[(((size_t)i * nt + j) * 5 + which) * n]; }
};

struct Thbp { double theta_00, p_val1, p_val2, p_coa1, p_val7, p_pen1, p_val4; };
struct ThbHeader { int cnt; Thbp prm[5]; };
struct Fbp { double V1, V2, V3, p_tor1, p_cot1; };
struct FbHeader { int cnt; Fbp prm[5]; };
struct Hbp { double r0_hb, p_hb1, p_hb2, p_hb3; };

struct Params {
  int nt = 0;  // number of element types in the force field
  std::vector<double> gp;
  int vdw_type = 0;
  std::vector<Sbp> sbp;
  std::vector<Tbp> tbp;        // [i*nt+j]
  std::vector<ThbHeader> thbp; // [(i*nt+j)*nt+k]
  std::vector<FbHeader> fbp;   // [((i*nt+j)*nt+k)*nt+l]
  std::vector<Hbp> hbp;        // [(i*nt+j)*nt+k]
  // control
  double bo_cut = 0, nonb_low = 0, nonb_cut = 0;
  double bond_cut = 5.0, hbond_cut = 7.5, bg_cut = 0.3, thb_cut = 0.001, thb_cutsq = 0.00001;
  int tabulate = 0, lgflag = 0, enobondsflag = 1, energy_update_freq = 0;
  double Tap[8];
  Lookup lookup;               // built on first use when tabulate > 0
  // LAMMPS type (1-based) -> ff element index, or -1 (NULL)
  std::vector<int> map;

  const Tbp& tb(int i, int j) const { return tbp[i * nt + j]; }
  Tbp& tb(int i, int j) { return tbp[i * nt + j]; }
  ThbHeader& thb(int i, int j, int k) { return thbp[(i * nt + j) * nt + k]; }
  const ThbHeader& thb(int i, int j, int k) const { return thbp[(i * nt + j) * nt + k]; }
  FbHeader& fb(int i, int j, int k, int l) { return fbp[((i * nt + j) * nt + k) * nt + l]; }
  const FbHeader& fb(int i, int j, int k, int l) const { return fbp[((i * nt + j) * nt + k) * nt + l]; }
  Hbp& hb(int i, int j, int k) { return hbp[(i * nt + j) * nt + k]; }
  const Hbp& hb(int i, int j, int k) const { return hbp[(i * nt + j) * nt + k]; }
};

// returns empty string on success, else error message
std::string read_force_field(const char* path, Params& p);
std::string read_control(const char* path, Params& p);
std::string set_element_map(Params& p, int ntypes, const char* const* elements);
void init_taper(Params& p);

}  // namespace orc
