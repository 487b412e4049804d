// ORACLE/_REF — TEST INFRASTRUCTURE ONLY.
// Harness (written for this repo) around the reference's UNMODIFIED parser sources, compiled from where they lie in
// /root/reference:  Read_Force_Field (reaxc_ffield_sunway.cpp:35), Read_Control_File (reaxc_control_sunway.cpp:34),
// Tokenize/scalloc (reaxc_tool_box_sunway.cpp).  It dumps every parsed parameter in the canonical order used by
// orc_params_dump / rxb_params_dump so that the oracle and the product parser can be compared with the reference itself.
#include <cstring>
#include <strings.h>
#include <vector>

#include "pair_reaxc_sunway.h"
#include "reaxc_control_sunway.h"
#include "reaxc_ffield_sunway.h"

using namespace REAXC_SUNWAY_NS;

extern "C" long ref_params_dump(const char* control_file, const char* ffield_file, int ntypes, const char** elements,
                                int lgflag, double* out, long cap) {
  reax_interaction rp;
  memset(&rp, 0, sizeof(rp));
  control_params control;
  memset(&control, 0, sizeof(control));
  output_controls out_control;
  memset(&out_control, 0, sizeof(out_control));
  control.lgflag = lgflag;
  bool null_control = (control_file == nullptr) || !strcmp(control_file, "NULL");
  if (null_control) {  // pair_reaxc_sunway.cpp:208-232
    control.bond_cut = 5.; control.hbond_cut = 7.50; control.thb_cut = 0.001; control.thb_cutsq = 0.00001; control.bg_cut = 0.3;
    control.tabulate = 0; out_control.energy_update_freq = 0;
  } else {
    Read_Control_File((char*)control_file, &control, &out_control);
  }
  FILE* fp = fopen(ffield_file, "r");
  if (!fp) return -1;
  Read_Force_Field(fp, &rp, &control);
  const int nt = rp.num_atom_types;
  std::vector<double> v;
  v.push_back(nt); v.push_back(rp.gp.vdw_type); v.push_back(rp.gp.n_global);
  for (int i = 0; i < rp.gp.n_global; i++) v.push_back(rp.gp.l[i]);
  v.push_back(control.bo_cut); v.push_back(control.nonb_low); v.push_back(control.nonb_cut); v.push_back(control.bond_cut);
  v.push_back(control.hbond_cut); v.push_back(control.bg_cut); v.push_back(control.thb_cut); v.push_back(control.thb_cutsq);
  v.push_back(control.tabulate); v.push_back(out_control.energy_update_freq);
  for (int i = 0; i < 8; i++) v.push_back(0.0);  // Tap[] is computed by Init_Taper (reaxc_init_md), not by the parser: slot left 0
  for (int i = 0; i < nt; i++) {
    const single_body_parameters& s = rp.sbp[i];
    double a[] = {s.r_s, s.valency, s.mass, s.r_vdw, s.epsilon, s.gamma, s.r_pi, s.valency_e, s.nlp_opt, s.alpha,
                  s.gamma_w, s.valency_boc, s.p_ovun5, s.chi, s.eta, (double)s.p_hbond, s.r_pi_pi, s.p_lp2, s.b_o_131,
                  s.b_o_132, s.b_o_133, s.p_ovun2, s.p_val3, s.valency_val, s.p_val5, s.rcore2, s.ecore2, s.acore2,
                  s.lgcij, s.lgre};
    v.insert(v.end(), a, a + sizeof(a) / sizeof(double));
  }
  for (int i = 0; i < nt; i++)
    for (int j = 0; j < nt; j++) {
      const two_body_parameters& t = rp.tbp[i][j];
      double a[] = {t.p_bo1, t.p_bo2, t.p_bo3, t.p_bo4, t.p_bo5, t.p_bo6, t.r_s, t.r_p, t.r_pp, t.p_boc3, t.p_boc4,
                    t.p_boc5, t.p_be1, t.p_be2, t.De_s, t.De_p, t.De_pp, t.p_ovun1, t.D, t.alpha, t.r_vdW, t.gamma_w,
                    t.rcore, t.ecore, t.acore, t.lgcij, t.lgre, t.gamma, t.v13cor, t.ovc};
      v.insert(v.end(), a, a + sizeof(a) / sizeof(double));
    }
  for (int i = 0; i < nt; i++)
    for (int j = 0; j < nt; j++)
      for (int k = 0; k < nt; k++) {
        const three_body_header& t = rp.thbp[i][j][k];
        v.push_back(t.cnt);
        for (int c = 0; c < 5; c++) {
          const three_body_parameters& q = t.prm[c];
          double a[] = {q.theta_00, q.p_val1, q.p_val2, q.p_coa1, q.p_val7, q.p_pen1, q.p_val4};
          v.insert(v.end(), a, a + 7);
        }
      }
  for (int i = 0; i < nt; i++)
    for (int j = 0; j < nt; j++)
      for (int k = 0; k < nt; k++)
        for (int l = 0; l < nt; l++) {
          const four_body_header& t = rp.fbp[i][j][k][l];
          v.push_back(t.cnt);
          const four_body_parameters& q = t.prm[0];
          double a[] = {q.V1, q.V2, q.V3, q.p_tor1, q.p_cot1};
          v.insert(v.end(), a, a + 5);
        }
  for (int i = 0; i < nt; i++)
    for (int j = 0; j < nt; j++)
      for (int k = 0; k < nt; k++) {
        const hbond_parameters& t = rp.hbp[i][j][k];
        v.push_back(t.r0_hb); v.push_back(t.p_hb1); v.push_back(t.p_hb2); v.push_back(t.p_hb3);
      }
  for (int i = 0; i < ntypes; i++) {  // element map, pair_reaxc_sunway.cpp:318-336
    int m = -1;
    if (strcmp(elements[i], "NULL"))
      for (int j = 0; j < nt; j++)
        if (strcasecmp(elements[i], rp.sbp[j].name) == 0) m = j;
    v.push_back(m);
  }
  if (out) for (long i = 0; i < (long)v.size() && i < cap; i++) out[i] = v[i];
  return (long)v.size();
}
