// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_params.h header for the reference ranges restated here).
#include "orc_params.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <strings.h>

namespace orc {

// reaxc_tool_box_sunway.cpp:42-57
static std::vector<std::string> tokenize(const char* s) {
  std::vector<std::string> out;
  char buf[1024];
  strncpy(buf, s, sizeof(buf) - 1);
  buf[sizeof(buf) - 1] = 0;
  const char* sep = "\t \n\r\f!=";
  for (char* w = strtok(buf, sep); w; w = strtok(nullptr, sep)) out.push_back(w);
  return out;
}

static double tokd(const std::vector<std::string>& t, size_t i) {
  return i < t.size() ? atof(t[i].c_str()) : 0.0;
}
static int toki(const std::vector<std::string>& t, size_t i) {
  return i < t.size() ? atoi(t[i].c_str()) : 0;
}

// reaxc_ffield_sunway.cpp:35-714
std::string read_force_field(const char* path, Params& p) {
  FILE* fp = fopen(path, "r");
  if (!fp) return std::string("cannot open force field file ") + path;
  char s[1024];
  // The reference tokenises into a persistent buffer and never clears it (reaxc_tool_box_sunway.cpp:42-57): a field
  // missing on the current line reads the token a previous, longer line left in that slot (this decides e.g. the
  // lgcij column of the off-diagonal section, reaxc_ffield_sunway.cpp:512).  Restated as-is.
  std::vector<std::string> slots;
  auto line = [&]() -> std::vector<std::string> {
    if (!fgets(s, sizeof(s), fp)) s[0] = 0;
    std::vector<std::string> now = tokenize(s);
    if (slots.size() < now.size()) slots.resize(now.size());
    for (size_t i = 0; i < now.size(); i++) slots[i] = now[i];
    return slots;
  };
  line();  // header comment
  auto t = line();
  int n = toki(t, 0);
  if (n < 1) { fclose(fp); return "number of globals in ffield file is 0"; }
  p.gp.assign(n, 0.0);
  for (int i = 0; i < n; i++) { t = line(); p.gp[i] = tokd(t, 0); }
  p.bo_cut = 0.01 * p.gp[29];
  p.nonb_low = p.gp[11];
  p.nonb_cut = p.gp[12];

  t = line();
  const int nt = p.nt = toki(t, 0);
  line(); line(); line();
  p.sbp.assign(nt, Sbp{});
  p.tbp.assign((size_t)nt * nt, Tbp{});
  p.thbp.assign((size_t)nt * nt * nt, ThbHeader{});
  p.fbp.assign((size_t)nt * nt * nt * nt, FbHeader{});
  p.hbp.assign((size_t)nt * nt * nt, Hbp{});
  std::vector<char> tor_flag((size_t)nt * nt * nt * nt, 0);
  auto tf = [&](int a, int b, int c, int d) -> char& { return tor_flag[((a * nt + b) * nt + c) * nt + d]; };
  p.vdw_type = 0;

  for (int i = 0; i < nt; i++) {
    Sbp& e = p.sbp[i];
    t = line();
    memset(e.name, 0, sizeof(e.name));
    for (size_t j = 0; j < t[0].size() && j < sizeof(e.name) - 1; j++) e.name[j] = toupper(t[0][j]);
    e.r_s = tokd(t, 1); e.valency = tokd(t, 2); e.mass = tokd(t, 3); e.r_vdw = tokd(t, 4);
    e.epsilon = tokd(t, 5); e.gamma = tokd(t, 6); e.r_pi = tokd(t, 7); e.valency_e = tokd(t, 8);
    e.nlp_opt = 0.5 * (e.valency_e - e.valency);
    t = line();
    e.alpha = tokd(t, 0); e.gamma_w = tokd(t, 1); e.valency_boc = tokd(t, 2); e.p_ovun5 = tokd(t, 3);
    e.chi = tokd(t, 5); e.eta = 2.0 * tokd(t, 6); e.p_hbond = (int)tokd(t, 7);
    t = line();
    e.r_pi_pi = tokd(t, 0); e.p_lp2 = tokd(t, 1); e.b_o_131 = tokd(t, 3); e.b_o_132 = tokd(t, 4);
    e.b_o_133 = tokd(t, 5);
    t = line();
    if (tokenize(s).size() < 3) { fclose(fp); return "Inconsistent ffield file"; }
    e.p_ovun2 = tokd(t, 0); e.p_val3 = tokd(t, 1); e.valency_val = tokd(t, 3); e.p_val5 = tokd(t, 4);
    e.rcore2 = tokd(t, 5); e.ecore2 = tokd(t, 6); e.acore2 = tokd(t, 7);
    if (p.lgflag) {
      t = line();
      if (tokenize(s).size() > 3) { fclose(fp); return "Inconsistent ffield file (lg)"; }
      e.lgcij = tokd(t, 0); e.lgre = tokd(t, 1);
    }
    // vdw_type detection, :241-293
    if (e.rcore2 > 0.01 && e.acore2 > 0.01) {
      if (e.gamma_w > 0.5) {
        if (!(p.vdw_type != 0 && p.vdw_type != 3)) p.vdw_type = 3;
      } else {
        if (!(p.vdw_type != 0 && p.vdw_type != 2)) p.vdw_type = 2;
      }
    } else {
      if (e.gamma_w > 0.5) {
        if (!(p.vdw_type != 0 && p.vdw_type != 1)) p.vdw_type = 1;
      } else {
        fclose(fp);
        return std::string("inconsistent vdWaals-parameters: no shielding or inner-wall set for ") + e.name;
      }
    }
  }
  // :296-304
  for (int i = 0; i < nt; i++)
    if (p.sbp[i].mass < 21 && p.sbp[i].valency_val != p.sbp[i].valency_boc)
      p.sbp[i].valency_val = p.sbp[i].valency_boc;

  // two-body, :306-363
  t = line();
  int l = toki(t, 0);
  line();
  for (int i = 0; i < l; i++) {
    t = line();
    int j = toki(t, 0) - 1, k = toki(t, 1) - 1;
    auto t2 = t;
    if (j < nt && k < nt) {
      Tbp &a = p.tb(j, k), &b = p.tb(k, j);
      a.De_s = b.De_s = tokd(t, 2); a.De_p = b.De_p = tokd(t, 3); a.De_pp = b.De_pp = tokd(t, 4);
      a.p_be1 = b.p_be1 = tokd(t, 5); a.p_bo5 = b.p_bo5 = tokd(t, 6); a.v13cor = b.v13cor = tokd(t, 7);
      a.p_bo6 = b.p_bo6 = tokd(t, 8); a.p_ovun1 = b.p_ovun1 = tokd(t, 9);
      t = line();
      a.p_be2 = b.p_be2 = tokd(t, 0); a.p_bo3 = b.p_bo3 = tokd(t, 1); a.p_bo4 = b.p_bo4 = tokd(t, 2);
      a.p_bo1 = b.p_bo1 = tokd(t, 4); a.p_bo2 = b.p_bo2 = tokd(t, 5); a.ovc = b.ovc = tokd(t, 6);
    }
    // NOTE: when the pair is out of range the reference does not consume the 2nd line (:320,343);
    // restated as-is.
  }
  // combination rules, :365-461
  for (int i = 0; i < nt; i++)
    for (int j = i; j < nt; j++) {
      const Sbp &si = p.sbp[i], &sj = p.sbp[j];
      Tbp &a = p.tb(i, j), &b = p.tb(j, i);
      a.r_s = 0.5 * (si.r_s + sj.r_s);         b.r_s = 0.5 * (sj.r_s + si.r_s);
      a.r_p = 0.5 * (si.r_pi + sj.r_pi);       b.r_p = 0.5 * (sj.r_pi + si.r_pi);
      a.r_pp = 0.5 * (si.r_pi_pi + sj.r_pi_pi); b.r_pp = 0.5 * (sj.r_pi_pi + si.r_pi_pi);
      a.p_boc3 = sqrt(si.b_o_132 * sj.b_o_132); b.p_boc3 = sqrt(sj.b_o_132 * si.b_o_132);
      a.p_boc4 = sqrt(si.b_o_131 * sj.b_o_131); b.p_boc4 = sqrt(sj.b_o_131 * si.b_o_131);
      a.p_boc5 = sqrt(si.b_o_133 * sj.b_o_133); b.p_boc5 = sqrt(sj.b_o_133 * si.b_o_133);
      a.D = sqrt(si.epsilon * sj.epsilon);      b.D = sqrt(sj.epsilon * si.epsilon);
      a.alpha = sqrt(si.alpha * sj.alpha);      b.alpha = sqrt(sj.alpha * si.alpha);
      a.r_vdW = 2.0 * sqrt(si.r_vdw * sj.r_vdw); b.r_vdW = 2.0 * sqrt(sj.r_vdw * si.r_vdw);
      a.gamma_w = sqrt(si.gamma_w * sj.gamma_w); b.gamma_w = sqrt(sj.gamma_w * si.gamma_w);
      a.gamma = pow(si.gamma * sj.gamma, -1.5);  b.gamma = pow(sj.gamma * si.gamma, -1.5);
      a.rcore = b.rcore = sqrt(si.rcore2 * sj.rcore2);
      a.ecore = b.ecore = sqrt(si.ecore2 * sj.ecore2);
      a.acore = b.acore = sqrt(si.acore2 * sj.acore2);
      a.lgcij = b.lgcij = sqrt(si.lgcij * sj.lgcij);
      a.lgre = b.lgre = 2.0 * p.gp[35] * sqrt(si.lgre * sj.lgre);
    }
  // off-diagonal, :463-517
  t = line();
  l = toki(t, 0);
  for (int i = 0; i < l; i++) {
    t = line();
    int j = toki(t, 0) - 1, k = toki(t, 1) - 1;
    if (j < nt && k < nt) {
      Tbp &a = p.tb(j, k), &b = p.tb(k, j);
      double v;
      v = tokd(t, 2); if (v > 0.0) a.D = b.D = v;
      v = tokd(t, 3); if (v > 0.0) a.r_vdW = b.r_vdW = 2 * v;
      v = tokd(t, 4); if (v > 0.0) a.alpha = b.alpha = v;
      v = tokd(t, 5); if (v > 0.0) a.r_s = b.r_s = v;
      v = tokd(t, 6); if (v > 0.0) a.r_p = b.r_p = v;
      v = tokd(t, 7); if (v > 0.0) a.r_pp = b.r_pp = v;
      v = tokd(t, 8); if (v >= 0.0) a.lgcij = b.lgcij = v;
    }
  }
  // angles, :519-570 (note the double cnt++ when j==m, restated as-is)
  for (auto& h : p.thbp) h.cnt = 0;
  t = line();
  l = toki(t, 0);
  for (int i = 0; i < l; i++) {
    t = line();
    int j = toki(t, 0) - 1, k = toki(t, 1) - 1, m = toki(t, 2) - 1;
    if (j < nt && k < nt && m < nt) {
      int cnt = p.thb(j, k, m).cnt;
      p.thb(j, k, m).cnt++;
      p.thb(m, k, j).cnt++;
      if (cnt < 5) {
        Thbp &a = p.thb(j, k, m).prm[cnt], &b = p.thb(m, k, j).prm[cnt];
        a.theta_00 = b.theta_00 = tokd(t, 3); a.p_val1 = b.p_val1 = tokd(t, 4);
        a.p_val2 = b.p_val2 = tokd(t, 5);     a.p_coa1 = b.p_coa1 = tokd(t, 6);
        a.p_val7 = b.p_val7 = tokd(t, 7);     a.p_pen1 = b.p_pen1 = tokd(t, 8);
        a.p_val4 = b.p_val4 = tokd(t, 9);
      }
    }
  }
  // torsions, :572-650
  for (auto& h : p.fbp) h.cnt = 0;
  t = line();
  l = toki(t, 0);
  for (int i = 0; i < l; i++) {
    t = line();
    int j = toki(t, 0) - 1, k = toki(t, 1) - 1, m = toki(t, 2) - 1, nn = toki(t, 3) - 1;
    if (j >= 0 && nn >= 0) {
      if (j < nt && k < nt && m < nt && nn < nt) {
        tf(j, k, m, nn) = 1; tf(nn, m, k, j) = 1;
        FbHeader &a = p.fb(j, k, m, nn), &b = p.fb(nn, m, k, j);
        a.cnt = 1; b.cnt = 1;
        a.prm[0].V1 = b.prm[0].V1 = tokd(t, 4); a.prm[0].V2 = b.prm[0].V2 = tokd(t, 5);
        a.prm[0].V3 = b.prm[0].V3 = tokd(t, 6); a.prm[0].p_tor1 = b.prm[0].p_tor1 = tokd(t, 7);
        a.prm[0].p_cot1 = b.prm[0].p_cot1 = tokd(t, 8);
      }
    } else {
      if (k < nt && m < nt)
        for (int pp = 0; pp < nt; pp++)
          for (int o = 0; o < nt; o++) {
            p.fb(pp, k, m, o).cnt = 1;
            p.fb(o, m, k, pp).cnt = 1;
            if (tf(pp, k, m, o) == 0) {
              Fbp& a = p.fb(pp, k, m, o).prm[0];
              a.V1 = tokd(t, 4); a.V2 = tokd(t, 5); a.V3 = tokd(t, 6); a.p_tor1 = tokd(t, 7); a.p_cot1 = tokd(t, 8);
            }
            if (tf(o, m, k, pp) == 0) {
              Fbp& a = p.fb(o, m, k, pp).prm[0];
              a.V1 = tokd(t, 4); a.V2 = tokd(t, 5); a.V3 = tokd(t, 6); a.p_tor1 = tokd(t, 7); a.p_cot1 = tokd(t, 8);
            }
          }
    }
  }
  // hydrogen bonds, :654-686
  t = line();
  l = toki(t, 0);
  for (auto& h : p.hbp) h.r0_hb = -1.0;
  for (int i = 0; i < l; i++) {
    t = line();
    int j = toki(t, 0) - 1, k = toki(t, 1) - 1, m = toki(t, 2) - 1;
    if (j < nt && m < nt) {
      Hbp& h = p.hb(j, k, m);
      h.r0_hb = tokd(t, 3); h.p_hb1 = tokd(t, 4); h.p_hb2 = tokd(t, 5); h.p_hb3 = tokd(t, 6);
    }
  }
  fclose(fp);
  init_taper(p);
  return "";
}

// reaxc_control_sunway.cpp:34-391 (only the keys that reach the force path are stored)
std::string read_control(const char* path, Params& p) {
  p.bond_cut = 5.0; p.bg_cut = 0.3; p.thb_cut = 0.001; p.thb_cutsq = 0.00001; p.hbond_cut = 7.5;
  p.tabulate = 0; p.energy_update_freq = 0;
  if (!path || !strcmp(path, "NULL")) return "";  // pair_reaxc_sunway.cpp:208-232
  FILE* fp = fopen(path, "r");
  if (!fp) return std::string("error opening the control file ") + path;
  static const char* known[] = {
      "simulation_name", "ensemble_type", "nsteps", "dt", "proc_by_dim", "random_vel", "restart_format",
      "restart_freq", "reposition_atoms", "restrict_bonds", "remove_CoM_vel", "debug_level", "energy_update_freq",
      "reneighbor", "vlist_buffer", "nbrhood_cutoff", "bond_graph_cutoff", "thb_cutoff", "thb_cutoff_sq",
      "hbond_cutoff", "ghost_cutoff", "tabulate_long_range", "qeq_freq", "q_err", "ilu_refactor", "ilu_droptol",
      "temp_init", "temp_final", "t_mass", "t_mode", "t_rate", "t_freq", "pressure", "p_mass", "pt_mass",
      "compress", "press_mode", "geo_format", "write_freq", "traj_compress", "traj_method", "traj_title",
      "atom_info", "atom_velocities", "atom_forces", "bond_info", "angle_info", "molecular_analysis", "ignore",
      "dipole_anal", "freq_dipole_anal", "diffusion_coef", "freq_diffusion_coef", "restrict_type", nullptr};
  char s[1024];
  while (fgets(s, sizeof(s), fp)) {
    auto t = tokenize(s);
    if (t.empty()) continue;
    const std::string& k = t[0];
    double v = tokd(t, 1);
    if (k == "nbrhood_cutoff") p.bond_cut = v;
    else if (k == "bond_graph_cutoff") p.bg_cut = v;
    else if (k == "thb_cutoff") p.thb_cut = v;
    else if (k == "thb_cutoff_sq") p.thb_cutsq = v;
    else if (k == "hbond_cutoff") p.hbond_cut = v;
    else if (k == "tabulate_long_range") p.tabulate = (int)atoi(t.size() > 1 ? t[1].c_str() : "0");
    else if (k == "energy_update_freq") p.energy_update_freq = (int)atoi(t.size() > 1 ? t[1].c_str() : "0");
    else {
      bool ok = false;
      for (int i = 0; known[i]; i++) if (k == known[i]) ok = true;
      if (!ok) { fclose(fp); return "WARNING: unknown parameter " + k; }  // :369-372 aborts
    }
  }
  fclose(fp);
  return "";
}

// pair_reaxc_sunway.cpp:318-336
std::string set_element_map(Params& p, int ntypes, const char* const* elements) {
  p.map.assign(ntypes + 1, -1);
  int itmp = 0;
  for (int i = 0; i < ntypes; i++) {
    if (!strcmp(elements[i], "NULL")) { p.map[i + 1] = -1; itmp++; continue; }
  }
  for (int i = 0; i < ntypes; i++)
    for (int j = 0; j < p.nt; j++)
      if (strcasecmp(elements[i], p.sbp[j].name) == 0) { p.map[i + 1] = j; itmp++; }
  if (itmp != ntypes) return "Non-existent ReaxFF type";
  return "";
}

// reaxc_init_md_sunway.cpp:100-136
void init_taper(Params& p) {
  double swa = p.nonb_low, swb = p.nonb_cut;
  double d1 = swb - swa, d7 = pow(d1, 7.0);
  double swa2 = swa * swa, swa3 = swa * swa * swa, swb2 = swb * swb, swb3 = swb * swb * swb;
  p.Tap[7] = 20.0 / d7;
  p.Tap[6] = -70.0 * (swa + swb) / d7;
  p.Tap[5] = 84.0 * (swa2 + 3.0 * swa * swb + swb2) / d7;
  p.Tap[4] = -35.0 * (swa3 + 9.0 * swa2 * swb + 9.0 * swa * swb2 + swb3) / d7;
  p.Tap[3] = 140.0 * (swa3 * swb + 3.0 * swa2 * swb2 + swa * swb3) / d7;
  p.Tap[2] = -210.0 * (swa3 * swb2 + swa2 * swb3) / d7;
  p.Tap[1] = 140.0 * swa3 * swb3 / d7;
  p.Tap[0] = (-35.0 * swa3 * swb2 * swb2 + 21.0 * swa2 * swb3 * swb2 + 7.0 * swa * swb3 * swb3 + swb3 * swb3 * swb) / d7;
}

}  // namespace orc
