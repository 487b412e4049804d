// ffield.reax / control-file front end of the B200 path (host side, C++).
//
// File formats and every default are the reference's, so existing LAMMPS inputs run unchanged:
//   force field : /root/reference/reaxc_ffield_sunway.cpp:35-714  (Read_Force_Field)
//   control     : /root/reference/reaxc_control_sunway.cpp:34-391 (Read_Control_File)
//   tokens      : /root/reference/reaxc_tool_box_sunway.cpp:42-57 (separators TAB SPACE NL CR FF '!' '=')
//   taper       : /root/reference/reaxc_init_md_sunway.cpp:100-136 (Init_Taper)
//   element map : /root/reference/pair_reaxc_sunway.cpp:318-336
// Unlike the reference (MPI_Abort / error->all) errors are returned as strings and surface as negative
// status codes through the C ABI.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <strings.h>

#include "rxb_params.h"

namespace rxb {
namespace {

struct Fields {
  std::vector<std::string> tok;
  size_t count = 0;  // tokens on THIS line; tok may hold more (see LineSource)
  double num(size_t i) const { return i < tok.size() ? atof(tok[i].c_str()) : 0.0; }
  int integer(size_t i) const { return i < tok.size() ? atoi(tok[i].c_str()) : 0; }
  size_t size() const { return count; }
};

Fields split_fields(const std::string& line) {
  static const char* kSep = "\t \n\r\f!=";
  Fields f;
  size_t pos = 0;
  while (pos < line.size()) {
    size_t b = line.find_first_not_of(kSep, pos);
    if (b == std::string::npos) break;
    size_t e = line.find_first_of(kSep, b);
    if (e == std::string::npos) e = line.size();
    f.tok.emplace_back(line.substr(b, e - b));
    pos = e;
  }
  f.count = f.tok.size();
  return f;
}

class LineSource {
 public:
  explicit LineSource(const char* path) : in_(path) {}
  bool ok() const { return in_.good() || in_.eof(); }
  bool opened() const { return in_.is_open(); }
  // Force-field lines are read the way the reference reads them: columns that are absent on a line keep the token a
  // previous, longer line left in that slot (the reference's token buffer is persistent and never cleared,
  // reaxc_tool_box_sunway.cpp:42-57).  Only the unused-without-lgvdw lgcij column of the off-diagonal section depends
  // on it (reaxc_ffield_sunway.cpp:512), but it keeps the parsed tables identical to the reference's, bit for bit.
  Fields next() {
    std::string line;
    Fields now;
    if (std::getline(in_, line)) {
      if (line.size() > 1023) line.resize(1023);  // MAX_LINE of the reference's fgets buffer
      now = split_fields(line);
    }
    if (slots_.size() < now.tok.size()) slots_.resize(now.tok.size());
    for (size_t i = 0; i < now.tok.size(); i++) slots_[i] = now.tok[i];
    Fields f;
    f.tok = slots_;
    f.count = now.count;
    return f;
  }
  void skip(int n) { for (int i = 0; i < n; i++) next(); }

 private:
  std::ifstream in_;
  std::vector<std::string> slots_;
};

template <class T>
void both(T& a, T& b, double T::*member, double v) { a.*member = v; b.*member = v; }

}  // namespace

std::string ForceField::load_control(const char* path) {
  // defaults: reaxc_control_sunway.cpp:48-111 (same values as the "NULL" branch, pair_reaxc_sunway.cpp:208-232)
  ctl.bond_cut = 5.0; ctl.hbond_cut = 7.5; ctl.bg_cut = 0.3; ctl.thb_cut = 0.001; ctl.thb_cutsq = 0.00001;
  ctl.tabulate = 0; ctl.energy_update_freq = 0;
  if (path == nullptr || std::strcmp(path, "NULL") == 0) return "";
  LineSource src(path);
  if (!src.opened()) return std::string("error opening the control file ") + path;
  static const char* kIgnored[] = {
      "simulation_name", "ensemble_type", "nsteps", "dt", "proc_by_dim", "random_vel", "restart_format", "restart_freq",
      "reposition_atoms", "restrict_bonds", "remove_CoM_vel", "debug_level", "reneighbor", "vlist_buffer", "ghost_cutoff",
      "qeq_freq", "q_err", "ilu_refactor", "ilu_droptol", "temp_init", "temp_final", "t_mass", "t_mode", "t_rate", "t_freq",
      "pressure", "p_mass", "pt_mass", "compress", "press_mode", "geo_format", "write_freq", "traj_compress", "traj_method",
      "traj_title", "atom_info", "atom_velocities", "atom_forces", "bond_info", "angle_info", "molecular_analysis", "ignore",
      "dipole_anal", "freq_dipole_anal", "diffusion_coef", "freq_diffusion_coef", "restrict_type"};
  std::ifstream in(path);
  std::string line;
  while (std::getline(in, line)) {
    Fields f = split_fields(line);
    if (f.size() == 0) continue;
    const std::string& key = f.tok[0];
    if (key == "nbrhood_cutoff") ctl.bond_cut = f.num(1);
    else if (key == "hbond_cutoff") ctl.hbond_cut = f.num(1);
    else if (key == "bond_graph_cutoff") ctl.bg_cut = f.num(1);
    else if (key == "thb_cutoff") ctl.thb_cut = f.num(1);
    else if (key == "thb_cutoff_sq") ctl.thb_cutsq = f.num(1);
    else if (key == "tabulate_long_range") ctl.tabulate = f.integer(1);
    else if (key == "energy_update_freq") ctl.energy_update_freq = f.integer(1);
    else {
      bool known = false;
      for (const char* k : kIgnored) known = known || key == k;
      if (!known) return "unknown parameter " + key + " in control file";  // reference aborts, :369-372
    }
  }
  return "";
}

std::string ForceField::load_ffield(const char* path) {
  LineSource src(path);
  if (!src.opened()) return std::string("Cannot open ReaxFF potential file ") + path;
  src.skip(1);
  const int nglob = src.next().integer(0);
  if (nglob < 1) return "number of globals in ffield file is 0";
  gp.resize(nglob);
  for (double& g : gp) g = src.next().num(0);
  ctl.bo_cut = 0.01 * gp[29];
  ctl.nonb_low = gp[11];
  ctl.nonb_cut = gp[12];

  nt = src.next().integer(0);
  src.skip(3);
  atom.assign(nt, AtomPar{});
  names.assign(nt, "");
  pair.assign((size_t)nt * nt, PairPar{});
  angle.assign((size_t)nt * nt * nt, AngleSet{});
  tors.assign((size_t)nt * nt * nt * nt, TorsPar{});
  hb.assign((size_t)nt * nt * nt, HbPar{});
  ctl.vdw_type = 0;

  // ---- atoms: 4 lines each (5 with lgvdw) ----
  for (int e = 0; e < nt; e++) {
    AtomPar& a = atom[e];
    Fields f = src.next();
    if (f.size() == 0) return "Inconsistent ffield file (atom section)";
    for (char c : f.tok[0]) names[e].push_back((char)toupper(c));
    a.is_carbon = names[e] == "C";
    a.r_s = f.num(1); a.valency = f.num(2); a.mass = f.num(3); a.r_vdw = f.num(4); a.epsilon = f.num(5);
    a.gamma = f.num(6); a.r_pi = f.num(7); a.valency_e = f.num(8);
    a.nlp_opt = 0.5 * (a.valency_e - a.valency);
    f = src.next();
    a.alpha = f.num(0); a.gamma_w = f.num(1); a.valency_boc = f.num(2); a.p_ovun5 = f.num(3);
    a.chi = f.num(5); a.eta = 2.0 * f.num(6); a.p_hbond = (int)f.num(7);
    f = src.next();
    a.r_pi_pi = f.num(0); a.p_lp2 = f.num(1); a.b_o_131 = f.num(3); a.b_o_132 = f.num(4); a.b_o_133 = f.num(5);
    f = src.next();
    if (f.size() < 3) return "Inconsistent ffield file (reaxc_ffield.cpp)";
    a.p_ovun2 = f.num(0); a.p_val3 = f.num(1); a.valency_val = f.num(3); a.p_val5 = f.num(4);
    a.rcore2 = f.num(5); a.ecore2 = f.num(6); a.acore2 = f.num(7);
    if (ctl.lgflag) {
      f = src.next();
      if (f.size() > 3) return "Inconsistent ffield file (reaxc_ffield.cpp)";
      a.lgcij = f.num(0); a.lgre = f.num(1);
    }
    // van der Waals flavour: 1 shielding, 2 inner wall, 3 both (first element decides, later ones only warn)
    const bool wall = a.rcore2 > 0.01 && a.acore2 > 0.01;
    const bool shield = a.gamma_w > 0.5;
    if (!wall && !shield) return "inconsistent vdWaals-parameters: no shielding or inner-wall set for element " + names[e];
    const int want = wall ? (shield ? 3 : 2) : 1;
    if (ctl.vdw_type == 0 || ctl.vdw_type == want) ctl.vdw_type = want;
  }
  for (AtomPar& a : atom)
    if (a.mass < 21 && a.valency_val != a.valency_boc) a.valency_val = a.valency_boc;

  auto P = [&](int i, int j) -> PairPar& { return pair[(size_t)i * nt + j]; };

  // ---- bonds: 2 lines each ----
  int count = src.next().integer(0);
  src.skip(1);
  for (int b = 0; b < count; b++) {
    Fields f = src.next();
    const int i = f.integer(0) - 1, j = f.integer(1) - 1;
    if (i >= nt || j >= nt) continue;  // the reference leaves the second line to be mis-read as the next entry
    PairPar &x = P(i, j), &y = P(j, i);
    both(x, y, &PairPar::De_s, f.num(2));   both(x, y, &PairPar::De_p, f.num(3));   both(x, y, &PairPar::De_pp, f.num(4));
    both(x, y, &PairPar::p_be1, f.num(5));  both(x, y, &PairPar::p_bo5, f.num(6));  both(x, y, &PairPar::v13cor, f.num(7));
    both(x, y, &PairPar::p_bo6, f.num(8));  both(x, y, &PairPar::p_ovun1, f.num(9));
    f = src.next();
    both(x, y, &PairPar::p_be2, f.num(0));  both(x, y, &PairPar::p_bo3, f.num(1));  both(x, y, &PairPar::p_bo4, f.num(2));
    both(x, y, &PairPar::p_bo1, f.num(4));  both(x, y, &PairPar::p_bo2, f.num(5));  both(x, y, &PairPar::ovc, f.num(6));
  }
  // ---- combination rules ----
  for (int i = 0; i < nt; i++)
    for (int j = 0; j < nt; j++) {
      const AtomPar &a = atom[i], &b = atom[j];
      PairPar& x = P(i, j);
      x.r_s = 0.5 * (a.r_s + b.r_s);
      x.r_p = 0.5 * (a.r_pi + b.r_pi);
      x.r_pp = 0.5 * (a.r_pi_pi + b.r_pi_pi);
      x.p_boc3 = sqrt(a.b_o_132 * b.b_o_132);
      x.p_boc4 = sqrt(a.b_o_131 * b.b_o_131);
      x.p_boc5 = sqrt(a.b_o_133 * b.b_o_133);
      x.D = sqrt(a.epsilon * b.epsilon);
      x.alpha = sqrt(a.alpha * b.alpha);
      x.r_vdW = 2.0 * sqrt(a.r_vdw * b.r_vdw);
      x.gamma_w = sqrt(a.gamma_w * b.gamma_w);
      x.gamma = pow(a.gamma * b.gamma, -1.5);
      x.rcore = sqrt(a.rcore2 * b.rcore2);
      x.ecore = sqrt(a.ecore2 * b.ecore2);
      x.acore = sqrt(a.acore2 * b.acore2);
      x.lgcij = sqrt(a.lgcij * b.lgcij);
      x.lgre = 2.0 * gp[35] * sqrt(a.lgre * b.lgre);
    }
  // ---- off-diagonal overrides ----
  count = src.next().integer(0);
  for (int b = 0; b < count; b++) {
    Fields f = src.next();
    const int i = f.integer(0) - 1, j = f.integer(1) - 1;
    if (i >= nt || j >= nt) continue;
    PairPar &x = P(i, j), &y = P(j, i);
    if (f.num(2) > 0.0) both(x, y, &PairPar::D, f.num(2));
    if (f.num(3) > 0.0) both(x, y, &PairPar::r_vdW, 2 * f.num(3));
    if (f.num(4) > 0.0) both(x, y, &PairPar::alpha, f.num(4));
    if (f.num(5) > 0.0) both(x, y, &PairPar::r_s, f.num(5));
    if (f.num(6) > 0.0) both(x, y, &PairPar::r_p, f.num(6));
    if (f.num(7) > 0.0) both(x, y, &PairPar::r_pp, f.num(7));
    if (f.num(8) >= 0.0) both(x, y, &PairPar::lgcij, f.num(8));
  }
  // ---- valence angles ----
  auto A = [&](int i, int j, int k) -> AngleSet& { return angle[((size_t)i * nt + j) * nt + k]; };
  count = src.next().integer(0);
  for (int b = 0; b < count; b++) {
    Fields f = src.next();
    const int i = f.integer(0) - 1, j = f.integer(1) - 1, k = f.integer(2) - 1;
    if (i >= nt || j >= nt || k >= nt) continue;
    const int slot = A(i, j, k).cnt;
    A(i, j, k).cnt++;
    A(k, j, i).cnt++;  // i==k bumps the same counter twice, leaving an all-zero slot (reference behaviour, :541-543)
    if (slot >= kMaxAngleSets) continue;
    AnglePar v{f.num(3), f.num(4), f.num(5), f.num(6), f.num(7), f.num(8), f.num(9)};
    A(i, j, k).prm[slot] = v;
    A(k, j, i).prm[slot] = v;
  }
  // ---- torsions (explicit entries win over 0-j-k-0 wildcards regardless of file order) ----
  auto T = [&](int i, int j, int k, int l) -> TorsPar& { return tors[(((size_t)i * nt + j) * nt + k) * nt + l]; };
  std::vector<char> explicit_entry(tors.size(), 0);
  auto X = [&](int i, int j, int k, int l) -> char& { return explicit_entry[(((size_t)i * nt + j) * nt + k) * nt + l]; };
  count = src.next().integer(0);
  for (int b = 0; b < count; b++) {
    Fields f = src.next();
    const int i = f.integer(0) - 1, j = f.integer(1) - 1, k = f.integer(2) - 1, l = f.integer(3) - 1;
    TorsPar v{f.num(4), f.num(5), f.num(6), f.num(7), f.num(8), 1, 0};
    if (i >= 0 && l >= 0) {
      if (i < nt && j < nt && k < nt && l < nt) {
        X(i, j, k, l) = X(l, k, j, i) = 1;
        T(i, j, k, l) = v;
        T(l, k, j, i) = v;
      }
    } else if (j < nt && k < nt) {
      for (int p = 0; p < nt; p++)
        for (int o = 0; o < nt; o++) {
          T(p, j, k, o).cnt = 1;
          T(o, k, j, p).cnt = 1;
          if (!X(p, j, k, o)) T(p, j, k, o) = v;
          if (!X(o, k, j, p)) T(o, k, j, p) = v;
        }
    }
  }
  // ---- hydrogen bonds ----
  for (HbPar& h : hb) h.r0_hb = -1.0;
  count = src.next().integer(0);
  for (int b = 0; b < count; b++) {
    Fields f = src.next();
    const int i = f.integer(0) - 1, j = f.integer(1) - 1, k = f.integer(2) - 1;
    if (i < nt && k < nt && i >= 0 && j >= 0 && k >= 0 && j < nt)
      hb[((size_t)i * nt + j) * nt + k] = HbPar{f.num(3), f.num(4), f.num(5), f.num(6)};
  }
  derive();
  return "";
}

std::string ForceField::set_elements(int ntypes, const char* const* el) {
  map.assign(ntypes + 1, -1);
  int matched = 0;
  for (int t = 0; t < ntypes; t++) {
    if (std::strcmp(el[t], "NULL") == 0) { matched++; continue; }
    for (int e = 0; e < nt; e++)
      if (strcasecmp(el[t], names[e].c_str()) == 0) { map[t + 1] = e; matched++; }
  }
  if (matched != ntypes) return "Non-existent ReaxFF type";
  bool any = false;
  for (int t = 1; t <= ntypes; t++) any = any || map[t] >= 0;
  if (!any) return "Incorrect args for pair coefficients";
  return "";
}

void ForceField::derive() {
  const double swa = ctl.nonb_low, swb = ctl.nonb_cut;
  const double d7 = pow(swb - swa, 7.0);
  const double a2 = swa * swa, a3 = a2 * swa, b2 = swb * swb, b3 = b2 * swb;
  double* T = ctl.Tap;
  T[7] = 20.0 / d7;
  T[6] = -70.0 * (swa + swb) / d7;
  T[5] = 84.0 * (a2 + 3.0 * swa * swb + b2) / d7;
  T[4] = -35.0 * (a3 + 9.0 * a2 * swb + 9.0 * swa * b2 + b3) / d7;
  T[3] = 140.0 * (a3 * swb + 3.0 * a2 * b2 + swa * b3) / d7;
  T[2] = -210.0 * (a3 * b2 + a2 * b3) / d7;
  T[1] = 140.0 * a3 * b3 / d7;
  T[0] = (-35.0 * a3 * b2 * b2 + 21.0 * a2 * b3 * b2 + 7.0 * swa * b3 * b3 + b3 * b3 * swb) / d7;
  const double p_vdW1 = gp.size() > 28 ? gp[28] : 0.0;
  for (PairPar& x : pair) {
    x.powgi_vdW1 = (x.gamma_w > 0.0) ? pow(1.0 / x.gamma_w, p_vdW1) : 0.0;
    x.inv_r_vdW = x.r_vdW != 0.0 ? 1.0 / x.r_vdW : 0.0;
    x.alpha_over_r_vdW = x.r_vdW != 0.0 ? x.alpha / x.r_vdW : 0.0;
    x.log_r_s = x.r_s > 0.0 ? log(x.r_s) : 0.0;
    x.log_r_p = x.r_p > 0.0 ? log(x.r_p) : 0.0;
    x.log_r_pp = x.r_pp > 0.0 ? log(x.r_pp) : 0.0;
  }
  // Reach of a bond per element pair: the uncorrected bond order of BOp_single (reaxc_forces_sunway.cpp:694-720) is a sum
  // of terms exp(p (d/r)^q) with p < 0 < q, i.e. it falls monotonically with d, so the distance where it crosses bo_cut
  // bounds every bond of that pair.  k_bond_list uses it (plus a 1e-6 relative margin) to keep candidates that cannot
  // bond away from the transcendental work; the accept test itself stays the exact one.
  for (int i = 0; i < nt; i++)
    for (int j = 0; j < nt; j++) {
      PairPar& x = pair[(size_t)i * nt + j];
      const AtomPar &ai = atom[i], &aj = atom[j];
      const bool use_s = ai.r_s > 0.0 && aj.r_s > 0.0, use_p = ai.r_pi > 0.0 && aj.r_pi > 0.0,
                 use_pp = ai.r_pi_pi > 0.0 && aj.r_pi_pi > 0.0;
      const bool monotone = (!use_s || (x.p_bo1 < 0.0 && x.p_bo2 > 0.0)) && (!use_p || (x.p_bo3 < 0.0 && x.p_bo4 > 0.0)) &&
                            (!use_pp || (x.p_bo5 < 0.0 && x.p_bo6 > 0.0));
      auto bop = [&](double d) {
        double b = 0.0;
        if (use_s) b += (1.0 + ctl.bo_cut) * exp(x.p_bo1 * pow(d / x.r_s, x.p_bo2));
        if (use_p) b += exp(x.p_bo3 * pow(d / x.r_p, x.p_bo4));
        if (use_pp) b += exp(x.p_bo5 * pow(d / x.r_pp, x.p_bo6));
        return b;
      };
      double reach = ctl.bond_cut;
      if (monotone && bop(ctl.bond_cut) < ctl.bo_cut) {
        double lo = 0.0, hi = ctl.bond_cut;          // bop(lo) >= bo_cut > bop(hi)
        for (int it = 0; it < 200; it++) {
          const double mid = 0.5 * (lo + hi);
          if (bop(mid) >= ctl.bo_cut) lo = mid; else hi = mid;
        }
        reach = std::min(ctl.bond_cut, hi * (1.0 + 1e-6));
      }
      x.d_bond_max = reach;
    }
}

std::vector<double> ForceField::dump() const {
  std::vector<double> v;
  auto put = [&](std::initializer_list<double> l) { v.insert(v.end(), l); };
  put({(double)nt, (double)ctl.vdw_type, (double)gp.size()});
  v.insert(v.end(), gp.begin(), gp.end());
  put({ctl.bo_cut, ctl.nonb_low, ctl.nonb_cut, ctl.bond_cut, ctl.hbond_cut, ctl.bg_cut, ctl.thb_cut, ctl.thb_cutsq,
       (double)ctl.tabulate, (double)ctl.energy_update_freq});
  v.insert(v.end(), ctl.Tap, ctl.Tap + 8);
  for (const AtomPar& a : atom)
    put({a.r_s, a.valency, a.mass, a.r_vdw, a.epsilon, a.gamma, a.r_pi, a.valency_e, a.nlp_opt, a.alpha, a.gamma_w,
         a.valency_boc, a.p_ovun5, a.chi, a.eta, (double)a.p_hbond, a.r_pi_pi, a.p_lp2, a.b_o_131, a.b_o_132, a.b_o_133,
         a.p_ovun2, a.p_val3, a.valency_val, a.p_val5, a.rcore2, a.ecore2, a.acore2, a.lgcij, a.lgre});
  for (const PairPar& t : pair)
    put({t.p_bo1, t.p_bo2, t.p_bo3, t.p_bo4, t.p_bo5, t.p_bo6, t.r_s, t.r_p, t.r_pp, t.p_boc3, t.p_boc4, t.p_boc5, t.p_be1,
         t.p_be2, t.De_s, t.De_p, t.De_pp, t.p_ovun1, t.D, t.alpha, t.r_vdW, t.gamma_w, t.rcore, t.ecore, t.acore, t.lgcij,
         t.lgre, t.gamma, t.v13cor, t.ovc});
  for (const AngleSet& s : angle) {
    v.push_back(s.cnt);
    for (const AnglePar& q : s.prm) put({q.theta_00, q.p_val1, q.p_val2, q.p_coa1, q.p_val7, q.p_pen1, q.p_val4});
  }
  for (const TorsPar& q : tors) put({(double)q.cnt, q.V1, q.V2, q.V3, q.p_tor1, q.p_cot1});
  for (const HbPar& h : hb) put({h.r0_hb, h.p_hb1, h.p_hb2, h.p_hb3});
  for (size_t t = 1; t < map.size(); t++) v.push_back(map[t]);
  return v;
}

}  // namespace rxb
