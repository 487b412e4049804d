// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under sw_reaxff_b200/ may include, link or call this.
//
// CPU restatement (double, libm) of the serial semantics of SW_REAXFF's force path.  Each routine
// cites the /root/reference range it follows.  Deviations from the reference, all deliberate and
// listed in SURVEY.md §8a:  (1) e_vdW/e_ele take 1/2 per directed pair (stock-correct) instead of
// the reference's double-counted my_en (reaxc_nonbonded_cpe.h:314,346);  (2) e_pol is the plain sum,
// not the running-prefix sum of reaxc_multi_body_sw64.c:102-103.
//
// Pinning (tests/test_oracle_vs_ref.py, against the reference's own code compiled unmodified into oracle/_ref/libref.so):
//   build_bond_list (BO') == BOp_single;  valence_torsion == Torsion_Angles (serial virial path);
//   hydrogen_bonds == Hydrogen_Bonds (serial virial path);  add_dbond_forces == Add_dBond_to_Forces;
//   bonds_atom_energy == Merge_Bonds_Atom_Energy_C_New;  nonbonded == vdW_Coulomb_Energy_Full_C_test_err  — all to 1e-10.
//   bond_orders == the serial body of BO() (reaxc_bond_orders_sunway.cpp:460-774, reached through
//   oracle/ref/stubs/prelude_bo_serial.h) to 1e-11.
//   PARITY UNPINNED by reference execution: only the tabulated mode (a9', commented out in the reference); its deviation
//   from the analytic form is measured in tests/test_oracle.py.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "orc_system.h"

namespace orc {

static const double C_ele = 332.06371;         // reaxc_defs_sunway.h:62
static const double constPI = 3.14159265;      // reaxc_defs_sunway.h:55 (8 digits, as in the reference)
static const double KCALpMOL_to_EV = 23.02;    // reaxc_defs_sunway.h:69
static const double HB_THRESHOLD = 1e-2;       // reaxc_defs_sunway.h:95
static const double MIN_SINE = 1e-10;          // reaxc_torsion_angles_sunway.cpp:40
static inline double SQR(double a) { return a * a; }
static inline double CUBE(double a) { return a * a * a; }
static inline double DEG2RAD(double a) { return a * constPI / 180.0; }
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

#ifdef _OPENMP
#define OMP_ATOMIC _Pragma("omp atomic")
#else
#define OMP_ATOMIC
#endif
static inline void atomic_add(double& dst, double v) {
  OMP_ATOMIC
  dst += v;
}
static inline void add4(double* fc, double c, const double* v) {  // rvec4_ScaledAdd on the xyz part
  atomic_add(fc[0], c * v[0]);
  atomic_add(fc[1], c * v[1]);
  atomic_add(fc[2], c * v[2]);
}

// ---------------------------------------------------------------------------------------------
// a4: Init_Forces_noQEq_Full + BOp_single   reaxc_forces_sunway.cpp:677-825, 852-935
//     far-list semantics (d = sqrt(r2), dvec = x_j - x_i): pair_reaxc_sw64.c:49-95
//     rows sorted by neighbour index: reaxc_forces_sw64.c:31-75,1289
void build_bond_list(System& s) {
  const Params& P = s.prm;
  const int N = s.N;
  s.b_start.assign(N, 0);
  s.b_end.assign(N, 0);
  s.total_bo.assign(N, 0.0);
  s.dDeltap_self.assign((size_t)N * 3, 0.0);
  std::vector<std::vector<Bond>> rows(N);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < N; i++) {
    int ti = s.type[i];
    if (ti < 0) continue;
    const Sbp& sbi = P.sbp[ti];
    std::vector<Bond>& row = rows[i];
    for (long pj = s.nb_off[i]; pj < s.nb_off[i + 1]; pj++) {
      int j = s.nb[pj];
      double dv[3] = {s.x[3 * j] - s.x[3 * i], s.x[3 * j + 1] - s.x[3 * i + 1], s.x[3 * j + 2] - s.x[3 * i + 2]};
      double r2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
      if (!(r2 <= P.nonb_cut * P.nonb_cut)) continue;  // far list filter
      double d = sqrt(r2);
      if (!(d <= P.bond_cut)) continue;
      int tj = s.type[j];
      if (tj < 0) continue;
      const Sbp& sbj = P.sbp[tj];
      const Tbp& tw = P.tb(ti, tj);
      double C12, C34, C56, BO_s, BO_pi, BO_pi2;
      double rr2 = SQR(d);
      if (sbi.r_s > 0.0 && sbj.r_s > 0.0) {
        C12 = tw.p_bo1 * pow(d / tw.r_s, tw.p_bo2);
        BO_s = (1.0 + P.bo_cut) * exp(C12);
      } else BO_s = C12 = 0.0;
      if (sbi.r_pi > 0.0 && sbj.r_pi > 0.0) {
        C34 = tw.p_bo3 * pow(d / tw.r_p, tw.p_bo4);
        BO_pi = exp(C34);
      } else BO_pi = C34 = 0.0;
      if (sbi.r_pi_pi > 0.0 && sbj.r_pi_pi > 0.0) {
        C56 = tw.p_bo5 * pow(d / tw.r_pp, tw.p_bo6);
        BO_pi2 = exp(C56);
      } else BO_pi2 = C56 = 0.0;
      double BO = BO_s + BO_pi + BO_pi2;
      if (BO >= P.bo_cut) {
        Bond b;
        memset(&b, 0, sizeof(b));
        b.nbr = j; b.sym = -1; b.d = d;
        b.dvec[0] = dv[0]; b.dvec[1] = dv[1]; b.dvec[2] = dv[2];
        b.BO = BO; b.BO_s = BO_s; b.BO_pi = BO_pi; b.BO_pi2 = BO_pi2;
        double Cln_BOp_s = tw.p_bo2 * C12 / rr2;
        double Cln_BOp_pi = tw.p_bo4 * C34 / rr2;
        double Cln_BOp_pi2 = tw.p_bo6 * C56 / rr2;
        for (int t = 0; t < 3; t++) {
          b.dln_BOp_s[t] = -b.BO_s * Cln_BOp_s * dv[t];
          b.dln_BOp_pi[t] = -b.BO_pi * Cln_BOp_pi * dv[t];
          b.dln_BOp_pi2[t] = -b.BO_pi2 * Cln_BOp_pi2 * dv[t];
          b.dBOp[t] = -(b.BO_s * Cln_BOp_s + b.BO_pi * Cln_BOp_pi + b.BO_pi2 * Cln_BOp_pi2) * dv[t];
        }
        b.BO_s -= P.bo_cut;
        b.BO -= P.bo_cut;
        b.Cdbo = b.Cdbopi = b.Cdbopi2 = 0.0;
        row.push_back(b);
      }
    }
    std::sort(row.begin(), row.end(), [](const Bond& a, const Bond& b) { return a.nbr < b.nbr; });
    // accumulate in row order (after the sort the order is deterministic)
    for (const Bond& b : row) {
      for (int t = 0; t < 3; t++) s.dDeltap_self[3 * i + t] += b.dBOp[t];
      s.total_bo[i] += b.BO;
    }
  }
  size_t tot = 0;
  for (int i = 0; i < N; i++) { s.b_start[i] = (int)tot; tot += rows[i].size(); s.b_end[i] = (int)tot; }
  s.bonds.resize(tot);
  for (int i = 0; i < N; i++) std::copy(rows[i].begin(), rows[i].end(), s.bonds.begin() + s.b_start[i]);
}

// a5: Init_Forces_noQEq_HB_Full_C  reaxc_forces_sw64.c:787-863;  Hindex: reaxc_reset_tools_sunway.cpp:36-57
void build_hbond_list(System& s) {
  const Params& P = s.prm;
  s.Hindex.assign(s.N, -1);
  s.hb_start.clear(); s.hb_end.clear(); s.hbonds.clear();
  if (P.hbond_cut <= 0) return;
  int numH = 0;
  for (int i = 0; i < s.n; i++) {
    if (s.type[i] < 0) continue;
    s.Hindex[i] = (P.sbp[s.type[i]].p_hbond == 1) ? numH++ : -1;
  }
  s.hb_start.assign(numH, 0); s.hb_end.assign(numH, 0);
  for (int i = 0; i < s.n; i++) {
    int ti = s.type[i];
    if (ti < 0) continue;
    if (P.sbp[ti].p_hbond != 1) continue;
    int h = s.Hindex[i];
    s.hb_start[h] = (int)s.hbonds.size();
    for (long pj = s.nb_off[i]; pj < s.nb_off[i + 1]; pj++) {
      int j = s.nb[pj];
      double dv[3] = {s.x[3 * j] - s.x[3 * i], s.x[3 * j + 1] - s.x[3 * i + 1], s.x[3 * j + 2] - s.x[3 * i + 2]};
      double r2 = dot3(dv, dv);
      if (!(r2 <= P.nonb_cut * P.nonb_cut)) continue;
      double d = sqrt(r2);
      if (!(d <= P.hbond_cut)) continue;
      int tj = s.type[j];
      if (tj < 0) continue;
      if (P.sbp[tj].p_hbond == 2) s.hbonds.push_back(HBond{j, d, {dv[0], dv[1], dv[2]}});
    }
    s.hb_end[h] = (int)s.hbonds.size();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// a9' / f4: tabulated long-range interactions.  NOT executed by the reference (reaxc_lookup_sunway.cpp is commented out in
// full and Init_Lookup_Tables is never called); restated from those comments because BASELINE's north_star names the
// spline tables: Tridiagonal_Solve :37-54, Natural_Cubic_Spline :57-106, Complete_Cubic_Spline :110-157,
// Init_Lookup_Tables :157-285 (including its quirks: v0 = CEvd / CEclmb at the first knot, vlast_ele = fele[last],
// the operator-precedence slip in d[n-1]), LR_vdW_Coulomb reaxc_nonbonded_sunway.cpp:573-668, and the evaluation form
// of Tabulated_vdW_Coulomb_Energy :498-519.  Parity unpinned (dead code); the deviation from the analytic form is measured
// in tests/test_oracle.py.
namespace {
struct LRpoint { double H, e_vdW, CEvd, e_ele, CEclmb; };

LRpoint lr_vdw_coulomb(const Params& P, int i, int j, double r_ij) {
  const double p_vdW1 = P.gp[28], p_vdW1i = 1.0 / p_vdW1;
  const double* Tap = P.Tap;
  const Tbp& tw = P.tb(i, j);
  LRpoint lr;
  double T = Tap[7] * r_ij + Tap[6];
  T = T * r_ij + Tap[5]; T = T * r_ij + Tap[4]; T = T * r_ij + Tap[3];
  T = T * r_ij + Tap[2]; T = T * r_ij + Tap[1]; T = T * r_ij + Tap[0];
  double dT = 7 * Tap[7] * r_ij + 6 * Tap[6];
  dT = dT * r_ij + 5 * Tap[5]; dT = dT * r_ij + 4 * Tap[4]; dT = dT * r_ij + 3 * Tap[3];
  dT = dT * r_ij + 2 * Tap[2];
  dT += Tap[1] / r_ij;
  if (P.vdw_type == 1 || P.vdw_type == 3) {
    double powr = pow(r_ij, p_vdW1);
    double powgi = pow(1.0 / tw.gamma_w, p_vdW1);
    double fn13 = pow(powr + powgi, p_vdW1i);
    double exp1 = exp(tw.alpha * (1.0 - fn13 / tw.r_vdW));
    double exp2 = exp(0.5 * tw.alpha * (1.0 - fn13 / tw.r_vdW));
    lr.e_vdW = T * tw.D * (exp1 - 2.0 * exp2);
    double dfn13 = pow(powr + powgi, p_vdW1i - 1.0) * pow(r_ij, p_vdW1 - 2.0);
    lr.CEvd = dT * tw.D * (exp1 - 2.0 * exp2) - T * tw.D * (tw.alpha / tw.r_vdW) * (exp1 - exp2) * dfn13;
  } else {
    double exp1 = exp(tw.alpha * (1.0 - r_ij / tw.r_vdW));
    double exp2 = exp(0.5 * tw.alpha * (1.0 - r_ij / tw.r_vdW));
    lr.e_vdW = T * tw.D * (exp1 - 2.0 * exp2);
    lr.CEvd = dT * tw.D * (exp1 - 2.0 * exp2) - T * tw.D * (tw.alpha / tw.r_vdW) * (exp1 - exp2) / r_ij;
  }
  if (P.vdw_type == 2 || P.vdw_type == 3) {
    double e_core = tw.ecore * exp(tw.acore * (1.0 - (r_ij / tw.rcore)));
    lr.e_vdW += T * e_core;
    double de_core = -(tw.acore / tw.rcore) * e_core;
    lr.CEvd += dT * e_core + T * de_core / r_ij;
    if (P.lgflag) {
      double r5 = pow(r_ij, 5.0), r6 = pow(r_ij, 6.0), re6 = pow(tw.lgre, 6.0);
      double e_lg = -(tw.lgcij / (r6 + re6));
      lr.e_vdW += T * e_lg;
      double de_lg = -6.0 * e_lg * r5 / (r6 + re6);
      lr.CEvd += dT * e_lg + T * de_lg / r_ij;
    }
  }
  double dr3gamij_1 = r_ij * r_ij * r_ij + tw.gamma;
  double dr3gamij_3 = pow(dr3gamij_1, 0.33333333333333);
  double tmp = T / dr3gamij_3;
  lr.H = 14.40 * tmp;   // EV_to_KCALpMOL
  lr.e_ele = C_ele * tmp;
  lr.CEclmb = C_ele * (dT - T * r_ij / dr3gamij_1) / dr3gamij_3;
  return lr;
}

void tridiagonal_solve(const double* a, const double* b, double* c, double* d, double* x, int n) {
  c[0] /= b[0];
  d[0] /= b[0];
  for (int i = 1; i < n; i++) {
    double id = (b[i] - c[i - 1] * a[i]);
    c[i] /= id;
    d[i] = (d[i] - d[i - 1] * a[i]) / id;
  }
  x[n - 1] = d[n - 1];
  for (int i = n - 2; i >= 0; i--) x[i] = d[i] - c[i] * x[i + 1];
}

void spline_coefs(const double* h, const double* f, const double* v, SplineCoef* coef, int n) {
  for (int i = 1; i < n; ++i) {
    coef[i - 1].d = (v[i] - v[i - 1]) / (6 * h[i - 1]);
    coef[i - 1].c = v[i] / 2;
    coef[i - 1].b = (f[i] - f[i - 1]) / h[i - 1] + h[i - 1] * (2 * v[i] + v[i - 1]) / 6;
    coef[i - 1].a = f[i];
  }
}

void natural_cubic_spline(const double* h, const double* f, SplineCoef* coef, int n) {
  std::vector<double> a(n), b(n), c(n), d(n), v(n);
  a[0] = a[1] = a[n - 1] = 0;
  for (int i = 2; i < n - 1; ++i) a[i] = h[i - 1];
  b[0] = b[n - 1] = 0;
  for (int i = 1; i < n - 1; ++i) b[i] = 2 * (h[i - 1] + h[i]);
  c[0] = c[n - 2] = c[n - 1] = 0;
  for (int i = 1; i < n - 2; ++i) c[i] = h[i];
  d[0] = d[n - 1] = 0;
  for (int i = 1; i < n - 1; ++i) d[i] = 6 * ((f[i + 1] - f[i]) / h[i] - (f[i] - f[i - 1]) / h[i - 1]);
  v[0] = 0;
  v[n - 1] = 0;
  tridiagonal_solve(&a[1], &b[1], &c[1], &d[1], &v[1], n - 2);
  spline_coefs(h, f, v.data(), coef, n);
}

void complete_cubic_spline(const double* h, const double* f, double v0, double vlast, SplineCoef* coef, int n) {
  std::vector<double> a(n), b(n), c(n), d(n), v(n);
  a[0] = 0;
  for (int i = 1; i < n; ++i) a[i] = h[i - 1];
  b[0] = 2 * h[0];
  for (int i = 1; i < n; ++i) b[i] = 2 * (h[i - 1] + h[i]);
  c[n - 1] = 0;
  for (int i = 0; i < n - 1; ++i) c[i] = h[i];
  d[0] = 6 * (f[1] - f[0]) / h[0] - 6 * v0;
  d[n - 1] = 6 * vlast - 6 * (f[n - 1] - f[n - 2] / h[n - 2]);   // sic (:143)
  for (int i = 1; i < n - 1; ++i) d[i] = 6 * ((f[i + 1] - f[i]) / h[i] - (f[i] - f[i - 1]) / h[i - 1]);
  tridiagonal_solve(&a[0], &b[0], &c[0], &d[0], &v[0], n);
  spline_coefs(h, f, v.data(), coef, n);
}
}  // namespace

void build_lookup_tables(Params& P) {
  Lookup& L = P.lookup;
  const int nt = P.nt, tab = P.tabulate;
  L.n = tab + 2;
  L.dx = P.nonb_cut / tab;
  L.inv_dx = tab / P.nonb_cut;
  L.tables.assign((size_t)nt * nt * 5 * L.n, SplineCoef{0, 0, 0, 0});
  // h[r] spans r = 1 .. tab+1 (the reference sets h[tab+1] too); one spare slot because Complete_Cubic_Spline reads h[n-1]
  std::vector<double> h(tab + 3, L.dx), fh(tab + 2), fvdw(tab + 2), fCEvd(tab + 2), fele(tab + 2), fCEclmb(tab + 2);
  for (int i = 0; i < nt; i++)
    for (int j = i; j < nt; j++) {
      LRpoint first{};
      int r;
      for (r = 1; r <= tab; ++r) {
        LRpoint y = lr_vdw_coulomb(P, i, j, r * L.dx);
        if (r == 1) first = y;
        fh[r] = y.H; fvdw[r] = y.e_vdW; fCEvd[r] = y.CEvd; fele[r] = y.e_ele; fCEclmb[r] = y.CEclmb;
      }
      const double v0_vdw = first.CEvd, v0_ele = first.CEclmb;
      fh[r] = fh[r - 1]; fvdw[r] = fvdw[r - 1]; fCEvd[r] = fCEvd[r - 1]; fele[r] = fele[r - 1]; fCEclmb[r] = fCEclmb[r - 1];
      const double vlast_vdw = fCEvd[r - 1], vlast_ele = fele[r - 1];
      natural_cubic_spline(&h[1], &fh[1], L.at(nt, i, j, 0) + 1, tab + 1);
      complete_cubic_spline(&h[1], &fvdw[1], v0_vdw, vlast_vdw, L.at(nt, i, j, 1) + 1, tab + 1);
      natural_cubic_spline(&h[1], &fCEvd[1], L.at(nt, i, j, 2) + 1, tab + 1);
      complete_cubic_spline(&h[1], &fele[1], v0_ele, vlast_ele, L.at(nt, i, j, 3) + 1, tab + 1);
      natural_cubic_spline(&h[1], &fCEclmb[1], L.at(nt, i, j, 4) + 1, tab + 1);
    }
}

// Tabulated_vdW_Coulomb_Energy evaluation (:498-519) in the production full-list / owner-computes form of a9
static void nonbonded_tabulated(System& s) {
  Params& P = s.prm;
  if (P.lookup.n != P.tabulate + 2) build_lookup_tables(P);
  const Lookup& L = P.lookup;
  const int nt = P.nt;
  double e_vdW_tot = 0, e_ele_tot = 0;
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0;
#pragma omp parallel for schedule(dynamic, 32) reduction(+ : e_vdW_tot, e_ele_tot, v0, v1, v2, v3, v4, v5)
  for (int i = 0; i < s.n; i++) {
    int ti = s.type[i];
    if (ti < 0) continue;
    double fi[3] = {0, 0, 0};
    for (long pj = s.nb_off[i]; pj < s.nb_off[i + 1]; pj++) {
      int j = s.nb[pj];
      int tj = s.type[j];
      double dv[3] = {s.x[3 * j] - s.x[3 * i], s.x[3 * j + 1] - s.x[3 * i + 1], s.x[3 * j + 2] - s.x[3 * i + 2]};
      double r2 = dot3(dv, dv);
      if (!(r2 <= P.nonb_cut * P.nonb_cut)) continue;
      if (tj < 0) continue;
      double r_ij = sqrt(r2);
      const int tmin = std::min(ti, tj), tmax = std::max(ti, tj);
      int r = (int)(r_ij * L.inv_dx);
      if (r == 0) ++r;
      const double base = (double)(r + 1) * L.dx;
      const double dif = r_ij - base;
      auto ev = [&](int which) {
        const SplineCoef& c = L.at(nt, tmin, tmax, which)[r];
        return ((c.d * dif + c.c) * dif + c.b) * dif + c.a;
      };
      const double qq = s.q[i] * s.q[j];
      const double e_vdW = ev(1), e_ele = ev(3) * qq, CEvd = ev(2), CEclmb = ev(4) * qq;
      e_vdW_tot += 0.5 * e_vdW;
      e_ele_tot += 0.5 * e_ele;
      double fpair = -(CEvd + CEclmb);
      v0 += 0.5 * dv[0] * dv[0] * fpair; v1 += 0.5 * dv[1] * dv[1] * fpair; v2 += 0.5 * dv[2] * dv[2] * fpair;
      v3 += 0.5 * dv[0] * dv[1] * fpair; v4 += 0.5 * dv[0] * dv[2] * fpair; v5 += 0.5 * dv[1] * dv[2] * fpair;
      for (int t = 0; t < 3; t++) fi[t] += -(CEvd + CEclmb) * dv[t];
    }
    for (int t = 0; t < 3; t++) s.fCd[4 * i + t] += fi[t];
    v0 += s.x[3 * i] * fi[0]; v1 += s.x[3 * i + 1] * fi[1]; v2 += s.x[3 * i + 2] * fi[2];
    v3 += s.x[3 * i] * fi[1]; v4 += s.x[3 * i] * fi[2];     v5 += s.x[3 * i + 1] * fi[2];
  }
  s.en.e_vdW += e_vdW_tot;
  s.en.e_ele += e_ele_tot;
  s.virial[0] += v0; s.virial[1] += v1; s.virial[2] += v2; s.virial[3] += v3; s.virial[4] += v4; s.virial[5] += v5;
}

// a9: vdW_Coulomb_Energy_Full_C  serial twin reaxc_nonbonded_sw64.c:40-258
//     (full list, local i only, force on i only; pair virial via ev_tally_full reaxc_inlines_sw64.h:178-222)
void nonbonded(System& s) {
  if (s.prm.tabulate > 0) { nonbonded_tabulated(s); return; }   // Compute_NonBonded_Forces: tabulate == 0 ? analytic : tables
  const Params& P = s.prm;
  const double p_vdW1 = P.gp[28], p_vdW1i = 1.0 / p_vdW1;
  const double* Tap = P.Tap;
  double e_vdW_tot = 0, e_ele_tot = 0;
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0;
#pragma omp parallel for schedule(dynamic, 32) reduction(+ : e_vdW_tot, e_ele_tot, v0, v1, v2, v3, v4, v5)
  for (int i = 0; i < s.n; i++) {
    int ti = s.type[i];
    if (ti < 0) continue;
    double fi[3] = {0, 0, 0};
    for (long pj = s.nb_off[i]; pj < s.nb_off[i + 1]; pj++) {
      int j = s.nb[pj];
      int tj = s.type[j];
      double dv[3] = {s.x[3 * j] - s.x[3 * i], s.x[3 * j + 1] - s.x[3 * i + 1], s.x[3 * j + 2] - s.x[3 * i + 2]};
      double r2 = dot3(dv, dv);
      if (!(r2 <= P.nonb_cut * P.nonb_cut)) continue;
      if (tj < 0) continue;
      double r_ij = sqrt(r2);
      const Tbp& tw = P.tb(ti, tj);
      double T = Tap[7] * r_ij + Tap[6];
      T = T * r_ij + Tap[5]; T = T * r_ij + Tap[4]; T = T * r_ij + Tap[3];
      T = T * r_ij + Tap[2]; T = T * r_ij + Tap[1]; T = T * r_ij + Tap[0];
      double dT = 7 * Tap[7] * r_ij + 6 * Tap[6];
      dT = dT * r_ij + 5 * Tap[5]; dT = dT * r_ij + 4 * Tap[4]; dT = dT * r_ij + 3 * Tap[3];
      dT = dT * r_ij + 2 * Tap[2];
      dT += Tap[1] / r_ij;
      double e_vdW, CEvd, e_core = 0, e_lg = 0;
      if (P.vdw_type == 1 || P.vdw_type == 3) {
        double powr = pow(r_ij, p_vdW1);
        double powgi = pow(1.0 / tw.gamma_w, p_vdW1);
        double fn13 = pow(powr + powgi, p_vdW1i);
        double exp1 = exp(tw.alpha * (1.0 - fn13 / tw.r_vdW));
        double exp2 = exp(0.5 * tw.alpha * (1.0 - fn13 / tw.r_vdW));
        e_vdW = tw.D * (exp1 - 2.0 * exp2);
        double dfn13 = pow(powr + powgi, p_vdW1i - 1.0) * pow(r_ij, p_vdW1 - 2.0);
        CEvd = dT * e_vdW - T * tw.D * (tw.alpha / tw.r_vdW) * (exp1 - exp2) * dfn13;
      } else {
        double exp1 = exp(tw.alpha * (1.0 - r_ij / tw.r_vdW));
        double exp2 = exp(0.5 * tw.alpha * (1.0 - r_ij / tw.r_vdW));
        e_vdW = tw.D * (exp1 - 2.0 * exp2);
        CEvd = dT * e_vdW - T * tw.D * (tw.alpha / tw.r_vdW) * (exp1 - exp2) / r_ij;
      }
      if (P.vdw_type == 2 || P.vdw_type == 3) {
        e_core = tw.ecore * exp(tw.acore * (1.0 - (r_ij / tw.rcore)));
        double de_core = -(tw.acore / tw.rcore) * e_core;
        CEvd += dT * e_core + T * de_core / r_ij;
        if (P.lgflag) {
          double r5 = pow(r_ij, 5.0), r6 = pow(r_ij, 6.0), re6 = pow(tw.lgre, 6.0);
          e_lg = -(tw.lgcij / (r6 + re6));
          double de_lg = -6.0 * e_lg * r5 / (r6 + re6);
          CEvd += dT * e_lg + T * de_lg / r_ij;
        }
      }
      double dr3gamij_1 = r_ij * r_ij * r_ij + tw.gamma;
      double dr3gamij_3 = pow(dr3gamij_1, 0.33333333333333);
      double tmp = T / dr3gamij_3;
      double e_ele = C_ele * s.q[i] * s.q[j] * tmp;
      double CEclmb = C_ele * s.q[i] * s.q[j] * (dT - T * r_ij / dr3gamij_1) / dr3gamij_3;
      double pe_vdw = T * (e_vdW + e_core + e_lg);
      e_vdW_tot += 0.5 * pe_vdw;
      e_ele_tot += 0.5 * e_ele;
      double fpair = -(CEvd + CEclmb);
      // del = x_i - x_j = -dv
      v0 += 0.5 * dv[0] * dv[0] * fpair; v1 += 0.5 * dv[1] * dv[1] * fpair; v2 += 0.5 * dv[2] * dv[2] * fpair;
      v3 += 0.5 * dv[0] * dv[1] * fpair; v4 += 0.5 * dv[0] * dv[2] * fpair; v5 += 0.5 * dv[1] * dv[2] * fpair;
      for (int t = 0; t < 3; t++) fi[t] += -(CEvd + CEclmb) * dv[t];
    }
    // nonbonded is the first writer of fCd[i] (Compute_Forces order), owner-computes
    for (int t = 0; t < 3; t++) s.fCd[4 * i + t] += fi[t];
    // reaxc_nonbonded_sw64.c:247-252: -x_i (x) f_i^nb correction so that fdotr later cancels it
    v0 += s.x[3 * i] * fi[0]; v1 += s.x[3 * i + 1] * fi[1]; v2 += s.x[3 * i + 2] * fi[2];
    v3 += s.x[3 * i] * fi[1]; v4 += s.x[3 * i] * fi[2];     v5 += s.x[3 * i + 1] * fi[2];
  }
  s.en.e_vdW += e_vdW_tot;
  s.en.e_ele += e_ele_tot;
  s.virial[0] += v0; s.virial[1] += v1; s.virial[2] += v2; s.virial[3] += v3; s.virial[4] += v4; s.virial[5] += v5;
}

// a6: BO sections 1-3  reaxc_bond_orders_sw64.c:25-66,67-378,380-484 (== reaxc_bond_orders_sunway.cpp:460-774)
void bond_orders(System& s) {
  const Params& P = s.prm;
  const int N = s.N;
  const double p_boc1 = P.gp[0], p_boc2 = P.gp[1], p_lp1 = P.gp[15];
  s.Deltap.assign(N, 0); s.Deltap_boc.assign(N, 0);
  // section 1
  for (int i = 0; i < N; i++) {
    int ti = s.type[i];
    if (ti < 0) continue;
    s.Deltap[i] = s.total_bo[i] - P.sbp[ti].valency;
    s.Deltap_boc[i] = s.total_bo[i] - P.sbp[ti].valency_val;
    s.total_bo[i] = 0;
  }
  // section 2
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < N; i++) {
    int ti = s.type[i];
    if (ti < 0) continue;
    double val_i = P.sbp[ti].valency;
    double Deltap_i = s.Deltap[i], Deltap_boc_i = s.Deltap_boc[i];
    for (int pj = s.b_start[i]; pj < s.b_end[i]; pj++) {
      Bond& b = s.bonds[pj];
      int j = b.nbr;
      int tj = s.type[j];
      if (tj < 0) continue;
      for (int pk = s.b_start[j]; pk < s.b_end[j]; pk++)
        if (s.bonds[pk].nbr == i) b.sym = pk;
      const Tbp& tw = P.tb(ti, tj);
      if (tw.ovc < 0.001 && tw.v13cor < 0.001) {
        b.C1dbo = 1.0; b.C2dbo = 0.0; b.C3dbo = 0.0;
        b.C1dbopi = b.BO_pi; b.C2dbopi = 0.0; b.C3dbopi = 0.0; b.C4dbopi = 0.0;
        b.C1dbopi2 = b.BO_pi2; b.C2dbopi2 = 0.0; b.C3dbopi2 = 0.0; b.C4dbopi2 = 0.0;
      } else {
        double val_j = P.sbp[tj].valency;
        double Deltap_j = s.Deltap[j], Deltap_boc_j = s.Deltap_boc[j];
        double f1, f4, f5, f4f5, Cf1_ij, Cf1_ji, Cf45_ij, Cf45_ji;
        if (tw.ovc >= 0.001) {
          double exp_p1i = exp(-p_boc1 * Deltap_i), exp_p2i = exp(-p_boc2 * Deltap_i);
          double exp_p1j = exp(-p_boc1 * Deltap_j), exp_p2j = exp(-p_boc2 * Deltap_j);
          double f2 = exp_p1i + exp_p1j;
          double f3 = -1.0 / p_boc2 * log(0.5 * (exp_p2i + exp_p2j));
          f1 = 0.5 * ((val_i + f2) / (val_i + f2 + f3) + (val_j + f2) / (val_j + f2 + f3));
          double temp = f2 + f3;
          double u1_ij = val_i + temp, u1_ji = val_j + temp;
          double Cf1A_ij = 0.5 * f3 * (1.0 / SQR(u1_ij) + 1.0 / SQR(u1_ji));
          double Cf1B_ij = -0.5 * ((u1_ij - f3) / SQR(u1_ij) + (u1_ji - f3) / SQR(u1_ji));
          Cf1_ij = 0.50 * (-p_boc1 * exp_p1i / u1_ij -
                           ((val_i + f2) / SQR(u1_ij)) * (-p_boc1 * exp_p1i + exp_p2i / (exp_p2i + exp_p2j)) +
                           -p_boc1 * exp_p1i / u1_ji -
                           ((val_j + f2) / SQR(u1_ji)) * (-p_boc1 * exp_p1i + exp_p2i / (exp_p2i + exp_p2j)));
          Cf1_ji = -Cf1A_ij * p_boc1 * exp_p1j + Cf1B_ij * exp_p2j / (exp_p2i + exp_p2j);
        } else {
          f1 = 1.0;
          Cf1_ij = Cf1_ji = 0.0;
        }
        if (tw.v13cor >= 0.001) {
          double exp_f4 = exp(-(tw.p_boc4 * SQR(b.BO) - Deltap_boc_i) * tw.p_boc3 + tw.p_boc5);
          double exp_f5 = exp(-(tw.p_boc4 * SQR(b.BO) - Deltap_boc_j) * tw.p_boc3 + tw.p_boc5);
          f4 = 1. / (1. + exp_f4);
          f5 = 1. / (1. + exp_f5);
          f4f5 = f4 * f5;
          Cf45_ij = -f4 * exp_f4;
          Cf45_ji = -f5 * exp_f5;
        } else {
          f4 = f5 = f4f5 = 1.0;
          Cf45_ij = Cf45_ji = 0.0;
        }
        double A0_ij = f1 * f4f5;
        double A1_ij = -2 * tw.p_boc3 * tw.p_boc4 * b.BO * (Cf45_ij + Cf45_ji);
        double A2_ij = Cf1_ij / f1 + tw.p_boc3 * Cf45_ij;
        double A2_ji = Cf1_ji / f1 + tw.p_boc3 * Cf45_ji;
        double A3_ij = A2_ij + Cf1_ij / f1;
        double A3_ji = A2_ji + Cf1_ji / f1;
        b.BO = b.BO * A0_ij;
        b.BO_pi = b.BO_pi * A0_ij * f1;
        b.BO_pi2 = b.BO_pi2 * A0_ij * f1;
        b.BO_s = b.BO - (b.BO_pi + b.BO_pi2);
        b.C1dbo = A0_ij + b.BO * A1_ij;
        b.C2dbo = b.BO * A2_ij;
        b.C3dbo = b.BO * A2_ji;
        b.C1dbopi = f1 * f1 * f4 * f5;
        b.C2dbopi = b.BO_pi * A1_ij;
        b.C3dbopi = b.BO_pi * A3_ij;
        b.C4dbopi = b.BO_pi * A3_ji;
        b.C1dbopi2 = f1 * f1 * f4 * f5;
        b.C2dbopi2 = b.BO_pi2 * A1_ij;
        b.C3dbopi2 = b.BO_pi2 * A3_ij;
        b.C4dbopi2 = b.BO_pi2 * A3_ji;
      }
      if (b.BO < 1e-10) b.BO = 0.0;
      if (b.BO_s < 1e-10) b.BO_s = 0.0;
      if (b.BO_pi < 1e-10) b.BO_pi = 0.0;
      if (b.BO_pi2 < 1e-10) b.BO_pi2 = 0.0;
      s.total_bo[i] += b.BO;
    }
  }
  // section 3
  s.Delta.assign(N, 0); s.Delta_e.assign(N, 0); s.Delta_boc.assign(N, 0); s.Delta_val.assign(N, 0);
  s.vlpex.assign(N, 0); s.nlp.assign(N, 0); s.Delta_lp.assign(N, 0); s.Clp.assign(N, 0); s.dDelta_lp.assign(N, 0);
  s.nlp_temp.assign(N, 0); s.Delta_lp_temp.assign(N, 0); s.dDelta_lp_temp.assign(N, 0);
  for (int j = 0; j < N; j++) {
    int tj = s.type[j];
    if (tj < 0) continue;  // the reference indexes sbp[-1] here; NULL types are excluded by setflag upstream
    const Sbp& sb = P.sbp[tj];
    s.Delta[j] = s.total_bo[j] - sb.valency;
    s.Delta_e[j] = s.total_bo[j] - sb.valency_e;
    s.Delta_boc[j] = s.total_bo[j] - sb.valency_boc;
    s.Delta_val[j] = s.total_bo[j] - sb.valency_val;
    s.vlpex[j] = s.Delta_e[j] - 2.0 * (int)(s.Delta_e[j] / 2.0);
    double explp1 = exp(-p_lp1 * SQR(2.0 + s.vlpex[j]));
    s.nlp[j] = explp1 - (int)(s.Delta_e[j] / 2.0);
    s.Delta_lp[j] = sb.nlp_opt - s.nlp[j];
    s.Clp[j] = 2.0 * p_lp1 * explp1 * (2.0 + s.vlpex[j]);
    s.dDelta_lp[j] = s.Clp[j];
    if (sb.mass > 21.0) {
      s.nlp_temp[j] = 0.5 * (sb.valency_e - sb.valency);
      s.Delta_lp_temp[j] = sb.nlp_opt - s.nlp_temp[j];
      s.dDelta_lp_temp[j] = 0.;
    } else {
      s.nlp_temp[j] = s.nlp[j];
      s.Delta_lp_temp[j] = sb.nlp_opt - s.nlp_temp[j];
      s.dDelta_lp_temp[j] = s.Clp[j];
    }
  }
}

// tag-ordered half selection with z,y,x tie-break:  reaxc_multi_body_sw64.c:256-268,
// reaxc_torsion_angles_sunway.cpp:982-992
static inline bool half_select(const System& s, int i, int j) {
  if (s.tag[i] > s.tag[j]) return false;
  if (s.tag[i] == s.tag[j]) {
    const double *xi = &s.x[3 * i], *xj = &s.x[3 * j];
    if (xj[2] < xi[2]) return false;
    if (xj[2] == xi[2] && xj[1] < xi[1]) return false;
    if (xj[2] == xi[2] && xj[1] == xi[1] && xj[0] < xi[0]) return false;
  }
  return true;
}

// a7: Merge_Bonds_Atom_Energy_C_New  reaxc_multi_body_sw64.c:21-333
void bonds_atom_energy(System& s) {
  const Params& P = s.prm;
  const double p_lp3 = P.gp[5], p_ovun3 = P.gp[32], p_ovun4 = P.gp[31], p_ovun6 = P.gp[6], p_ovun7 = P.gp[8],
               p_ovun8 = P.gp[9];
  const double gp3 = P.gp[3], gp4 = P.gp[4], gp7 = P.gp[7], gp10 = P.gp[10];
  const int gp37 = (int)P.gp[37];
  double e_lph = 0, e_lp = 0, e_ov = 0, e_un_tot = 0, ebond = 0, estriph = 0, e_pol = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : e_lph, e_lp, e_ov, e_un_tot, ebond, estriph, e_pol)
  for (int i = 0; i < s.n; i++) {
    int ti = s.type[i];
    if (ti < 0) continue;
    const Sbp& sbi = P.sbp[ti];
    double q = s.q[i];
    e_pol += KCALpMOL_to_EV * (sbi.chi * q + (sbi.eta / 2.) * SQR(q));

    double p_lp2 = sbi.p_lp2;
    double expvd2 = exp(-75 * s.Delta_lp[i]);
    double inv_expvd2 = 1. / (1. + expvd2);
    int numbonds = 0;
    double p_ovun2 = sbi.p_ovun2;
    double sum_ovun1 = 0, sum_ovun2 = 0;
    double dfvl = (sbi.mass > 21.0) ? 0.0 : 1.0;
    for (int pj = s.b_start[i]; pj < s.b_end[i]; pj++) {
      numbonds++;
      const Bond& b = s.bonds[pj];
      int j = b.nbr, tj = s.type[j];
      if (tj < 0) continue;
      const Tbp& tw = P.tb(ti, tj);
      sum_ovun1 += tw.p_ovun1 * tw.De_s * b.BO;
      sum_ovun2 += (s.Delta[j] - dfvl * s.Delta_lp_temp[j]) * (b.BO_pi + b.BO_pi2);
    }
    double dElp = p_lp2 * inv_expvd2 + 75 * p_lp2 * s.Delta_lp[i] * expvd2 * SQR(inv_expvd2);
    double CElp = dElp * s.dDelta_lp[i];

    if (p_lp3 > 0.001 && !strcmp(sbi.name, "C")) {
      for (int pj = s.b_start[i]; pj < s.b_end[i]; pj++) {
        Bond& b = s.bonds[pj];
        int j = b.nbr, tj = s.type[j];
        if (tj < 0) continue;
        if (!strcmp(P.sbp[tj].name, "C")) {
          double Di = s.Delta[i];
          double vov3 = b.BO - Di - 0.040 * pow(Di, 4.);
          if (vov3 > 3.) {
            e_lph += p_lp3 * SQR(vov3 - 3.0);
            double deahu2dbo = 2. * p_lp3 * (vov3 - 3.);
            double deahu2dsbo = 2. * p_lp3 * (vov3 - 3.) * (-1. - 0.16 * pow(Di, 3.));
            atomic_add(b.Cdbo, deahu2dbo);
            atomic_add(s.fCd[4 * i + 3], deahu2dsbo);
          }
        }
      }
    }

    double exp_ovun1 = p_ovun3 * exp(p_ovun4 * sum_ovun2);
    double inv_exp_ovun1 = 1.0 / (1 + exp_ovun1);
    double Delta_lpcorr = s.Delta[i] - (dfvl * s.Delta_lp_temp[i]) * inv_exp_ovun1;
    double exp_ovun2 = exp(p_ovun2 * Delta_lpcorr);
    double inv_exp_ovun2 = 1.0 / (1.0 + exp_ovun2);
    double DlpVi = 1.0 / (Delta_lpcorr + sbi.valency + 1e-8);
    double CEover1 = Delta_lpcorr * DlpVi * inv_exp_ovun2;
    e_ov += sum_ovun1 * CEover1;
    double CEover2 = sum_ovun1 * DlpVi * inv_exp_ovun2 *
                     (1.0 - Delta_lpcorr * (DlpVi + p_ovun2 * exp_ovun2 * inv_exp_ovun2));
    double CEover3 = CEover2 * (1.0 - dfvl * s.dDelta_lp[i] * inv_exp_ovun1);
    double CEover4 = CEover2 * (dfvl * s.Delta_lp_temp[i]) * p_ovun4 * exp_ovun1 * SQR(inv_exp_ovun1);

    double p_ovun5 = sbi.p_ovun5;
    double exp_ovun2n = 1.0 / exp_ovun2;
    double exp_ovun6 = exp(p_ovun6 * Delta_lpcorr);
    double exp_ovun8 = p_ovun7 * exp(p_ovun8 * sum_ovun2);
    double inv_exp_ovun2n = 1.0 / (1.0 + exp_ovun2n);
    double inv_exp_ovun8 = 1.0 / (1.0 + exp_ovun8);
    double e_un = 0.0;
    atomic_add(s.fCd[4 * i + 3], CEover3);
    double CEunder1, CEunder2, CEunder3, CEunder4;
    if (numbonds > 0 || P.enobondsflag) {
      e_un = -p_ovun5 * (1.0 - exp_ovun6) * inv_exp_ovun2n * inv_exp_ovun8;
      e_un_tot += e_un;
      CEunder1 = inv_exp_ovun2n * (p_ovun5 * p_ovun6 * exp_ovun6 * inv_exp_ovun8 + p_ovun2 * e_un * exp_ovun2n);
      CEunder2 = -e_un * p_ovun8 * exp_ovun8 * inv_exp_ovun8;
      CEunder3 = CEunder1 * (1.0 - dfvl * s.dDelta_lp[i] * inv_exp_ovun1);
      CEunder4 = CEunder1 * (dfvl * s.Delta_lp_temp[i]) * p_ovun4 * exp_ovun1 * SQR(inv_exp_ovun1) + CEunder2;
      atomic_add(s.fCd[4 * i + 3], CEunder3);
      atomic_add(s.fCd[4 * i + 3], CElp);
      e_lp += p_lp2 * s.Delta_lp[i] * inv_expvd2;
    } else {
      CEunder1 = inv_exp_ovun2n * (p_ovun5 * p_ovun6 * exp_ovun6 * inv_exp_ovun8 + p_ovun2 * e_un * exp_ovun2n);
      CEunder2 = -e_un * p_ovun8 * exp_ovun8 * inv_exp_ovun8;
      CEunder3 = CEunder1 * (1.0 - dfvl * s.dDelta_lp[i] * inv_exp_ovun1);
      CEunder4 = CEunder1 * (dfvl * s.Delta_lp_temp[i]) * p_ovun4 * exp_ovun1 * SQR(inv_exp_ovun1) + CEunder2;
    }
    (void)CEunder3;

    for (int pj = s.b_start[i]; pj < s.b_end[i]; pj++) {
      Bond& b = s.bonds[pj];
      int j = b.nbr, tj = s.type[j];
      if (tj < 0) continue;  // reference indexes tbp[ti][-1]; unreachable when all types are mapped
      const Tbp& tw = P.tb(ti, tj);
      const Sbp& sbj = P.sbp[tj];
      double cdbo = 0, cdbopi = 0, cdbopi2 = 0;
      cdbo += CEover1 * tw.p_ovun1 * tw.De_s;
      double ftmp2 = (1.0 - dfvl * s.dDelta_lp[j]) * (b.BO_pi + b.BO_pi2);
      atomic_add(s.fCd[4 * j + 3], CEover4 * ftmp2);
      double ftmp = CEover4 * (s.Delta[j] - dfvl * s.Delta_lp_temp[j]);
      cdbopi += ftmp; cdbopi2 += ftmp;
      atomic_add(s.fCd[4 * j + 3], CEunder4 * ftmp2);
      double fc3 = CEunder4 * (s.Delta[j] - dfvl * s.Delta_lp_temp[j]);
      cdbopi += fc3; cdbopi2 += fc3;
      if (half_select(s, i, j)) {
        double pow_BOs_be2 = (b.BO_s == 0.0) ? 0.0 : pow(b.BO_s, tw.p_be2);
        double exp_be12 = exp(tw.p_be1 * (1.0 - pow_BOs_be2));
        double CEbo = -tw.De_s * exp_be12 * (1.0 - tw.p_be1 * tw.p_be2 * pow_BOs_be2);
        ebond += -tw.De_s * b.BO_s * exp_be12 - tw.De_p * b.BO_pi - tw.De_pp * b.BO_pi2;
        cdbo += CEbo;
        cdbopi -= (CEbo + tw.De_p);
        cdbopi2 -= (CEbo + tw.De_pp);
        if (b.BO >= 1.00) {
          if (gp37 == 2 || (sbi.mass == 12.0000 && sbj.mass == 15.9990) || (sbj.mass == 12.0000 && sbi.mass == 15.9990)) {
            double exphu = exp(-gp7 * SQR(b.BO - 2.50));
            double exphua1 = exp(-gp3 * (s.total_bo[i] - b.BO));
            double exphub1 = exp(-gp3 * (s.total_bo[j] - b.BO));
            double exphuov = exp(gp4 * (s.Delta[i] + s.Delta[j]));
            double hulpov = 1.0 / (1.0 + 25.0 * exphuov);
            estriph += gp10 * exphu * hulpov * (exphua1 + exphub1);
            double decobdbo = gp10 * exphu * hulpov * (exphua1 + exphub1) * (gp3 - 2.0 * gp7 * (b.BO - 2.50));
            double decobdboua = -gp10 * exphu * hulpov * (gp3 * exphua1 + 25.0 * gp4 * exphuov * hulpov * (exphua1 + exphub1));
            double decobdboub = -gp10 * exphu * hulpov * (gp3 * exphub1 + 25.0 * gp4 * exphuov * hulpov * (exphua1 + exphub1));
            cdbo += decobdbo;
            atomic_add(s.fCd[4 * i + 3], decobdboua);
            atomic_add(s.fCd[4 * j + 3], decobdboub);
          }
        }
      }
      atomic_add(b.Cdbo, cdbo);
      atomic_add(b.Cdbopi, cdbopi);
      atomic_add(b.Cdbopi2, cdbopi2);
    }
  }
  s.en.e_pol += e_pol;
  s.en.e_lp += (e_lph + e_lp);
  s.en.e_ov += e_ov;
  s.en.e_un += e_un_tot;
  s.en.e_bond += (ebond + estriph);
}

// reaxc_valence_angles_sunway.cpp:50-84
static inline void calc_theta(const double* dji, double d_ji, const double* djk, double d_jk, double* theta, double* cos_theta) {
  *cos_theta = dot3(dji, djk) / (d_ji * d_jk);
  if (*cos_theta > 1.) *cos_theta = 1.0;
  if (*cos_theta < -1.) *cos_theta = -1.0;
  *theta = acos(*cos_theta);
}
static inline void calc_dcos_theta(const double* dji, double d_ji, const double* djk, double d_jk, double* di, double* dj, double* dk) {
  double sqr_d_ji = SQR(d_ji), sqr_d_jk = SQR(d_jk);
  double inv_dists = 1.0 / (d_ji * d_jk);
  double inv_dists3 = pow(inv_dists, 3.0);
  double dot_dvecs = dot3(dji, djk);
  double Cdot_inv3 = dot_dvecs * inv_dists3;
  for (int t = 0; t < 3; t++) {
    di[t] = djk[t] * inv_dists - Cdot_inv3 * sqr_d_jk * dji[t];
    dj[t] = -(djk[t] + dji[t]) * inv_dists + Cdot_inv3 * (sqr_d_jk * dji[t] + sqr_d_ji * djk[t]);
    dk[t] = dji[t] * inv_dists - Cdot_inv3 * sqr_d_ji * djk[t];
  }
}

// a8: Hydrogen_Bonds  live serial reaxc_hydrogen_bonds_sunway.cpp:310-436
void hydrogen_bonds(System& s) {
  const Params& P = s.prm;
  double e_hb_tot = 0;
#pragma omp parallel for schedule(dynamic, 32) reduction(+ : e_hb_tot)
  for (int j = 0; j < s.n; j++) {
    int tj = s.type[j];
    if (tj < 0) continue;
    if (P.sbp[tj].p_hbond != 1) continue;
    if (s.Hindex[j] < 0) continue;
    std::vector<int> hblist;
    for (int pi = s.b_start[j]; pi < s.b_end[j]; pi++) {
      const Bond& b = s.bonds[pi];
      int ti = s.type[b.nbr];
      if (ti < 0) continue;
      if (P.sbp[ti].p_hbond == 2 && b.BO >= HB_THRESHOLD) hblist.push_back(pi);
    }
    int h = s.Hindex[j];
    for (int pk = s.hb_start[h]; pk < s.hb_end[h]; pk++) {
      const HBond& hb = s.hbonds[pk];
      int k = hb.nbr, tk = s.type[k];
      if (tk < 0) continue;
      double r_jk = hb.d;
      const double* dvec_jk = hb.dvec;
      for (int pi : hblist) {
        Bond& b = s.bonds[pi];
        int i = b.nbr;
        if (s.tag[i] == s.tag[k]) continue;
        int ti = s.type[i];
        if (ti < 0) continue;
        const Hbp& hp = P.hb(ti, tj, tk);
        if (hp.r0_hb <= 0.0) continue;
        double theta, cos_theta, di[3], dj[3], dk[3];
        calc_theta(b.dvec, b.d, dvec_jk, r_jk, &theta, &cos_theta);
        calc_dcos_theta(b.dvec, b.d, dvec_jk, r_jk, di, dj, dk);
        double sin_theta2 = sin(theta / 2.0);
        double sin_xhz4 = SQR(sin_theta2);
        sin_xhz4 *= sin_xhz4;
        double cos_xhz1 = (1.0 - cos_theta);
        double exp_hb2 = exp(-hp.p_hb2 * b.BO);
        double exp_hb3 = exp(-hp.p_hb3 * (hp.r0_hb / r_jk + r_jk / hp.r0_hb - 2.0));
        double e_hb = hp.p_hb1 * (1.0 - exp_hb2) * exp_hb3 * sin_xhz4;
        e_hb_tot += e_hb;
        double CEhb1 = hp.p_hb1 * hp.p_hb2 * exp_hb2 * exp_hb3 * sin_xhz4;
        double CEhb2 = -hp.p_hb1 / 2.0 * (1.0 - exp_hb2) * exp_hb3 * cos_xhz1;
        double CEhb3 = -hp.p_hb3 * (-hp.r0_hb / SQR(r_jk) + 1.0 / hp.r0_hb) * e_hb;
        atomic_add(b.Cdbo, CEhb1);
        add4(&s.fCd[4 * i], +CEhb2, di);
        add4(&s.fCd[4 * j], +CEhb2, dj);
        add4(&s.fCd[4 * k], +CEhb2, dk);
        add4(&s.fCd[4 * j], -CEhb3 / r_jk, dvec_jk);
        add4(&s.fCd[4 * k], +CEhb3 / r_jk, dvec_jk);
      }
    }
  }
  s.en.e_hb += e_hb_tot;
}

struct Thb { double theta, dcos_di[3], dcos_dj[3], dcos_dk[3]; };

// reaxc_torsion_angles_sunway.cpp:44-125
static double calc_omega(const double* dvec_ij, double r_ij, const double* dvec_jk, double r_jk, const double* dvec_kl,
                         double r_kl, const double* dvec_li, double r_li, const Thb* p_ijk, const Thb* p_jkl,
                         double* dcos_omega_di, double* dcos_omega_dj, double* dcos_omega_dk, double* dcos_omega_dl) {
  double sin_ijk = sin(p_ijk->theta), cos_ijk = cos(p_ijk->theta);
  double sin_jkl = sin(p_jkl->theta), cos_jkl = cos(p_jkl->theta);
  double unnorm_cos_omega = -dot3(dvec_ij, dvec_jk) * dot3(dvec_jk, dvec_kl) + SQR(r_jk) * dot3(dvec_ij, dvec_kl);
  double cross_jk_kl[3] = {dvec_jk[1] * dvec_kl[2] - dvec_jk[2] * dvec_kl[1], dvec_jk[2] * dvec_kl[0] - dvec_jk[0] * dvec_kl[2],
                           dvec_jk[0] * dvec_kl[1] - dvec_jk[1] * dvec_kl[0]};
  double unnorm_sin_omega = -r_jk * dot3(dvec_ij, cross_jk_kl);
  double omega = atan2(unnorm_sin_omega, unnorm_cos_omega);
  double htra = r_ij + cos_ijk * (r_kl * cos_jkl - r_jk);
  double htrb = r_jk - r_ij * cos_ijk - r_kl * cos_jkl;
  double htrc = r_kl + cos_jkl * (r_ij * cos_ijk - r_jk);
  double hthd = r_ij * sin_ijk * (r_jk - r_kl * cos_jkl);
  double hthe = r_kl * sin_jkl * (r_jk - r_ij * cos_ijk);
  double hnra = r_kl * sin_ijk * sin_jkl;
  double hnrc = r_ij * sin_ijk * sin_jkl;
  double hnhd = r_ij * r_kl * cos_ijk * sin_jkl;
  double hnhe = r_ij * r_kl * sin_ijk * cos_jkl;
  double poem = 2.0 * r_ij * r_kl * sin_ijk * sin_jkl;
  if (poem < 1e-20) poem = 1e-20;
  double tel = SQR(r_ij) + SQR(r_jk) + SQR(r_kl) - SQR(r_li) -
               2.0 * (r_ij * r_jk * cos_ijk - r_ij * r_kl * cos_ijk * cos_jkl + r_jk * r_kl * cos_jkl);
  double arg = tel / poem;
  if (arg > 1.0) arg = 1.0;
  if (arg < -1.0) arg = -1.0;
  if (sin_ijk >= 0 && sin_ijk <= MIN_SINE) sin_ijk = MIN_SINE;
  else if (sin_ijk <= 0 && sin_ijk >= -MIN_SINE) sin_ijk = -MIN_SINE;
  if (sin_jkl >= 0 && sin_jkl <= MIN_SINE) sin_jkl = MIN_SINE;
  else if (sin_jkl <= 0 && sin_jkl >= -MIN_SINE) sin_jkl = -MIN_SINE;
  for (int t = 0; t < 3; t++) {
    double di = (htra - arg * hnra) / r_ij * dvec_ij[t] + -1. * dvec_li[t];
    di += -(hthd - arg * hnhd) / sin_ijk * p_ijk->dcos_dk[t];
    dcos_omega_di[t] = 2.0 / poem * di;
    double dj = -(htra - arg * hnra) / r_ij * dvec_ij[t] + -htrb / r_jk * dvec_jk[t];
    dj += -(hthd - arg * hnhd) / sin_ijk * p_ijk->dcos_dj[t];
    dj += -(hthe - arg * hnhe) / sin_jkl * p_jkl->dcos_di[t];
    dcos_omega_dj[t] = 2.0 / poem * dj;
    double dk = -(htrc - arg * hnrc) / r_kl * dvec_kl[t] + htrb / r_jk * dvec_jk[t];
    dk += -(hthd - arg * hnhd) / sin_ijk * p_ijk->dcos_di[t];
    dk += -(hthe - arg * hnhe) / sin_jkl * p_jkl->dcos_dj[t];
    dcos_omega_dk[t] = 2.0 / poem * dk;
    double dl = (htrc - arg * hnrc) / r_kl * dvec_kl[t] + 1. * dvec_li[t];
    dl += -(hthe - arg * hnhe) / sin_jkl * p_jkl->dcos_dk[t];
    dcos_omega_dl[t] = 2.0 / poem * dl;
  }
  return omega;
}

// a10: merged valence + torsion, no stored three-body list
//      live serial reaxc_torsion_angles_sunway.cpp:652-1312 (== reaxc_torsion_angles_cpe.h, j in [0,n))
void valence_torsion(System& s) {
  const Params& P = s.prm;
  const double p_tor2 = P.gp[23], p_tor3 = P.gp[24], p_tor4 = P.gp[25], p_cot2 = P.gp[27];
  const double p_val6 = P.gp[14], p_val8 = P.gp[33], p_val9 = P.gp[16], p_val10 = P.gp[17];
  const double p_pen2 = P.gp[19], p_pen3 = P.gp[20], p_pen4 = P.gp[21];
  const double p_coa2 = P.gp[2], p_coa3 = P.gp[38], p_coa4 = P.gp[30];
  const double thb_cut = P.thb_cut, thb_cutsq = P.thb_cutsq;
  double e_ang_t = 0, e_pen_t = 0, e_coa_t = 0, e_tor_t = 0, e_con_t = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : e_ang_t, e_pen_t, e_coa_t, e_tor_t, e_con_t)
  for (int j = 0; j < s.N; j++) {
    int type_j = s.type[j];
    double Delta_j = s.Delta_boc[j];
    int start_j = s.b_start[j], end_j = s.b_end[j];
    if (type_j < 0 || end_j <= start_j) continue;
    double p_val3 = P.sbp[type_j].p_val3, p_val5 = P.sbp[type_j].p_val5;
    double SBOp = 0, prod_SBO = 1;
    for (int t = start_j; t < end_j; t++) {
      const Bond& bt = s.bonds[t];
      SBOp += (bt.BO_pi + bt.BO_pi2);
      double temp = SQR(bt.BO);
      temp *= temp;
      temp *= temp;
      prod_SBO *= exp(-temp);
    }
    double vlpadj, dSBO2;
    if (s.vlpex[j] >= 0) { vlpadj = 0; dSBO2 = prod_SBO - 1; }
    else { vlpadj = s.nlp[j]; dSBO2 = (prod_SBO - 1) * (1 - p_val8 * s.dDelta_lp[j]); }
    double SBO = SBOp + (1 - prod_SBO) * (-s.Delta_boc[j] - p_val8 * vlpadj);
    double dSBO1 = -8 * prod_SBO * (s.Delta_boc[j] + p_val8 * vlpadj);
    double SBO2, CSBO2;
    if (SBO <= 0) { SBO2 = 0; CSBO2 = 0; }
    else if (SBO > 0 && SBO <= 1) { SBO2 = pow(SBO, p_val9); CSBO2 = p_val9 * pow(SBO, p_val9 - 1); }
    else if (SBO > 1 && SBO < 2) { SBO2 = 2 - pow(2 - SBO, p_val9); CSBO2 = p_val9 * pow(2 - SBO, p_val9 - 1); }
    else { SBO2 = 2; CSBO2 = 0; }
    double expval6 = exp(p_val6 * s.Delta_boc[j]);

    for (int pk = start_j; pk < end_j; pk++) {
      Bond& bjk = s.bonds[pk];
      int k = bjk.nbr;
      double BOA_jk = bjk.BO - thb_cut;
      if (!(BOA_jk > 0.0 && (j < s.n || k < s.n))) continue;
      int pj = bjk.sym;
      Bond& bpj = s.bonds[pj];
      int start_k = s.b_start[k], end_k = s.b_end[k];
      if (!(end_k > start_k)) continue;
      int type_k = s.type[k];
      double Delta_k = s.Delta_boc[k];
      double r_jk = bjk.d;
      double exp_tor2_jk = exp(-p_tor2 * BOA_jk);
      double exp_cot2_jk = exp(-p_cot2 * SQR(BOA_jk - 1.5));
      double exp_tor3_DjDk = exp(-p_tor3 * (Delta_j + Delta_k));
      double exp_tor4_DjDk = exp(p_tor4 * (Delta_j + Delta_k));
      double exp_tor34_inv = 1.0 / (1.0 + exp_tor3_DjDk + exp_tor4_DjDk);
      double f11_DjDk = (2.0 + exp_tor3_DjDk) * exp_tor34_inv;

      for (int ph = start_j; ph < end_j; ph++) {
        if (ph == pk) continue;
        Bond& bhj = s.bonds[ph];
        double BOA_hj = bhj.BO - thb_cut;
        int h = bhj.nbr;
        int type_h = s.type[h];
        Thb p_hjk;
        double theta_hjk, cos_theta_hjk;
        calc_theta(bjk.dvec, bjk.d, bhj.dvec, bhj.d, &theta_hjk, &cos_theta_hjk);
        calc_dcos_theta(bjk.dvec, bjk.d, bhj.dvec, bhj.d, p_hjk.dcos_di, p_hjk.dcos_dj, p_hjk.dcos_dk);
        p_hjk.theta = theta_hjk;
        double sin_theta_hjk = sin(theta_hjk);
        if (sin_theta_hjk < 1.0e-5) sin_theta_hjk = 1.0e-5;

        // ---- valence angle k-j-h ----
        if ((ph > pk) && (j < s.n) && (bjk.BO > thb_cut) && (bhj.BO > thb_cut) && (bjk.BO * bhj.BO > thb_cutsq)) {
          const ThbHeader& thbh = P.thb(type_k, type_j, type_h);
          for (int cnt = 0; cnt < thbh.cnt; cnt++) {
            if (fabs(thbh.prm[cnt].p_val1) > 0.001) {
              const Thbp& tp = thbh.prm[cnt];
              double p_val1 = tp.p_val1, p_val2 = tp.p_val2, p_val4 = tp.p_val4, p_val7 = tp.p_val7;
              double theta_00 = tp.theta_00;
              double exp3jk = exp(-p_val3 * pow(BOA_jk, p_val4));
              double f7_jk = 1.0 - exp3jk;
              double Cf7jk = p_val3 * p_val4 * pow(BOA_jk, p_val4 - 1.0) * exp3jk;
              double exp3hj = exp(-p_val3 * pow(BOA_hj, p_val4));
              double f7_hj = 1.0 - exp3hj;
              double Cf7hj = p_val3 * p_val4 * pow(BOA_hj, p_val4 - 1.0) * exp3hj;
              double expval7 = exp(-p_val7 * s.Delta_boc[j]);
              double trm8 = 1.0 + expval6 + expval7;
              double f8_Dj = p_val5 - ((p_val5 - 1.0) * (2.0 + expval6) / trm8);
              double Cf8j = ((1.0 - p_val5) / SQR(trm8)) *
                            (p_val6 * expval6 * trm8 - (2.0 + expval6) * (p_val6 * expval6 - p_val7 * expval7));
              double theta_0 = 180.0 - theta_00 * (1.0 - exp(-p_val10 * (2.0 - SBO2)));
              theta_0 = DEG2RAD(theta_0);
              double expval2theta = exp(-p_val2 * SQR(theta_0 - theta_hjk));
              double expval12theta;
              if (p_val1 >= 0) expval12theta = p_val1 * (1.0 - expval2theta);
              else expval12theta = p_val1 * -expval2theta;
              double CEval1 = Cf7jk * f7_hj * f8_Dj * expval12theta;
              double CEval2 = Cf7hj * f7_jk * f8_Dj * expval12theta;
              double CEval3 = Cf8j * f7_hj * f7_jk * expval12theta;
              double CEval4 = -2.0 * p_val1 * p_val2 * f7_jk * f7_hj * f8_Dj * expval2theta * (theta_0 - theta_hjk);
              double Ctheta_0 = p_val10 * DEG2RAD(theta_00) * exp(-p_val10 * (2.0 - SBO2));
              double CEval5 = -CEval4 * Ctheta_0 * CSBO2;
              double CEval6 = CEval5 * dSBO1;
              double CEval7 = CEval5 * dSBO2;
              double CEval8 = -CEval4 / sin_theta_hjk;
              double e_ang = f7_jk * f7_hj * f8_Dj * expval12theta;
              e_ang_t += e_ang;

              double p_pen1 = tp.p_pen1;
              double exp_pen2jk = exp(-p_pen2 * SQR(BOA_jk - 2.0));
              double exp_pen2hj = exp(-p_pen2 * SQR(BOA_hj - 2.0));
              double exp_pen3 = exp(-p_pen3 * s.Delta[j]);
              double exp_pen4 = exp(p_pen4 * s.Delta[j]);
              double trm_pen34 = 1.0 + exp_pen3 + exp_pen4;
              double f9_Dj = (2.0 + exp_pen3) / trm_pen34;
              double Cf9j = (-p_pen3 * exp_pen3 * trm_pen34 - (2.0 + exp_pen3) * (-p_pen3 * exp_pen3 + p_pen4 * exp_pen4)) /
                            SQR(trm_pen34);
              double e_pen = p_pen1 * f9_Dj * exp_pen2jk * exp_pen2hj;
              e_pen_t += e_pen;
              double CEpen1 = e_pen * Cf9j / f9_Dj;
              double temp = -2.0 * p_pen2 * e_pen;
              double CEpen2 = temp * (BOA_jk - 2.0);
              double CEpen3 = temp * (BOA_hj - 2.0);

              double p_coa1 = tp.p_coa1;
              double exp_coa2 = exp(p_coa2 * s.Delta_val[j]);
              double e_coa = p_coa1 / (1. + exp_coa2) * exp(-p_coa3 * SQR(s.total_bo[k] - BOA_jk)) *
                             exp(-p_coa3 * SQR(s.total_bo[h] - BOA_hj)) * exp(-p_coa4 * SQR(BOA_jk - 1.5)) *
                             exp(-p_coa4 * SQR(BOA_hj - 1.5));
              e_coa_t += e_coa;
              double CEcoa1 = -2 * p_coa4 * (BOA_jk - 1.5) * e_coa;
              double CEcoa2 = -2 * p_coa4 * (BOA_hj - 1.5) * e_coa;
              double CEcoa3 = -p_coa2 * exp_coa2 * e_coa / (1 + exp_coa2);
              double CEcoa4 = -2 * p_coa3 * (s.total_bo[k] - BOA_jk) * e_coa;
              double CEcoa5 = -2 * p_coa3 * (s.total_bo[h] - BOA_hj) * e_coa;

              atomic_add(bjk.Cdbo, (CEval1 + CEpen2 + (CEcoa1 - CEcoa4)));
              atomic_add(bhj.Cdbo, (CEval2 + CEpen3 + (CEcoa2 - CEcoa5)));
              atomic_add(s.fCd[4 * j + 3], ((CEval3 + CEval7) + CEpen1 + CEcoa3));
              atomic_add(s.fCd[4 * k + 3], CEcoa4);
              atomic_add(s.fCd[4 * h + 3], CEcoa5);
              for (int t = start_j; t < end_j; t++) {
                Bond& bt = s.bonds[t];
                double temp_bo_jt = bt.BO;
                double tmp3 = CUBE(temp_bo_jt);
                double pBOjt7 = tmp3 * tmp3 * temp_bo_jt;
                atomic_add(bt.Cdbo, (CEval6 * pBOjt7));
                atomic_add(bt.Cdbopi, CEval5);
                atomic_add(bt.Cdbopi2, CEval5);
              }
              add4(&s.fCd[4 * k], CEval8, p_hjk.dcos_di);
              add4(&s.fCd[4 * j], CEval8, p_hjk.dcos_dj);
              add4(&s.fCd[4 * h], CEval8, p_hjk.dcos_dk);
            }
          }
        }

        // ---- torsion h-j-k-w ----
        if (!half_select(s, j, k)) continue;
        if (!(bhj.BO > thb_cut && j < s.n && end_k > start_k)) continue;
        int i = h, type_i = type_h;
        const Thb* p_ijk = &p_hjk;
        Bond& bij = bhj;
        double r_ij = bij.d;
        double BOA_ij = bij.BO - thb_cut;
        double theta_ijk = p_ijk->theta;
        double sin_ijk = sin(theta_ijk), cos_ijk = cos(theta_ijk);
        double tan_ijk_i;
        if (sin_ijk >= 0 && sin_ijk <= MIN_SINE) tan_ijk_i = cos_ijk / MIN_SINE;
        else if (sin_ijk <= 0 && sin_ijk >= -MIN_SINE) tan_ijk_i = cos_ijk / -MIN_SINE;
        else tan_ijk_i = cos_ijk / sin_ijk;
        double exp_tor2_ij = exp(-p_tor2 * BOA_ij);
        double exp_cot2_ij = exp(-p_cot2 * SQR(BOA_ij - 1.5));

        for (int pw = start_k; pw < end_k; pw++) {
          if (pw == pj) continue;
          Bond& bkl = s.bonds[pw];
          int l = bkl.nbr;
          int type_l = s.type[l];
          if (type_l < 0 || type_i < 0 || type_k < 0) continue;
          Thb p_jkl;
          double theta_jkl, cos_theta_jkw;
          calc_theta(bpj.dvec, bpj.d, bkl.dvec, bkl.d, &theta_jkl, &cos_theta_jkw);
          calc_dcos_theta(bpj.dvec, bpj.d, bkl.dvec, bkl.d, p_jkl.dcos_di, p_jkl.dcos_dj, p_jkl.dcos_dk);
          p_jkl.theta = theta_jkl;
          const FbHeader& fbh = P.fb(type_i, type_j, type_k, type_l);
          const Fbp& fp = fbh.prm[0];
          if (!(i != l && fbh.cnt && bkl.BO > thb_cut && bij.BO * bjk.BO * bkl.BO > thb_cut)) continue;
          double r_kl = bkl.d;
          double BOA_kl = bkl.BO - thb_cut;
          double sin_jkl = sin(theta_jkl), cos_jkl = cos(theta_jkl);
          double tan_jkl_i;
          if (sin_jkl >= 0 && sin_jkl <= MIN_SINE) tan_jkl_i = cos_jkl / MIN_SINE;
          else if (sin_jkl <= 0 && sin_jkl >= -MIN_SINE) tan_jkl_i = cos_jkl / -MIN_SINE;
          else tan_jkl_i = cos_jkl / sin_jkl;
          double dvec_li[3] = {s.x[3 * i] - s.x[3 * l], s.x[3 * i + 1] - s.x[3 * l + 1], s.x[3 * i + 2] - s.x[3 * l + 2]};
          double r_li = sqrt(dot3(dvec_li, dvec_li));
          double dco_di[3], dco_dj[3], dco_dk[3], dco_dl[3];
          double omega = calc_omega(bij.dvec, r_ij, bjk.dvec, r_jk, bkl.dvec, r_kl, dvec_li, r_li, p_ijk, &p_jkl, dco_di,
                                    dco_dj, dco_dk, dco_dl);
          double cos_omega = cos(omega), cos2omega = cos(2. * omega), cos3omega = cos(3. * omega);
          double exp_tor1 = exp(fp.p_tor1 * SQR(2.0 - bjk.BO_pi - f11_DjDk));
          double exp_tor2_kl = exp(-p_tor2 * BOA_kl);
          double exp_cot2_kl = exp(-p_cot2 * SQR(BOA_kl - 1.5));
          double fn10 = (1.0 - exp_tor2_ij) * (1.0 - exp_tor2_jk) * (1.0 - exp_tor2_kl);
          double CV = 0.5 * (fp.V1 * (1.0 + cos_omega) + fp.V2 * exp_tor1 * (1.0 - cos2omega) + fp.V3 * (1.0 + cos3omega));
          double e_tor = fn10 * sin_ijk * sin_jkl * CV;
          e_tor_t += e_tor;
          double dfn11 = (-p_tor3 * exp_tor3_DjDk +
                          (p_tor3 * exp_tor3_DjDk - p_tor4 * exp_tor4_DjDk) * (2.0 + exp_tor3_DjDk) * exp_tor34_inv) *
                         exp_tor34_inv;
          double CEtors1 = sin_ijk * sin_jkl * CV;
          double CEtors2 = -fn10 * 2.0 * fp.p_tor1 * fp.V2 * exp_tor1 * (2.0 - bjk.BO_pi - f11_DjDk) *
                           (1.0 - SQR(cos_omega)) * sin_ijk * sin_jkl;
          double CEtors3 = CEtors2 * dfn11;
          double CEtors4 = CEtors1 * p_tor2 * exp_tor2_ij * (1.0 - exp_tor2_jk) * (1.0 - exp_tor2_kl);
          double CEtors5 = CEtors1 * p_tor2 * (1.0 - exp_tor2_ij) * exp_tor2_jk * (1.0 - exp_tor2_kl);
          double CEtors6 = CEtors1 * p_tor2 * (1.0 - exp_tor2_ij) * (1.0 - exp_tor2_jk) * exp_tor2_kl;
          double cmn = -fn10 * CV;
          double CEtors7 = cmn * sin_jkl * tan_ijk_i;
          double CEtors8 = cmn * sin_ijk * tan_jkl_i;
          double CEtors9 = fn10 * sin_ijk * sin_jkl *
                           (0.5 * fp.V1 - 2.0 * fp.V2 * exp_tor1 * cos_omega + 1.5 * fp.V3 * (cos2omega + 2.0 * SQR(cos_omega)));
          double fn12 = exp_cot2_ij * exp_cot2_jk * exp_cot2_kl;
          double e_con = fp.p_cot1 * fn12 * (1.0 + (SQR(cos_omega) - 1.0) * sin_ijk * sin_jkl);
          e_con_t += e_con;
          double Cconj = -2.0 * fn12 * fp.p_cot1 * p_cot2 * (1.0 + (SQR(cos_omega) - 1.0) * sin_ijk * sin_jkl);
          double CEconj1 = Cconj * (BOA_ij - 1.5e0);
          double CEconj2 = Cconj * (BOA_jk - 1.5e0);
          double CEconj3 = Cconj * (BOA_kl - 1.5e0);
          double CEconj4 = -fp.p_cot1 * fn12 * (SQR(cos_omega) - 1.0) * sin_jkl * tan_ijk_i;
          double CEconj5 = -fp.p_cot1 * fn12 * (SQR(cos_omega) - 1.0) * sin_ijk * tan_jkl_i;
          double CEconj6 = 2.0 * fp.p_cot1 * fn12 * cos_omega * sin_ijk * sin_jkl;

          atomic_add(bjk.Cdbopi, CEtors2);
          atomic_add(s.fCd[4 * j + 3], CEtors3);
          atomic_add(s.fCd[4 * k + 3], CEtors3);
          atomic_add(bij.Cdbo, (CEtors4 + CEconj1));
          atomic_add(bjk.Cdbo, (CEtors5 + CEconj2));
          atomic_add(bkl.Cdbo, (CEtors6 + CEconj3));
          add4(&s.fCd[4 * i], CEtors7 + CEconj4, p_ijk->dcos_dk);
          add4(&s.fCd[4 * j], CEtors7 + CEconj4, p_ijk->dcos_dj);
          add4(&s.fCd[4 * k], CEtors7 + CEconj4, p_ijk->dcos_di);
          add4(&s.fCd[4 * j], CEtors8 + CEconj5, p_jkl.dcos_di);
          add4(&s.fCd[4 * k], CEtors8 + CEconj5, p_jkl.dcos_dj);
          add4(&s.fCd[4 * l], CEtors8 + CEconj5, p_jkl.dcos_dk);
          add4(&s.fCd[4 * i], CEtors9 + CEconj6, dco_di);
          add4(&s.fCd[4 * j], CEtors9 + CEconj6, dco_dj);
          add4(&s.fCd[4 * k], CEtors9 + CEconj6, dco_dk);
          add4(&s.fCd[4 * l], CEtors9 + CEconj6, dco_dl);
        }
      }
    }
  }
  s.en.e_ang += e_ang_t; s.en.e_pen += e_pen_t; s.en.e_coa += e_coa_t; s.en.e_tor += e_tor_t; s.en.e_con += e_con_t;
}

// a11: Add_All_dBond_to_Forces_C  no-branch serial form reaxc_forces_sw64.c:500-587
void add_dbond_forces(System& s) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < s.N; i++) {
    double buf_c = 0.0;
    const double* dself = &s.dDeltap_self[3 * i];
    double fi[3] = {0, 0, 0};
    for (int pj = s.b_start[i]; pj < s.b_end[i]; pj++) {
      const Bond& b = s.bonds[pj];
      int j = b.nbr;
      const Bond& bs = s.bonds[b.sym];
      double C1dbo = b.C1dbo * (b.Cdbo + bs.Cdbo);
      double C2dbo = b.C2dbo * (b.Cdbo + bs.Cdbo);
      double C1dbopi = b.C1dbopi * (b.Cdbopi + bs.Cdbopi);
      double C2dbopi = b.C2dbopi * (b.Cdbopi + bs.Cdbopi);
      double C3dbopi = b.C3dbopi * (b.Cdbopi + bs.Cdbopi);
      double C1dbopi2 = b.C1dbopi2 * (b.Cdbopi2 + bs.Cdbopi2);
      double C2dbopi2 = b.C2dbopi2 * (b.Cdbopi2 + bs.Cdbopi2);
      double C3dbopi2 = b.C3dbopi2 * (b.Cdbopi2 + bs.Cdbopi2);
      double cdd = s.fCd[4 * i + 3] + s.fCd[4 * j + 3];
      double C1dDelta = b.C1dbo * cdd;
      double C2dDelta = b.C2dbo * cdd;
      for (int t = 0; t < 3; t++) {
        double temp = C1dbo * b.dBOp[t];
        temp += C2dbo * dself[t];
        temp += C1dDelta * b.dBOp[t];
        temp += C2dDelta * dself[t];
        temp += C1dbopi * b.dln_BOp_pi[t];
        temp += C2dbopi * b.dBOp[t];
        temp += C3dbopi * dself[t];
        temp += C1dbopi2 * b.dln_BOp_pi2[t];
        temp += C2dbopi2 * b.dBOp[t];
        temp += C3dbopi2 * dself[t];
        fi[t] += temp;
      }
      buf_c += (-C2dbo - C2dDelta - C3dbopi - C3dbopi2);
    }
    for (int t = 0; t < 3; t++) atomic_add(s.fCd[4 * i + t], fi[t]);
    for (int pj = s.b_start[i]; pj < s.b_end[i]; pj++) {
      const Bond& b = s.bonds[pj];
      add4(&s.fCd[4 * b.nbr], buf_c, b.dBOp);
    }
  }
}

// Compute_Forces order  reaxc_forces_sunway.cpp:1297-1365, Reset  reaxc_reset_tools_sunway.cpp:200-212
void compute_forces(System& s) {
  s.en = Energies();
  for (int t = 0; t < 6; t++) s.virial[t] = 0;
  s.fCd.assign((size_t)s.N * 4, 0.0);
  build_bond_list(s);
  build_hbond_list(s);
  nonbonded(s);
  bond_orders(s);
  bonds_atom_energy(s);
  hydrogen_bonds(s);
  valence_torsion(s);
  add_dbond_forces(s);
  // fdotr over all atoms (pair_reaxc_sunway.cpp:674-702), f = -fCd
  for (int i = 0; i < s.N; i++) {
    const double* xi = &s.x[3 * i];
    double f[3] = {-s.fCd[4 * i], -s.fCd[4 * i + 1], -s.fCd[4 * i + 2]};
    s.virial[0] += f[0] * xi[0]; s.virial[1] += f[1] * xi[1]; s.virial[2] += f[2] * xi[2];
    s.virial[3] += f[1] * xi[0]; s.virial[4] += f[2] * xi[0]; s.virial[5] += f[2] * xi[1];
  }
}

}  // namespace orc
