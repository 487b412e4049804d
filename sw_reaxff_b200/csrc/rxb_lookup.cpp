// Spline tables of the tabulated long-range mode (control file: tabulate_long_range N > 0), built on the host once per
// force field and uploaded for k_nonbonded_tab (rxb_nonbonded.cu).
//
// What the tables are is fixed by the reference's (commented-out, never executed) reaxc_lookup_sunway.cpp:157-285:
// knots r_k = k * dx, k = 1..N, dx = nonb_cut / N, one extra knot repeating the last value; natural cubic splines for
// CEvd and CEclmb, clamped ("complete") splines for the tapered e_vdW and e_ele with the end slopes that file uses;
// coefficient set k describes the interval ending at knot k+1 in the variable (r - r_{k+1}).
// This file is laid out for the kernel instead of for LR_lookup_table: one 128-byte record per (type pair, interval)
// holding the four coefficient sets a pair needs — {CEvd, CEclmb, e_vdW, e_ele} x {a, b, c, d} — so a force-only step
// touches the first 64 bytes and an energy step one full line.  Records exist for both (i,j) and (j,i).
#include <cmath>
#include <vector>

#include "rxb_params.h"

namespace rxb {

namespace {

struct Knot { double e_vdW, CEvd, e_ele, CEclmb; };

// LR_vdW_Coulomb (reaxc_nonbonded_sunway.cpp:573-668): tapered pair terms per unit q_i q_j at distance r
Knot knot_values(const ForceField& ff, int i, int j, double r) {
  const Control& c = ff.ctl;
  const PairPar& tw = ff.pair[(size_t)i * ff.nt + j];
  const double p = ff.gp[28], pi = 1.0 / p;
  const double* Tp = c.Tap;
  double T = Tp[7];
  for (int k = 6; k >= 0; k--) T = T * r + Tp[k];
  double dT = 7 * Tp[7] * r + 6 * Tp[6];
  for (int k = 5; k >= 2; k--) dT = dT * r + k * Tp[k];
  dT += Tp[1] / r;
  Knot y;
  double e, ce;
  if (c.vdw_type == 1 || c.vdw_type == 3) {
    const double powr = pow(r, p), powgi = pow(1.0 / tw.gamma_w, p);
    const double fn13 = pow(powr + powgi, pi);
    const double exp1 = exp(tw.alpha * (1.0 - fn13 / tw.r_vdW)), exp2 = exp(0.5 * tw.alpha * (1.0 - fn13 / tw.r_vdW));
    const double dfn13 = pow(powr + powgi, pi - 1.0) * pow(r, p - 2.0);
    e = T * tw.D * (exp1 - 2.0 * exp2);
    ce = dT * tw.D * (exp1 - 2.0 * exp2) - T * tw.D * (tw.alpha / tw.r_vdW) * (exp1 - exp2) * dfn13;
  } else {
    const double exp1 = exp(tw.alpha * (1.0 - r / tw.r_vdW)), exp2 = exp(0.5 * tw.alpha * (1.0 - r / tw.r_vdW));
    e = T * tw.D * (exp1 - 2.0 * exp2);
    ce = dT * tw.D * (exp1 - 2.0 * exp2) - T * tw.D * (tw.alpha / tw.r_vdW) * (exp1 - exp2) / r;
  }
  if (c.vdw_type == 2 || c.vdw_type == 3) {
    const double e_core = tw.ecore * exp(tw.acore * (1.0 - (r / tw.rcore)));
    const double de_core = -(tw.acore / tw.rcore) * e_core;
    e += T * e_core;
    ce += dT * e_core + T * de_core / r;
    if (c.lgflag) {
      const double r5 = pow(r, 5.0), r6 = pow(r, 6.0), re6 = pow(tw.lgre, 6.0);
      const double e_lg = -(tw.lgcij / (r6 + re6));
      const double de_lg = -6.0 * e_lg * r5 / (r6 + re6);
      e += T * e_lg;
      ce += dT * e_lg + T * de_lg / r;
    }
  }
  y.e_vdW = e; y.CEvd = ce;
  const double g1 = r * r * r + tw.gamma, g3 = pow(g1, 0.33333333333333);
  y.e_ele = kCeleConst * (T / g3);
  y.CEclmb = kCeleConst * (dT - T * r / g1) / g3;
  return y;
}

// Thomas algorithm on the (sub, diag, super, rhs) system of size m; super and rhs are overwritten
void thomas(const double* sub, const double* diag, double* sup, double* rhs, double* x, int m) {
  sup[0] /= diag[0];
  rhs[0] /= diag[0];
  for (int k = 1; k < m; k++) {
    const double piv = diag[k] - sup[k - 1] * sub[k];
    sup[k] /= piv;
    rhs[k] = (rhs[k] - rhs[k - 1] * sub[k]) / piv;
  }
  x[m - 1] = rhs[m - 1];
  for (int k = m - 2; k >= 0; k--) x[k] = rhs[k] - sup[k] * x[k + 1];
}

// second derivatives v[0..m-1] of the spline through f[0..m-1] on a uniform grid of spacing h
// clamped = false: natural ends (v = 0 at both ends, interior system only)
// clamped = true : the reference's "complete" system with end slopes s0, s1 (incl. its last-row right-hand side)
std::vector<double> second_derivatives(const std::vector<double>& f, double h, bool clamped, double s0, double s1) {
  const int m = (int)f.size();
  std::vector<double> sub(m), diag(m), sup(m), rhs(m), v(m, 0.0);
  for (int k = 1; k < m - 1; k++) rhs[k] = 6 * ((f[k + 1] - f[k]) / h - (f[k] - f[k - 1]) / h);
  if (!clamped) {
    sub[0] = sub[1] = sub[m - 1] = 0;
    for (int k = 2; k < m - 1; k++) sub[k] = h;
    diag[0] = diag[m - 1] = 0;
    for (int k = 1; k < m - 1; k++) diag[k] = 2 * (h + h);
    sup[0] = sup[m - 2] = sup[m - 1] = 0;
    for (int k = 1; k < m - 2; k++) sup[k] = h;
    rhs[0] = rhs[m - 1] = 0;
    thomas(&sub[1], &diag[1], &sup[1], &rhs[1], &v[1], m - 2);
  } else {
    sub[0] = 0;
    for (int k = 1; k < m; k++) sub[k] = h;
    diag[0] = 2 * h;
    for (int k = 1; k < m; k++) diag[k] = 2 * (h + h);
    sup[m - 1] = 0;
    for (int k = 0; k < m - 1; k++) sup[k] = h;
    rhs[0] = 6 * (f[1] - f[0]) / h - 6 * s0;
    rhs[m - 1] = 6 * s1 - 6 * (f[m - 1] - f[m - 2] / h);   // as written in reaxc_lookup_sunway.cpp:143
    thomas(&sub[0], &diag[0], &sup[0], &rhs[0], &v[0], m);
  }
  return v;
}

}  // namespace

std::vector<double> ForceField::lookup_tables(int* n_out, double* dx_out) const {
  const int N = ctl.tabulate, n = N + 2, m = N + 1;   // m knots enter each spline: k = 1..N and the repeated end knot
  const double dx = ctl.nonb_cut / N;
  std::vector<double> out((size_t)nt * nt * n * 16, 0.0);
  std::vector<double> f[4];
  for (auto& a : f) a.resize(m);
  for (int i = 0; i < nt; i++)
    for (int j = i; j < nt; j++) {
      Knot first{};
      for (int k = 1; k <= N; k++) {
        const Knot y = knot_values(*this, i, j, k * dx);
        if (k == 1) first = y;
        f[0][k - 1] = y.CEvd; f[1][k - 1] = y.CEclmb; f[2][k - 1] = y.e_vdW; f[3][k - 1] = y.e_ele;
      }
      for (auto& a : f) a[m - 1] = a[m - 2];
      const double slopes[4][2] = {{0, 0}, {0, 0}, {first.CEvd, f[0][m - 2]}, {first.CEclmb, f[3][m - 2]}};
      for (int t = 0; t < 4; t++) {
        const std::vector<double> v = second_derivatives(f[t], dx, t >= 2, slopes[t][0], slopes[t][1]);
        for (int k = 1; k < m; k++) {   // coefficient set with table index k (the reference stores it at [1 + (k-1)])
          const double a = f[t][k], b = (f[t][k] - f[t][k - 1]) / dx + dx * (2 * v[k] + v[k - 1]) / 6, c = v[k] / 2,
                       d = (v[k] - v[k - 1]) / (6 * dx);
          for (int rep = 0; rep < 2; rep++) {
            const int p = rep ? j * nt + i : i * nt + j;
            double* o = &out[((size_t)p * n + k) * 16 + 4 * t];
            o[0] = a; o[1] = b; o[2] = c; o[3] = d;
          }
        }
      }
    }
  *n_out = n;
  *dx_out = dx;
  return out;
}

}  // namespace rxb
