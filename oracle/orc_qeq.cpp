// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under sw_reaxff_b200/ may include, link or call this.
//
// CPU restatement of fix qeq/reax as implemented by SW_REAXFF (fix_qeq_reax_sunway.cpp):
//   init_shielding :440-454, init_taper :458-484, calculate_H :965-981,
//   H build (full symmetric rows, no tag filter) fix_qeq_reax_sw64.c:147-189 (serial comment),
//   init_matvec :696-717 (s cubic / t quadratic extrapolation), sparse_matvec :1601-1622,
//   CG_v2 :983-1167 (pipelined Jacobi-PCG, imax 200), calculate_Q :1697-1755.
// Single rank: forward_comm_fix == copy owner -> ghost (ghost_owner map).
// Pinned (tests/test_oracle_vs_ref.py::test_qeq_equals_reference_fix_qeq_reax): the reference's FixQEqReaxSunway, compiled
// unmodified against a LAMMPS-core stand-in (oracle/ref/ref_qeq.cpp), gives the same CG_v2 iteration counts, charges
// (1e-10), s/t and histories (1e-9) and H rows (1e-13) on identical inputs.  Its H build and SpMV exist only as Sunway
// slave-core kernels; the harness supplies serial loops for those two.
#include <cmath>
#include <cstdio>

#include "orc_qeq.h"

namespace orc {

static const double EV_TO_KCAL_PER_MOL = 14.4;

void QEq::init(const Params& P, double swa_, double swb_, double tol_) {
  swa = swa_; swb = swb_; tolerance = tol_;
  const int nt = P.nt;
  chi.assign(nt, 0); eta.assign(nt, 0); gamma.assign(nt, 0);
  for (int i = 0; i < nt; i++) { chi[i] = P.sbp[i].chi; eta[i] = P.sbp[i].eta; gamma[i] = P.sbp[i].gamma; }
  shld.assign((size_t)nt * nt, 0);
  for (int i = 0; i < nt; i++)
    for (int j = 0; j < nt; j++) shld[i * nt + j] = pow(gamma[i] * gamma[j], -1.5);
  double d7 = pow(swb - swa, 7);
  double swa2 = swa * swa, swa3 = swa2 * swa, swb2 = swb * swb, swb3 = swb2 * swb;
  Tap[7] = 20.0 / d7;
  Tap[6] = -70.0 * (swa + swb) / d7;
  Tap[5] = 84.0 * (swa2 + 3.0 * swa * swb + swb2) / d7;
  Tap[4] = -35.0 * (swa3 + 9.0 * swa2 * swb + 9.0 * swa * swb2 + swb3) / d7;
  Tap[3] = 140.0 * (swa3 * swb + 3.0 * swa2 * swb2 + swa * swb3) / d7;
  Tap[2] = -210.0 * (swa3 * swb2 + swa2 * swb3) / d7;
  Tap[1] = 140.0 * swa3 * swb3 / d7;
  Tap[0] = (-35.0 * swa3 * swb2 * swb2 + 21.0 * swa2 * swb3 * swb2 + 7.0 * swa * swb3 * swb3 + swb3 * swb3 * swb) / d7;
}

double QEq::calculate_H(double r, double g) const {
  double Taper = Tap[7] * r + Tap[6];
  Taper = Taper * r + Tap[5];
  Taper = Taper * r + Tap[4];
  Taper = Taper * r + Tap[3];
  Taper = Taper * r + Tap[2];
  Taper = Taper * r + Tap[1];
  Taper = Taper * r + Tap[0];
  double denom = r * r * r + g;
  denom = pow(denom, 0.3333333333333);
  return Taper * EV_TO_KCAL_PER_MOL / denom;
}

void QEq::compute_H(const System& s) {
  const int n = s.n, nt = s.prm.nt;
  H_off.assign(n + 1, 0);
  for (int i = 0; i < n; i++) H_off[i + 1] = H_off[i] + (s.nb_off[i + 1] - s.nb_off[i]);
  H_j.assign(H_off[n], 0); H_val.assign(H_off[n], 0.0); H_num.assign(n, 0);
  const double swbsq = swb * swb;
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; i++) {
    long o = H_off[i];
    int cnt = 0;
    for (long pj = s.nb_off[i]; pj < s.nb_off[i + 1]; pj++) {
      int j = s.nb[pj];
      double dx = s.x[3 * j] - s.x[3 * i], dy = s.x[3 * j + 1] - s.x[3 * i + 1], dz = s.x[3 * j + 2] - s.x[3 * i + 2];
      double r_sqr = dx * dx + dy * dy + dz * dz;
      if (r_sqr <= swbsq) {
        H_j[o + cnt] = j;
        H_val[o + cnt] = calculate_H(sqrt(r_sqr), shld[s.type[i] * nt + s.type[j]]);
        cnt++;
      }
    }
    H_num[i] = cnt;
  }
}

void QEq::matvec(const System& s, const std::vector<double>& x, std::vector<double>& b) const {
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < s.n; i++) {
    double acc = eta[s.type[i]] * x[i];
    for (long p = H_off[i]; p < H_off[i] + H_num[i]; p++) acc += H_val[p] * x[H_j[p]];
    b[i] = acc;
  }
}

void QEq::forward(const System& s, const std::vector<int>& ghost_owner, std::vector<double>& v) const {
  for (int g = s.n; g < s.N; g++) v[g] = v[ghost_owner[g - s.n]];
}

// CG_v2, fix_qeq_reax_sunway.cpp:983-1167
int QEq::cg(const System& s, const std::vector<int>& go, const std::vector<double>& b, std::vector<double>& x) {
  const int nn = s.n, N = s.N, imax = 200;
  std::vector<double> q(N, 0), r(N, 0), u(N, 0), d(N, 0), w(N, 0), m(N, 0), p(N, 0), ss(N, 0), v(N, 0), z(N, 0);
  matvec(s, x, q);
  for (int j = 0; j < nn; j++) { r[j] = b[j] - 1 * q[j]; u[j] = r[j] * Hdia_inv[j]; d[j] = u[j]; }
  forward(s, go, d);
  matvec(s, d, q);
  for (int j = 0; j < nn; j++) { w[j] = q[j]; m[j] = q[j] * Hdia_inv[j]; d[j] = m[j]; }
  forward(s, go, d);
  matvec(s, d, q);  // spawn ... join
  double my0 = 0, my1 = 0, my2 = 0;
  for (int j = 0; j < nn; j++) {
    p[j] = u[j]; ss[j] = w[j]; v[j] = m[j];
    my0 += b[j] * b[j]; my1 += u[j] * r[j]; my2 += u[j] * w[j];
  }
  double b_norm = sqrt(my0), sig_old = my1, deta = my2, heta = deta;
  double alpha = sig_old / deta, beta;
  double dot0 = sig_old, dot1;
  for (int j = 0; j < nn; j++) z[j] = q[j];
  int i;
  for (i = 1; i < imax && sqrt(dot0) / b_norm > tolerance; ++i) {
    dot0 = dot1 = 0.0;
    for (int j = 0; j < nn; j++) {
      r[j] -= alpha * ss[j];
      u[j] -= alpha * v[j];
      w[j] -= alpha * z[j];
      dot0 += u[j] * r[j];
      dot1 += u[j] * w[j];
      d[j] = w[j] * Hdia_inv[j];
    }
    forward(s, go, d);
    matvec(s, d, q);
    beta = dot0 / sig_old;
    heta = dot1 - beta * beta * heta;
    for (int j = 0; j < nn; j++) {
      x[j] += alpha * p[j];
      p[j] = u[j] + p[j] * beta;
      ss[j] = w[j] + ss[j] * beta;
      v[j] = d[j] + v[j] * beta;
    }
    alpha = dot0 / heta;
    for (int j = 0; j < nn; j++) z[j] = q[j] + z[j] * beta;
    sig_old = dot0;
  }
  return i;
}

// pre_force :539-600
void QEq::pre_force(System& s, const std::vector<int>& ghost_owner) {
  const int n = s.n, N = s.N;
  if ((int)s_hist.size() != 5 * n) { s_hist.assign((size_t)5 * n, 0.0); t_hist.assign((size_t)5 * n, 0.0); }
  compute_H(s);
  Hdia_inv.assign(N, 0); b_s.assign(N, 0); b_t.assign(N, 0); sv.assign(N, 0); tv.assign(N, 0);
  for (int i = 0; i < n; i++) {
    int ti = s.type[i];
    Hdia_inv[i] = 1. / eta[ti];
    b_s[i] = -chi[ti];
    b_t[i] = -1.0;
    const double *sh = &s_hist[5 * i], *th = &t_hist[5 * i];
    tv[i] = th[2] + 3 * (th[0] - th[1]);
    sv[i] = 4 * (sh[0] + sh[2]) - (6 * sh[1] + sh[3]);
  }
  forward(s, ghost_owner, sv);
  forward(s, ghost_owner, tv);
  matvecs_s = cg(s, ghost_owner, b_s, sv);
  matvecs_t = cg(s, ghost_owner, b_t, tv);
  // calculate_Q
  double s_sum = 0, t_sum = 0;
  for (int i = 0; i < n; i++) { s_sum += sv[i]; t_sum += tv[i]; }
  double u = s_sum / t_sum;
  for (int i = 0; i < n; i++) {
    s.q[i] = sv[i] - u * tv[i];
    for (int k = 4; k > 0; --k) { s_hist[5 * i + k] = s_hist[5 * i + k - 1]; t_hist[5 * i + k] = t_hist[5 * i + k - 1]; }
    s_hist[5 * i] = sv[i];
    t_hist[5 * i] = tv[i];
  }
  forward(s, ghost_owner, s.q);
}

}  // namespace orc
