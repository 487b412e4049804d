// ORACLE — TEST INFRASTRUCTURE ONLY.  See orc_qeq.cpp for the reference ranges restated.
#pragma once
#include <vector>

#include "orc_system.h"

namespace orc {

struct QEq {
  double swa = 0, swb = 10, tolerance = 1e-6;
  double Tap[8];
  std::vector<double> chi, eta, gamma, shld;
  std::vector<long> H_off;
  std::vector<int> H_num, H_j;
  std::vector<double> H_val;
  std::vector<double> Hdia_inv, b_s, b_t, sv, tv;
  std::vector<double> s_hist, t_hist;  // [n][5]
  int matvecs_s = 0, matvecs_t = 0;

  void init(const Params& P, double swa, double swb, double tol);
  double calculate_H(double r, double gamma) const;
  void compute_H(const System& s);
  void matvec(const System& s, const std::vector<double>& x, std::vector<double>& b) const;
  void forward(const System& s, const std::vector<int>& ghost_owner, std::vector<double>& v) const;
  int cg(const System& s, const std::vector<int>& ghost_owner, const std::vector<double>& b, std::vector<double>& x);
  void pre_force(System& s, const std::vector<int>& ghost_owner);
};

}  // namespace orc
