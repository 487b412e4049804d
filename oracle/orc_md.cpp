// ORACLE — TEST INFRASTRUCTURE ONLY.  See orc_md.h.
#include "orc_md.h"

#include <algorithm>
#include <cmath>
#include <cstdio>

namespace orc {

static const double FTM2V = 1.0 / 48.88821291 / 48.88821291;  // LAMMPS units real
static const double MVV2E = 48.88821291 * 48.88821291;

void Box::set(double xprd, double yprd, double zprd, double xy, double xz, double yz) {
  h[0] = xprd; h[1] = yprd; h[2] = zprd; h[3] = yz; h[4] = xz; h[5] = xy;
  h_inv[0] = 1.0 / h[0]; h_inv[1] = 1.0 / h[1]; h_inv[2] = 1.0 / h[2];
  h_inv[3] = -h[3] / (h[1] * h[2]);
  h_inv[4] = (h[3] * h[5] - h[1] * h[4]) / (h[0] * h[1] * h[2]);
  h_inv[5] = -h[5] / (h[0] * h[1]);
}
void Box::x2lamda(const double* x, double* l) const {
  double d0 = x[0] - lo[0], d1 = x[1] - lo[1], d2 = x[2] - lo[2];
  l[0] = h_inv[0] * d0 + h_inv[5] * d1 + h_inv[4] * d2;
  l[1] = h_inv[1] * d1 + h_inv[3] * d2;
  l[2] = h_inv[2] * d2;
}
void Box::shift(int sx, int sy, int sz, double* d) const {
  d[0] = sx * h[0] + sy * h[5] + sz * h[4];
  d[1] = sy * h[1] + sz * h[3];
  d[2] = sz * h[2];
}
void Box::cutghost_lamda(double cut, double* cg) const {  // LAMMPS Comm::setup, triclinic
  cg[0] = cut * sqrt(h_inv[0] * h_inv[0] + h_inv[5] * h_inv[5] + h_inv[4] * h_inv[4]);
  cg[1] = cut * sqrt(h_inv[1] * h_inv[1] + h_inv[3] * h_inv[3]);
  cg[2] = cut * h_inv[2];
}

void MD::remap() {
  for (int i = 0; i < nlocal; i++) {
    double l[3];
    box.x2lamda(&sys.x[3 * i], l);
    int s[3] = {(int)floor(l[0]), (int)floor(l[1]), (int)floor(l[2])};
    if (s[0] || s[1] || s[2]) {
      double d[3];
      box.shift(s[0], s[1], s[2], d);
      for (int t = 0; t < 3; t++) sys.x[3 * i + t] -= d[t];
    }
  }
}

void MD::make_ghosts() {
  double cg[3];
  box.cutghost_lamda(cutneigh, cg);
  int m[3] = {(int)ceil(cg[0]), (int)ceil(cg[1]), (int)ceil(cg[2])};
  sys.x.resize((size_t)3 * nlocal);
  sys.type.resize(nlocal); sys.tag.resize(nlocal); sys.q.resize(nlocal);
  ghost_owner.clear(); ghost_shift.clear();
  std::vector<double> lam((size_t)3 * nlocal);
  for (int i = 0; i < nlocal; i++) box.x2lamda(&sys.x[3 * i], &lam[3 * i]);
  for (int sz = -m[2]; sz <= m[2]; sz++)
    for (int sy = -m[1]; sy <= m[1]; sy++)
      for (int sx = -m[0]; sx <= m[0]; sx++) {
        if (!sx && !sy && !sz) continue;
        double d[3];
        box.shift(sx, sy, sz, d);
        for (int i = 0; i < nlocal; i++) {
          double l0 = lam[3 * i] + sx, l1 = lam[3 * i + 1] + sy, l2 = lam[3 * i + 2] + sz;
          if (l0 >= -cg[0] && l0 < 1.0 + cg[0] && l1 >= -cg[1] && l1 < 1.0 + cg[1] && l2 >= -cg[2] && l2 < 1.0 + cg[2]) {
            ghost_owner.push_back(i);
            ghost_shift.push_back(sx); ghost_shift.push_back(sy); ghost_shift.push_back(sz);
            sys.x.push_back(sys.x[3 * i] + d[0]); sys.x.push_back(sys.x[3 * i + 1] + d[1]); sys.x.push_back(sys.x[3 * i + 2] + d[2]);
            sys.type.push_back(sys.type[i]); sys.tag.push_back(sys.tag[i]); sys.q.push_back(sys.q[i]);
          }
        }
      }
  sys.n = nlocal;
  sys.N = nlocal + (int)ghost_owner.size();
}

void MD::forward_x() {
  for (int g = 0; g < (int)ghost_owner.size(); g++) {
    double d[3];
    box.shift(ghost_shift[3 * g], ghost_shift[3 * g + 1], ghost_shift[3 * g + 2], d);
    int o = ghost_owner[g], i = nlocal + g;
    for (int t = 0; t < 3; t++) sys.x[3 * i + t] = sys.x[3 * o + t] + d[t];
  }
}

// binned full neighbour list with ghost rows; rows sorted by neighbour index
void build_full_neighbor_list(System& s, double cutneigh) {
  const int N = s.N;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < N; i++)
    for (int t = 0; t < 3; t++) { lo[t] = std::min(lo[t], s.x[3 * i + t]); hi[t] = std::max(hi[t], s.x[3 * i + t]); }
  int nb[3];
  for (int t = 0; t < 3; t++) nb[t] = std::max(1, (int)((hi[t] - lo[t]) / cutneigh));
  double inv[3];
  for (int t = 0; t < 3; t++) inv[t] = nb[t] / std::max(hi[t] - lo[t], 1e-12);
  auto binof = [&](const double* x, int* b) {
    for (int t = 0; t < 3; t++) { b[t] = (int)((x[t] - lo[t]) * inv[t]); b[t] = std::min(std::max(b[t], 0), nb[t] - 1); }
  };
  std::vector<int> head((size_t)nb[0] * nb[1] * nb[2], -1), next(N, -1);
  for (int i = N - 1; i >= 0; i--) {
    int b[3];
    binof(&s.x[3 * i], b);
    int id = (b[2] * nb[1] + b[1]) * nb[0] + b[0];
    next[i] = head[id];
    head[id] = i;
  }
  std::vector<std::vector<int>> rows(N);
  const double c2 = cutneigh * cutneigh;
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < N; i++) {
    int b[3];
    binof(&s.x[3 * i], b);
    std::vector<int>& row = rows[i];
    for (int bz = std::max(0, b[2] - 1); bz <= std::min(nb[2] - 1, b[2] + 1); bz++)
      for (int by = std::max(0, b[1] - 1); by <= std::min(nb[1] - 1, b[1] + 1); by++)
        for (int bx = std::max(0, b[0] - 1); bx <= std::min(nb[0] - 1, b[0] + 1); bx++)
          for (int j = head[(bz * nb[1] + by) * nb[0] + bx]; j >= 0; j = next[j]) {
            if (j == i) continue;
            double dx = s.x[3 * i] - s.x[3 * j], dy = s.x[3 * i + 1] - s.x[3 * j + 1], dz = s.x[3 * i + 2] - s.x[3 * j + 2];
            if (dx * dx + dy * dy + dz * dz <= c2) row.push_back(j);
          }
    std::sort(row.begin(), row.end());
  }
  s.nb_off.assign(N + 1, 0);
  for (int i = 0; i < N; i++) s.nb_off[i + 1] = s.nb_off[i] + (long)rows[i].size();
  s.nb.resize(s.nb_off[N]);
  for (int i = 0; i < N; i++) std::copy(rows[i].begin(), rows[i].end(), s.nb.begin() + s.nb_off[i]);
}

void MD::build_neighbors() { build_full_neighbor_list(sys, cutneigh); }

void MD::force() {
  if (qeq_on) qeq.pre_force(sys, ghost_owner);
  compute_forces(sys);
  f.assign((size_t)3 * nlocal, 0.0);
  for (int i = 0; i < nlocal; i++)
    for (int t = 0; t < 3; t++) f[3 * i + t] = -sys.fCd[4 * i + t];
  for (int g = 0; g < (int)ghost_owner.size(); g++)  // reverse_comm
    for (int t = 0; t < 3; t++) f[3 * ghost_owner[g] + t] += -sys.fCd[4 * (nlocal + g) + t];
}

void MD::setup() {
  remap();
  make_ghosts();
  build_neighbors();
  ago = 0;
  force();
}

void MD::run(int nsteps) {
  const double dtv = dt, dtf = 0.5 * dt * FTM2V;
  for (int step = 0; step < nsteps; step++) {
    ntimestep++;
    for (int i = 0; i < nlocal; i++) {
      double dtfm = dtf / mass[ltype[i]];
      for (int t = 0; t < 3; t++) {
        v[3 * i + t] += dtfm * f[3 * i + t];
        sys.x[3 * i + t] += dtv * v[3 * i + t];
      }
    }
    ago++;
    if (ago % every == 0) {
      // history and charges are per local atom; ghosts are regenerated
      std::vector<double> qsave(sys.q.begin(), sys.q.begin() + nlocal);
      remap();
      make_ghosts();
      for (int i = 0; i < nlocal; i++) sys.q[i] = qsave[i];
      for (int g = 0; g < (int)ghost_owner.size(); g++) sys.q[nlocal + g] = sys.q[ghost_owner[g]];
      build_neighbors();
      ago = 0;
    } else {
      forward_x();
    }
    force();
    for (int i = 0; i < nlocal; i++) {
      double dtfm = dtf / mass[ltype[i]];
      for (int t = 0; t < 3; t++) v[3 * i + t] += dtfm * f[3 * i + t];
    }
  }
}

double MD::kinetic() const {
  double ke = 0;
  for (int i = 0; i < nlocal; i++)
    ke += mass[ltype[i]] * (v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]);
  return 0.5 * MVV2E * ke;
}

double MD::potential() const {
  const Energies& e = sys.en;
  return e.e_bond + e.e_ov + e.e_un + e.e_lp + e.e_ang + e.e_pen + e.e_coa + e.e_hb + e.e_tor + e.e_con + e.e_vdW +
         e.e_ele + e.e_pol;
}

}  // namespace orc
