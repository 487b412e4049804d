// mini_lammps.cpp — see mini_lammps.h.  Verlet::setup / Verlet::run order as in LAMMPS (and SURVEY.md §1):
//   initial_integrate -> [decide: pbc, borders, neighbor | forward_comm] -> force_clear -> pre_force (qeq) ->
//   pair->compute -> reverse_comm -> final_integrate -> thermo.
#include "mini_lammps.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

#include "styles_b200.h"

namespace LAMMPS_MINI {

// ---------------------------------------------------------------------------------------------------------------
void Domain::set_box(double xprd, double yprd, double zprd, double xy, double xz, double yz) {
  h[0] = xprd; h[1] = yprd; h[2] = zprd; h[3] = yz; h[4] = xz; h[5] = xy;
  h_inv[0] = 1.0 / h[0]; h_inv[1] = 1.0 / h[1]; h_inv[2] = 1.0 / h[2];
  h_inv[3] = -h[3] / (h[1] * h[2]);
  h_inv[4] = (h[3] * h[5] - h[1] * h[4]) / (h[0] * h[1] * h[2]);
  h_inv[5] = -h[5] / (h[0] * h[1]);
}
void Domain::x2lamda(const double* x, double* l) const {
  const double d0 = x[0] - boxlo[0], d1 = x[1] - boxlo[1], d2 = x[2] - boxlo[2];
  l[0] = h_inv[0] * d0 + h_inv[5] * d1 + h_inv[4] * d2;
  l[1] = h_inv[1] * d1 + h_inv[3] * d2;
  l[2] = h_inv[2] * d2;
}
void Domain::image_shift(int sx, int sy, int sz, double* d) const {
  d[0] = sx * h[0] + sy * h[5] + sz * h[4];
  d[1] = sy * h[1] + sz * h[3];
  d[2] = sz * h[2];
}
void Domain::remap(Atom& a) const {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < a.nlocal; i++) {
    double l[3];
    x2lamda(&a.x[3 * i], l);
    const int s[3] = {(int)floor(l[0]), (int)floor(l[1]), (int)floor(l[2])};
    if (s[0] || s[1] || s[2]) {
      double d[3];
      image_shift(s[0], s[1], s[2], d);
      for (int t = 0; t < 3; t++) a.x[3 * i + t] -= d[t];
    }
  }
}

// Ghost atoms = periodic images within `cut` of the box, ordered shift-major (z, y, x shift loops outermost, atoms
// innermost) like the oracle's stand-in.  Selection runs per shift in parallel; the fill keeps that order.
void Comm::borders(const Domain& dom, Atom& a, double cut) {
  const double* hi = dom.h_inv;
  const double cg[3] = {cut * sqrt(hi[0] * hi[0] + hi[5] * hi[5] + hi[4] * hi[4]), cut * sqrt(hi[1] * hi[1] + hi[3] * hi[3]), cut * hi[2]};
  const int m[3] = {(int)ceil(cg[0]), (int)ceil(cg[1]), (int)ceil(cg[2])};
  const int n = a.nlocal;
  std::vector<double> lam((size_t)3 * n);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) dom.x2lamda(&a.x[3 * i], &lam[3 * i]);
  struct Shift { int s[3]; double d[3]; std::vector<int> who; };
  std::vector<Shift> shifts;
  for (int sz = -m[2]; sz <= m[2]; sz++)
    for (int sy = -m[1]; sy <= m[1]; sy++)
      for (int sx = -m[0]; sx <= m[0]; sx++) {
        if (!sx && !sy && !sz) continue;
        Shift sh;
        sh.s[0] = sx; sh.s[1] = sy; sh.s[2] = sz;
        dom.image_shift(sx, sy, sz, sh.d);
        shifts.push_back(std::move(sh));
      }
  const int ns = (int)shifts.size();
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < ns; k++) {
    Shift& sh = shifts[k];
    for (int i = 0; i < n; i++) {
      const double l0 = lam[3 * i] + sh.s[0], l1 = lam[3 * i + 1] + sh.s[1], l2 = lam[3 * i + 2] + sh.s[2];
      if (l0 >= -cg[0] && l0 < 1.0 + cg[0] && l1 >= -cg[1] && l1 < 1.0 + cg[1] && l2 >= -cg[2] && l2 < 1.0 + cg[2]) sh.who.push_back(i);
    }
  }
  std::vector<size_t> off(ns + 1, 0);
  for (int k = 0; k < ns; k++) off[k + 1] = off[k] + shifts[k].who.size();
  const size_t ng = off[ns], nall = n + ng;
  auto grow = [&](auto& v, size_t want) { if (v.capacity() < want) v.reserve(want + want / 4); v.resize(want); };
  grow(a.x, 3 * nall); grow(a.type, nall); grow(a.tag, nall); grow(a.q, nall);
  grow(ghost_owner, ng); grow(ghost_shift, 3 * ng);
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < ns; k++) {
    const Shift& sh = shifts[k];
    size_t g = off[k];
    for (int i : sh.who) {
      ghost_owner[g] = i;
      for (int t = 0; t < 3; t++) { ghost_shift[3 * g + t] = sh.d[t]; a.x[3 * (n + g) + t] = a.x[3 * i + t] + sh.d[t]; }
      a.type[n + g] = a.type[i]; a.tag[n + g] = a.tag[i]; a.q[n + g] = a.q[i];
      g++;
    }
  }
  block_off.assign(off.begin(), off.end());
  a.nghost = (int)ng;
  grow(a.f, 3 * nall);
  std::fill(a.f.begin(), a.f.end(), 0.0);
}
void Comm::forward_comm(Atom& a) const {
  const int n = a.nlocal;
#pragma omp parallel for schedule(static)
  for (int g = 0; g < a.nghost; g++)
    for (int t = 0; t < 3; t++) a.x[3 * (n + g) + t] = a.x[3 * ghost_owner[g] + t] + ghost_shift[3 * g + t];
}
void Comm::reverse_comm(Atom& a) const {
  // Ghosts of one image shift have distinct owners, so a shift block is summed in parallel; blocks run in order, which
  // keeps the per-owner summation order of the serial loop (bitwise the same result).
  const int n = a.nlocal;
  double* f = a.f.data();
  for (size_t b = 0; b + 1 < block_off.size(); b++) {
    const long g0 = (long)block_off[b], g1 = (long)block_off[b + 1];
#pragma omp parallel for schedule(static) if (g1 - g0 > 4096)
    for (long g = g0; g < g1; g++)
      for (int t = 0; t < 3; t++) f[3 * ghost_owner[g] + t] += f[3 * (n + g) + t];
  }
}
void Comm::forward_comm_q(Atom& a) const {
  for (int g = 0; g < a.nghost; g++) a.q[a.nlocal + g] = a.q[ghost_owner[g]];
}

// ---------------------------------------------------------------------------------------------------------------
namespace {
struct Parser {
  const std::string& s;
  size_t p = 0;
  explicit Parser(const std::string& str) : s(str) {}
  void ws() { while (p < s.size() && isspace((unsigned char)s[p])) p++; }
  double number() {
    ws();
    size_t q = p;
    double v = strtod(s.c_str() + p, nullptr);
    char* end;
    strtod(s.c_str() + p, &end);
    p = end - s.c_str();
    if (p == q) throw std::runtime_error("ERROR: invalid expression '" + s + "'");
    return v;
  }
  double factor() {
    ws();
    if (p < s.size() && s[p] == '(') { p++; double v = expr(); ws(); if (p < s.size() && s[p] == ')') p++; return v; }
    if (p < s.size() && s[p] == '-') { p++; return -factor(); }
    if (p < s.size() && s[p] == '+') { p++; return factor(); }
    return number();
  }
  double term() {
    double v = factor();
    for (;;) {
      ws();
      if (p < s.size() && s[p] == '*') { p++; v *= factor(); }
      else if (p < s.size() && s[p] == '/') { p++; v /= factor(); }
      else return v;
    }
  }
  double expr() {
    double v = term();
    for (;;) {
      ws();
      if (p < s.size() && s[p] == '+') { p++; v += term(); }
      else if (p < s.size() && s[p] == '-') { p++; v -= term(); }
      else return v;
    }
  }
};

std::vector<std::string> words(const std::string& line) {
  std::vector<std::string> w;
  std::istringstream is(line);
  std::string t;
  while (is >> t) {
    if (t[0] == '#') break;
    w.push_back(t);
  }
  return w;
}

// Park-Miller minimal standard generator + Marsaglia polar Gaussian (own RNG: not bit-compatible with LAMMPS' velocity create)
struct Rng {
  long seed;
  explicit Rng(long s) : seed(s) {}
  double uniform() {
    const long IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836;
    long k = seed / IQ;
    seed = IA * (seed - k * IQ) - IR * k;
    if (seed < 0) seed += IM;
    return seed * (1.0 / IM);
  }
  double gaussian() {
    double v1, v2, rsq;
    do { v1 = 2.0 * uniform() - 1.0; v2 = 2.0 * uniform() - 1.0; rsq = v1 * v1 + v2 * v2; } while (rsq >= 1.0 || rsq == 0.0);
    return v2 * sqrt(-2.0 * log(rsq) / rsq);
  }
};
}  // namespace

double evaluate(const std::string& e) { Parser p(e); return p.expr(); }

std::string LAMMPS::substitute(const std::string& in) const {
  std::string out;
  for (size_t i = 0; i < in.size();) {
    if (in[i] == '$' && i + 1 < in.size()) {
      std::string name;
      size_t j;
      if (in[i + 1] == '{') { j = in.find('}', i); name = in.substr(i + 2, j - i - 2); j++; }
      else { name = in.substr(i + 1, 1); j = i + 2; }
      auto it = vars.find(name);
      if (it == vars.end()) throw std::runtime_error("ERROR: Substitution for illegal variable " + name);
      out += it->second;
      i = j;
    } else out += in[i++];
  }
  return out;
}

void LAMMPS::file(const std::string& path) {
  std::ifstream in(path);
  if (!in.is_open()) error->all(FLERR, "Cannot open input script " + path);
  std::string line, acc;
  while (std::getline(in, line)) {
    size_t h = line.find('#');
    if (h != std::string::npos) line.resize(h);
    size_t e = line.find_last_not_of(" \t\r\n");
    if (e == std::string::npos) {   // blank line: terminates a command whose last line ended in '&' (in.reaxc.lattice:409-410)
      if (!acc.empty()) { one(acc); acc.clear(); }
      continue;
    }
    line.resize(e + 1);
    if (line.back() == '&') { acc += line.substr(0, line.size() - 1) + " "; continue; }
    acc += line;
    one(acc);
    acc.clear();
  }
  if (!acc.empty()) one(acc);
}

void LAMMPS::one(const std::string& raw) {
  std::vector<std::string> w = words(substitute(raw));
  if (w.empty()) return;
  const std::string& c = w[0];
  auto need = [&](size_t n) { if (w.size() < n) error->all(FLERR, "Illegal " + c + " command"); };
  std::vector<char*> argv;
  for (size_t i = 1; i < w.size(); i++) argv.push_back(const_cast<char*>(w[i].c_str()));

  if (c == "variable") {
    need(4);
    if (w[2] == "index") { if (!vars.count(w[1])) vars[w[1]] = w[3]; }        // -var on the command line wins
    else if (w[2] == "equal") {
      std::string e;
      for (size_t i = 3; i < w.size(); i++) e += w[i];
      char buf[64];
      snprintf(buf, sizeof(buf), "%.15g", evaluate(e));
      vars[w[1]] = buf;
    } else error->all(FLERR, "Illegal variable command");
  } else if (c == "units") { need(2); if (w[1] != "real") error->all(FLERR, "only units real are supported by this driver"); }
  else if (c == "atom_style") { need(2); if (w[1] != "charge") error->all(FLERR, "Pair style reax/c requires atom attribute q"); }
  else if (c == "compute") {
    need(4);
    if (w[3] == "SPEC/ATOM" || w[3] == "SPEC/ATOM/b200" || w[3] == "reax/c/atom")
      computes.emplace_back(new ComputeSpecAtomB200(this, (int)argv.size(), argv.data()));
    // (any other compute style of the reference's script - its thermo-only `compute reax all pair reax/c` - is a no-op here)
  }
  else if (c == "boundary" || c == "thermo_style" || c == "dump" || c == "dump_modify" || c == "echo" || c == "log") {}
  else if (c == "lattice") {
    need(3);
    if (w[1] != "custom") error->all(FLERR, "only 'lattice custom' is supported by this driver");
    const double scale = atof(w[2].c_str());
    basis.clear();
    for (size_t i = 3; i < w.size();) {
      if ((w[i] == "a1" || w[i] == "a2" || w[i] == "a3") && i + 3 < w.size() + 0) {
        double* a = w[i] == "a1" ? a1 : (w[i] == "a2" ? a2 : a3);
        for (int t = 0; t < 3; t++) a[t] = scale * atof(w[i + 1 + t].c_str());
        i += 4;
      } else if (w[i] == "basis") { for (int t = 0; t < 3; t++) basis.push_back(atof(w[i + 1 + t].c_str())); i += 4; }
      else error->all(FLERR, "Illegal lattice command");
    }
  } else if (c == "region") {
    need(12);
    if (w[2] != "prism") error->all(FLERR, "only 'region ID prism' is supported by this driver");
    double v[9];
    for (int t = 0; t < 9; t++) v[t] = atof(w[3 + t].c_str());
    domain->boxlo[0] = v[0]; domain->boxlo[1] = v[2]; domain->boxlo[2] = v[4];
    domain->set_box(v[1] - v[0], v[3] - v[2], v[5] - v[4], v[6], v[7], v[8]);
  } else if (c == "create_box") { need(3); atom->ntypes = atoi(w[1].c_str()); atom->mass.assign(atom->ntypes + 1, 0.0); }
  else if (c == "create_atoms") {
    need(3);
    std::vector<int> btype(basis.size() / 3, atoi(w[1].c_str()));
    for (size_t i = 3; i + 2 < w.size() + 0 && i < w.size(); ) {
      if (w[i] == "basis") { btype.at(atoi(w[i + 1].c_str()) - 1) = atoi(w[i + 2].c_str()); i += 3; }
      else error->all(FLERR, "Illegal create_atoms command");
    }
    // lattice points i*a1 + j*a2 + k*a3 + basis inside the (triclinic) box, LAMMPS create_atoms box semantics
    const double* hh = domain->h;
    const int ni = (int)ceil(hh[0] / std::max(fabs(a1[0]), 1e-12)) + 2, nj = (int)ceil(hh[1] / std::max(fabs(a2[1]), 1e-12)) + 2,
              nk = (int)ceil(hh[2] / std::max(fabs(a3[2]), 1e-12)) + 2;
    const double eps = 1e-8;
    for (int k = -1; k < nk; k++)
      for (int j = -1; j < nj; j++)
        for (int i = -1; i < ni; i++)
          for (size_t b = 0; b < basis.size() / 3; b++) {
            const double fi = i + basis[3 * b], fj = j + basis[3 * b + 1], fk = k + basis[3 * b + 2];
            const double x[3] = {fi * a1[0] + fj * a2[0] + fk * a3[0], fi * a1[1] + fj * a2[1] + fk * a3[1], fi * a1[2] + fj * a2[2] + fk * a3[2]};
            double l[3];
            domain->x2lamda(x, l);
            if (l[0] < -eps || l[0] >= 1.0 - eps || l[1] < -eps || l[1] >= 1.0 - eps || l[2] < -eps || l[2] >= 1.0 - eps) continue;
            for (int t = 0; t < 3; t++) atom->x.push_back(x[t]);
            atom->type.push_back(btype[b]);
            atom->tag.push_back((int)atom->tag.size() + 1);
          }
    atom->nlocal = (int)atom->tag.size();
    atom->natoms = atom->nlocal;
    atom->q.assign(atom->nlocal, 0.0);
    atom->v.assign((size_t)3 * atom->nlocal, 0.0);
  } else if (c == "read_data") {
    need(2);
    std::ifstream in(w[1]);
    if (!in.is_open()) error->all(FLERR, "Cannot open file " + w[1]);
    std::string ln;
    double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1}, tilt[3] = {0, 0, 0};
    long natoms = 0;
    std::string section;
    while (std::getline(in, ln)) {
      std::vector<std::string> t = words(ln);
      if (t.empty()) continue;
      if (t.size() >= 2 && t[1] == "atoms") natoms = atol(t[0].c_str());
      else if (t.size() >= 3 && t[1] == "atom" && t[2] == "types") { atom->ntypes = atoi(t[0].c_str()); atom->mass.assign(atom->ntypes + 1, 0.0); }
      else if (t.size() >= 4 && t[2] == "xlo") { lo[0] = atof(t[0].c_str()); hi[0] = atof(t[1].c_str()); }
      else if (t.size() >= 4 && t[2] == "ylo") { lo[1] = atof(t[0].c_str()); hi[1] = atof(t[1].c_str()); }
      else if (t.size() >= 4 && t[2] == "zlo") { lo[2] = atof(t[0].c_str()); hi[2] = atof(t[1].c_str()); }
      else if (t.size() >= 6 && t[3] == "xy") { for (int k = 0; k < 3; k++) tilt[k] = atof(t[k].c_str()); }
      else if (t[0] == "Masses" || t[0] == "Atoms" || t[0] == "Velocities") section = t[0];
      else if (section == "Masses" && t.size() >= 2) atom->mass.at(atoi(t[0].c_str())) = atof(t[1].c_str());
      else if (section == "Atoms" && t.size() >= 6) {   // atom_style charge: id type q x y z
        atom->tag.push_back(atoi(t[0].c_str())); atom->type.push_back(atoi(t[1].c_str())); atom->q.push_back(atof(t[2].c_str()));
        for (int k = 0; k < 3; k++) atom->x.push_back(atof(t[3 + k].c_str()));
      }
    }
    for (int k = 0; k < 3; k++) domain->boxlo[k] = lo[k];
    domain->set_box(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], tilt[0], tilt[1], tilt[2]);
    atom->nlocal = (int)atom->tag.size();
    if (natoms != atom->nlocal) error->all(FLERR, "Did not assign all atoms correctly");
    atom->natoms = natoms;
    atom->v.assign((size_t)3 * atom->nlocal, 0.0);
  } else if (c == "replicate") {
    need(4);
    const int nx = atoi(w[1].c_str()), ny = atoi(w[2].c_str()), nz = atoi(w[3].c_str());
    const int n0 = atom->nlocal;
    const double* hh = domain->h;
    std::vector<double> x0 = atom->x, q0 = atom->q;
    std::vector<int> t0 = atom->type;
    atom->x.clear(); atom->q.clear(); atom->type.clear(); atom->tag.clear();
    for (int iz = 0; iz < nz; iz++)
      for (int iy = 0; iy < ny; iy++)
        for (int ix = 0; ix < nx; ix++)
          for (int i = 0; i < n0; i++) {
            atom->x.push_back(x0[3 * i] + ix * hh[0] + iy * hh[5] + iz * hh[4]);
            atom->x.push_back(x0[3 * i + 1] + iy * hh[1] + iz * hh[3]);
            atom->x.push_back(x0[3 * i + 2] + iz * hh[2]);
            atom->q.push_back(q0[i]); atom->type.push_back(t0[i]); atom->tag.push_back((int)atom->tag.size() + 1);
          }
    domain->set_box(hh[0] * nx, hh[1] * ny, hh[2] * nz, hh[5] * ny, hh[4] * nz, hh[3] * nz);
    atom->nlocal = (int)atom->tag.size();
    atom->natoms = atom->nlocal;
    atom->v.assign((size_t)3 * atom->nlocal, 0.0);
  } else if (c == "mass") { need(3); atom->mass.at(atoi(w[1].c_str())) = atof(w[2].c_str()); }
  else if (c == "pair_style") {
    need(2);
    if (w[1] != "reax/c" && w[1] != "reax/c/b200") error->all(FLERR, "Unknown pair style " + w[1]);
    pair.reset(new PairReaxCB200(this));
    pair->settings((int)argv.size() - 1, argv.data() + 1);
  } else if (c == "pair_coeff") {
    if (!pair) error->all(FLERR, "Pair_coeff command before pair_style is defined");
    pair->coeff((int)argv.size(), argv.data());
  } else if (c == "neighbor") { need(3); neighbor->skin = atof(w[1].c_str()); }
  else if (c == "neigh_modify") {
    for (size_t i = 1; i + 1 < w.size(); i += 2) {
      if (w[i] == "every") neighbor->every = atoi(w[i + 1].c_str());
      else if (w[i] == "delay") neighbor->delay = atoi(w[i + 1].c_str());
      else if (w[i] == "check") neighbor->dist_check = w[i + 1] == "yes";
      else if (w[i] == "one" || w[i] == "page") {}
      else error->all(FLERR, "Illegal neigh_modify command");
    }
    if (neighbor->dist_check) error->warning(FLERR, "neigh_modify check yes is treated as check no by this driver");
  } else if (c == "fix") {
    need(4);
    if (w[3] == "nve" || w[3] == "nve/b200") fixes.emplace_back(new FixNVEB200(this, (int)argv.size(), argv.data()));
    else if (w[3] == "qeq/reax" || w[3] == "qeq/reax/b200") fixes.emplace_back(new FixQEqReaxB200(this, (int)argv.size(), argv.data()));
    else if (w[3] == "reax/c/bonds" || w[3] == "reax/c/bonds/b200") fixes.emplace_back(new FixReaxCBondsB200(this, (int)argv.size(), argv.data()));
    else if (w[3] == "reax/c/species" || w[3] == "reax/c/species/b200") fixes.emplace_back(new FixReaxCSpeciesB200(this, (int)argv.size(), argv.data()));
    else error->all(FLERR, "Unknown fix style " + w[3]);
  } else if (c == "velocity") {
    need(5);
    if (w[2] != "create") error->all(FLERR, "only 'velocity all create T seed' is supported by this driver");
    const double T = atof(w[3].c_str());
    Rng rng(atol(w[4].c_str()));
    double p[3] = {0, 0, 0}, mtot = 0;
    for (int i = 0; i < atom->nlocal; i++) {
      const double m = atom->mass[atom->type[i]];
      for (int t = 0; t < 3; t++) { atom->v[3 * i + t] = rng.gaussian() / sqrt(m); p[t] += m * atom->v[3 * i + t]; }
      mtot += m;
    }
    for (int i = 0; i < atom->nlocal; i++) for (int t = 0; t < 3; t++) atom->v[3 * i + t] -= p[t] / mtot;
    const double tcur = 2 * kinetic() / (3.0 * (atom->nlocal - 1) * force->boltz);
    const double s = tcur > 0 ? sqrt(T / tcur) : 0.0;
    for (double& vv : atom->v) vv *= s;
  } else if (c == "thermo") { need(2); thermo_every = atoi(w[1].c_str()); }
  else if (c == "timestep") { need(2); update->dt = atof(w[1].c_str()); }
  else if (c == "run") { need(2); run(atol(w[1].c_str())); }
  else error->all(FLERR, "Unknown command: " + c);
}

double LAMMPS::kinetic() const {
  double ke = 0;
  for (int i = 0; i < atom->nlocal; i++) {
    const double* v = &atom->v[3 * i];
    ke += atom->mass[atom->type[i]] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  }
  return 0.5 * force->mvv2e * ke;
}

void LAMMPS::thermo_line(int) {
  Thermo t;
  t.step = update->ntimestep;
  t.ke = kinetic();
  t.pe = pair->eng_vdwl + pair->eng_coul;
  t.etotal = t.pe + t.ke;
  t.temp = atom->nlocal > 1 ? 2 * t.ke / (3.0 * (atom->nlocal - 1) * force->boltz) : 0.0;
  memcpy(t.pvector, pair->pvector, sizeof(t.pvector));
  thermo_log.push_back(t);
  if (echo_thermo) printf("%8ld %14.6f %18.6f %18.6f\n", t.step, t.temp, t.pe, t.etotal);
}

void LAMMPS::setup() {
  if (!pair) error->all(FLERR, "No pair style defined");
  for (auto& f : fixes) f->init();
  pair->init_style();
  domain->remap(*atom);
  comm->borders(*domain, *atom, pair->cutghost_request() + neighbor->skin);
  neighbor->ago = 0;
  neighbor->ncalls++;
  const int ev = 1;
  for (auto& f : fixes) f->setup_pre_force(ev);
  std::fill(atom->f.begin(), atom->f.end(), 0.0);
  pair->compute(ev, ev);
  comm->reverse_comm(*atom);
  for (auto& f : fixes) f->setup(ev);
  for (auto& c : computes) { c->init(); c->compute_peratom(); }
  if (echo_thermo) printf("    Step           Temp             PotEng             TotEng\n");
  thermo_line(ev);
  setup_done_ = true;
}

void LAMMPS::run(long nsteps) {
  setup();
  iterate(nsteps);
}

void LAMMPS::iterate(long nsteps) {
  using clk = std::chrono::steady_clock;
  auto lap = [&](int k, clk::time_point& t0) { auto t1 = clk::now(); phase_s[k] += std::chrono::duration<double>(t1 - t0).count(); t0 = t1; };
  for (long s = 0; s < nsteps; s++) {
    auto t0 = clk::now();
    update->ntimestep++;
    const int ev = (thermo_every && update->ntimestep % thermo_every == 0) || s == nsteps - 1;
    for (auto& f : fixes) f->initial_integrate(ev);
    for (auto& f : fixes) f->post_integrate();
    lap(0, t0);
    if (neighbor->decide()) {
      domain->remap(*atom);
      comm->borders(*domain, *atom, pair->cutghost_request() + neighbor->skin);
      neighbor->ago = 0;
      neighbor->ncalls++;
      lap(1, t0);
    } else {
      comm->forward_comm(*atom);
      lap(2, t0);
    }
    {
      double* f = atom->f.data();
      const long n3 = (long)atom->f.size();
#pragma omp parallel for schedule(static)
      for (long k = 0; k < n3; k++) f[k] = 0.0;
    }
    lap(3, t0);
    for (auto& f : fixes) f->pre_force(ev);
    lap(4, t0);
    pair->compute(ev, ev);
    lap(5, t0);
    comm->reverse_comm(*atom);
    lap(6, t0);
    for (auto& f : fixes) f->final_integrate();
    for (auto& f : fixes) if (f->nevery > 0 && update->ntimestep % f->nevery == 0) f->end_of_step();
    if (ev) { for (auto& c : computes) c->compute_peratom(); thermo_line(ev); }
    lap(7, t0);
  }
}

}  // namespace LAMMPS_MINI
