// ORACLE/_REF — TEST INFRASTRUCTURE ONLY.
// Harness (written for this repo) that runs the reference's own fix qeq/reax — fix_qeq_reax_sunway.cpp, compiled
// UNMODIFIED from /root/reference against the LAMMPS-core stand-in in stubs/lammps_stub.h — on atoms, ghosts and a full
// neighbour list supplied by the caller.  Executed reference code: the constructor's argument handling,
// pertype_parameters (parameter-file form), init_shielding, init_taper, allocate/init_storage, init_matvec (history
// extrapolation), CG_v2 (the pipelined PCG, its stopping rule and iteration count), calculate_Q (charges + history shift),
// pack/unpack_forward_comm, calculate_H.
// The four Sunway slave-core entry points the file calls (compute_H_Full_C, sparse_matvec_C, sparse_matvec_C_spawn/join;
// athread kernels in fix_qeq_reax_sw64.c) are provided here as plain serial loops over the same param pack, following
// the serial forms the reference keeps in comments (fix_qeq_reax_sw64.c:147-189; fix_qeq_reax_sunway.cpp:1601-1622);
// every H entry is evaluated with the reference's own FixQEqReaxSunway::calculate_H.
#include <math.h>

#include <vector>

#include "fix_qeq_reax_sunway.h"
#include "pair_reaxc_sunway.h"
#include "reaxc_ctypes_sunway.h"

using namespace LAMMPS_NS;
using namespace REAXC_SUNWAY_NS;

namespace {
struct FixOpen : public FixQEqReaxSunway {   // opens the protected members to the harness
  FixOpen(LAMMPS* l, int narg, char** arg) : FixQEqReaxSunway(l, narg, arg) {}
  double H_of(double r, double g) { return calculate_H(r, g); }
  double** s_hist_() { return s_hist; }
  double** t_hist_() { return t_hist; }
  double* s_() { return s; }
  double* t_() { return t; }
  double* tap_() { return Tap; }
  int ms() const { return matvecs_s; }
  int mt() const { return matvecs_t; }
  int m_fill_() const { return m_fill; }
  int* H_num() { return H.numnbrs; }
  int* H_first() { return H.firstnbr; }
  int* H_j() { return H.jlist; }
  double* H_v() { return H.val; }
};
FixOpen* g_fix = nullptr;
}  // namespace

extern "C" {

// H rows for the local atoms: firstnbr[i] = maxHlist * i, every list neighbour with r^2 <= swb^2 (no tag filter)
void compute_H_Full_C(void* vp) {
  fix_qeq_pack_t* p = (fix_qeq_pack_t*)vp;
  const int nlocal = p->atom.nlocal;
  int m_fill = 0;
  for (int i = 0; i < nlocal; i++) {
    p->H.numnbrs[i] = 0;
    p->H.firstnbr[i] = p->maxHlist * i;
    if (!(p->atom.mask[i] & p->groupbit)) continue;
    int foff = 0;
    const int* jlist = p->list.firstneigh[i];
    for (int jj = 0; jj < p->list.numneigh[i]; jj++) {
      const int j = jlist[jj];
      const double dx = p->atom.x[j][0] - p->atom.x[i][0], dy = p->atom.x[j][1] - p->atom.x[i][1],
                   dz = p->atom.x[j][2] - p->atom.x[i][2];
      const double r_sqr = dx * dx + dy * dy + dz * dz;
      if (r_sqr <= p->swbsq) {
        p->H.jlist[p->H.firstnbr[i] + foff] = j;
        p->H.val[p->H.firstnbr[i] + foff] = g_fix->H_of(sqrt(r_sqr), p->shld[p->atom.type[i]][p->atom.type[j]]);
        foff++;
      }
    }
    p->H.numnbrs[i] = foff;
    m_fill += foff;
  }
  p->m_fill = m_fill;
}

// b = (H + diag(eta)) x over the local rows
void sparse_matvec_C(void* vp) {
  fix_qeq_pack_t* p = (fix_qeq_pack_t*)vp;
  for (int ii = 0; ii < p->list.inum; ii++) {
    const int i = p->list.ilist[ii];
    if (!(p->atom.mask[i] & p->groupbit)) continue;
    double acc = p->eta[p->atom.type[i]] * p->x[i];
    for (int k = p->H.firstnbr[i]; k < p->H.firstnbr[i] + p->H.numnbrs[i]; k++) acc += p->H.val[k] * p->x[p->H.jlist[k]];
    p->b[i] = acc;
  }
}
void sparse_matvec_C_spawn(void* vp) { sparse_matvec_C(vp); }
void sparse_matvec_C_join() {}

// One fix qeq/reax pre_force on (nlocal + nghost) atoms.  types/mask/tag per atom (LAMMPS types, 1-based), full neighbour
// list rows for the local atoms (CSR), ghost_owner per ghost, per-type chi/eta/gamma (index 1..ntypes), histories
// [nlocal][5] in and out.  Returns 0; out_q[nall], matvecs2 = CG_v2 return values of the s and t solves.
int ref_qeq_pre_force(int nlocal, int nghost, int ntypes, const double* x, const int* type, const int* tag, const long* nb_off,
                      const int* nb, const int* ghost_owner, const double* chi, const double* eta, const double* gamma,
                      double swa, double swb, double tol, double* s_hist, double* t_hist, double* out_q, double* out_s,
                      double* out_t, int* matvecs2, double* tap8, long* H_count, double* H_rowsum) {
  const int nall = nlocal + nghost;
  Atom atom; Comm comm; Memory memory; Error error; Force force; Neighbor neighbor; Update update; Group group; CiteMe cite;
  Domain domain; Modify modify;
  LAMMPS lmp;
  lmp.atom = &atom; lmp.comm = &comm; lmp.memory = &memory; lmp.error = &error; lmp.force = &force; lmp.neighbor = &neighbor;
  lmp.update = &update; lmp.group = &group; lmp.citeme = &cite; lmp.domain = &domain; lmp.modify = &modify;
  std::vector<double> xs(x, x + (size_t)3 * nall), q(nall, 0.0);
  std::vector<double*> xrow(nall);
  for (int i = 0; i < nall; i++) xrow[i] = &xs[(size_t)3 * i];
  std::vector<int> ty(type, type + nall), tg(tag, tag + nall), mask(nall, 1);
  atom.x = xrow.data(); atom.type = ty.data(); atom.tag = tg.data(); atom.mask = mask.data(); atom.q = q.data();
  atom.nlocal = nlocal; atom.nghost = nghost; atom.nmax = nall + 64; atom.ntypes = ntypes;
  comm.atom = &atom;
  comm.ghost_owner.assign(ghost_owner, ghost_owner + nghost);
  group.natoms = nlocal;
  // the patched core's full list: rows for local atoms, inum = nlocal, gnum = nghost
  NeighList list;
  std::vector<int> ilist(nall), numneigh(nall, 0);
  std::vector<int*> first(nall, nullptr);
  std::vector<int> cols(nb, nb + nb_off[nlocal]);
  for (int i = 0; i < nall; i++) ilist[i] = i;
  for (int i = 0; i < nlocal; i++) { numneigh[i] = (int)(nb_off[i + 1] - nb_off[i]); first[i] = cols.data() + nb_off[i]; }
  list.inum = nlocal; list.gnum = nghost; list.ilist = ilist.data(); list.numneigh = numneigh.data(); list.firstneigh = first.data();
  // a PairReaxCSunway-shaped object: only its data members listfull / system are ever read by the fix on this path
  void* raw = calloc(1, sizeof(PairReaxCSunway) + 64);
  Pair* pair = reinterpret_cast<Pair*>(raw);
  pair->listfull = &list;
  force.pair = pair;
  // per-type parameters through the fix's parameter-file form (no virtual call into the pair style needed)
  char fname[] = "/tmp/ref_qeq_param_XXXXXX";
  int fd = mkstemp(fname);
  if (fd < 0) return -1;
  FILE* pf = fdopen(fd, "w");
  for (int k = 1; k <= ntypes; k++) fprintf(pf, "%d %.17g %.17g %.17g\n", k, chi[k], eta[k], gamma[k]);
  fclose(pf);
  char a0[] = "qeq", a1[] = "all", a2[] = "qeq/reax", a3[] = "1", a4[32], a5[32], a6[32];
  snprintf(a4, sizeof a4, "%.17g", swa); snprintf(a5, sizeof a5, "%.17g", swb); snprintf(a6, sizeof a6, "%.17g", tol);
  char* args[8] = {a0, a1, a2, a3, a4, a5, a6, fname};
  int rc = 0;
  try {
    FixOpen fix(&lmp, 8, args);
    g_fix = &fix;
    fix.post_constructor();
    fix.init();
    for (int i = 0; i < nlocal; i++)
      for (int k = 0; k < 5; k++) { fix.s_hist_()[i][k] = s_hist[5 * i + k]; fix.t_hist_()[i][k] = t_hist[5 * i + k]; }
    fix.setup_pre_force(0);          // allocate_storage, init_storage, allocate_matrix, pre_force
    for (int i = 0; i < nall; i++) { out_q[i] = q[i]; }
    for (int i = 0; i < nlocal; i++) {
      out_s[i] = fix.s_()[i]; out_t[i] = fix.t_()[i];
      for (int k = 0; k < 5; k++) { s_hist[5 * i + k] = fix.s_hist_()[i][k]; t_hist[5 * i + k] = fix.t_hist_()[i][k]; }
      if (H_rowsum) {
        double a = 0;
        for (int k = fix.H_first()[i]; k < fix.H_first()[i] + fix.H_num()[i]; k++) a += fix.H_v()[k];
        H_rowsum[i] = a;
      }
    }
    matvecs2[0] = fix.ms(); matvecs2[1] = fix.mt();
    for (int k = 0; k < 8; k++) tap8[k] = fix.tap_()[k];
    if (H_count) { long c = 0; for (int i = 0; i < nlocal; i++) c += fix.H_num()[i]; *H_count = c; }
    g_fix = nullptr;
  } catch (const std::exception& e) {
    fprintf(stderr, "oracle/_ref qeq: %s\n", e.what());
    rc = -2;
  }
  remove(fname);
  free(raw);
  return rc;
}

}  // extern "C"
