// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under sw_reaxff_b200/ may include, link or call this.
// Plain C entry points for ctypes (tests/, __graft_entry__.smoke(), bench.py cpu_baseline only).
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "orc_analysis.h"
#include "orc_md.h"

using namespace orc;

struct OrcHandle {
  MD md;
  SpeciesFix species;
  std::string err;
  std::string text;
};

static void put_err(char* err, int errlen, const std::string& s) {
  if (err && errlen > 0) { strncpy(err, s.c_str(), errlen - 1); err[errlen - 1] = 0; }
}

extern "C" {

void* orc_create(const char* ffield, const char* control, int ntypes, const char** elements, int lgflag,
                 int enobondsflag, char* err, int errlen) {
  OrcHandle* h = new OrcHandle();
  Params& P = h->md.sys.prm;
  P.lgflag = lgflag;
  P.enobondsflag = enobondsflag;
  std::string e = read_control(control, P);
  if (e.empty()) e = read_force_field(ffield, P);
  if (e.empty()) e = set_element_map(P, ntypes, elements);
  if (!e.empty()) { put_err(err, errlen, e); delete h; return nullptr; }
  h->md.qeq.init(P, 0.0, 10.0, 1e-6);
  return h;
}
void orc_destroy(void* hh) { delete (OrcHandle*)hh; }

// flatten every parameter table in a fixed order (must match rxb_params_dump in the product, which is
// written independently): returns number of doubles written (or needed if out==NULL)
long orc_params_dump(void* hh, double* out, long cap) {
  Params& P = ((OrcHandle*)hh)->md.sys.prm;
  std::vector<double> v;
  v.push_back(P.nt); v.push_back(P.vdw_type); v.push_back((double)P.gp.size());
  for (double g : P.gp) v.push_back(g);
  v.push_back(P.bo_cut); v.push_back(P.nonb_low); v.push_back(P.nonb_cut); v.push_back(P.bond_cut);
  v.push_back(P.hbond_cut); v.push_back(P.bg_cut); v.push_back(P.thb_cut); v.push_back(P.thb_cutsq);
  v.push_back(P.tabulate); v.push_back(P.energy_update_freq);
  for (int i = 0; i < 8; i++) v.push_back(P.Tap[i]);
  for (const Sbp& s : P.sbp) {
    double a[] = {s.r_s, s.valency, s.mass, s.r_vdw, s.epsilon, s.gamma, s.r_pi, s.valency_e, s.nlp_opt, s.alpha,
                  s.gamma_w, s.valency_boc, s.p_ovun5, s.chi, s.eta, (double)s.p_hbond, s.r_pi_pi, s.p_lp2, s.b_o_131,
                  s.b_o_132, s.b_o_133, s.p_ovun2, s.p_val3, s.valency_val, s.p_val5, s.rcore2, s.ecore2, s.acore2,
                  s.lgcij, s.lgre};
    v.insert(v.end(), a, a + sizeof(a) / sizeof(double));
  }
  for (const Tbp& t : P.tbp) {
    double a[] = {t.p_bo1, t.p_bo2, t.p_bo3, t.p_bo4, t.p_bo5, t.p_bo6, t.r_s, t.r_p, t.r_pp, t.p_boc3, t.p_boc4,
                  t.p_boc5, t.p_be1, t.p_be2, t.De_s, t.De_p, t.De_pp, t.p_ovun1, t.D, t.alpha, t.r_vdW, t.gamma_w,
                  t.rcore, t.ecore, t.acore, t.lgcij, t.lgre, t.gamma, t.v13cor, t.ovc};
    v.insert(v.end(), a, a + sizeof(a) / sizeof(double));
  }
  for (const ThbHeader& t : P.thbp) {
    v.push_back(t.cnt);
    for (int c = 0; c < 5; c++) {
      const Thbp& q = t.prm[c];
      double a[] = {q.theta_00, q.p_val1, q.p_val2, q.p_coa1, q.p_val7, q.p_pen1, q.p_val4};
      v.insert(v.end(), a, a + 7);
    }
  }
  for (const FbHeader& t : P.fbp) {
    v.push_back(t.cnt);
    const Fbp& q = t.prm[0];
    double a[] = {q.V1, q.V2, q.V3, q.p_tor1, q.p_cot1};
    v.insert(v.end(), a, a + 5);
  }
  for (const Hbp& t : P.hbp) { v.push_back(t.r0_hb); v.push_back(t.p_hb1); v.push_back(t.p_hb2); v.push_back(t.p_hb3); }
  for (size_t i = 1; i < P.map.size(); i++) v.push_back(P.map[i]);
  if (out) for (long i = 0; i < (long)v.size() && i < cap; i++) out[i] = v[i];
  return (long)v.size();
}

// ---- static configuration (the index space pair reax/c sees) ----
void orc_set_atoms(void* hh, int n, int N, const double* x, const int* ltype, const int* tag, const double* q) {
  System& s = ((OrcHandle*)hh)->md.sys;
  s.n = n; s.N = N;
  s.x.assign(x, x + (size_t)3 * N);
  s.type.resize(N);
  for (int i = 0; i < N; i++) s.type[i] = s.prm.map[ltype[i]];
  s.tag.assign(tag, tag + N);
  s.q.assign(q, q + N);
}
void orc_build_neighbors(void* hh, double cutneigh) { build_full_neighbor_list(((OrcHandle*)hh)->md.sys, cutneigh); }
long orc_num_neighbors(void* hh) { return (long)((OrcHandle*)hh)->md.sys.nb.size(); }
void orc_get_neighbors(void* hh, long* off, int* nb) {
  System& s = ((OrcHandle*)hh)->md.sys;
  for (int i = 0; i <= s.N; i++) off[i] = s.nb_off[i];
  memcpy(nb, s.nb.data(), s.nb.size() * sizeof(int));
}
void orc_set_neighbors(void* hh, const long* off, const int* nb) {
  System& s = ((OrcHandle*)hh)->md.sys;
  s.nb_off.assign(off, off + s.N + 1);
  s.nb.assign(nb, nb + off[s.N]);
}
void orc_compute(void* hh) { compute_forces(((OrcHandle*)hh)->md.sys); }
// individual phases (for per-kernel parity tests); caller is responsible for the order
void orc_phase(void* hh, int which) {
  System& s = ((OrcHandle*)hh)->md.sys;
  switch (which) {
    case 0: s.en = Energies(); for (int t = 0; t < 6; t++) s.virial[t] = 0; s.fCd.assign((size_t)s.N * 4, 0.0); break;
    case 1: build_bond_list(s); break;
    case 2: build_hbond_list(s); break;
    case 3: nonbonded(s); break;
    case 4: bond_orders(s); break;
    case 5: bonds_atom_energy(s); break;
    case 6: hydrogen_bonds(s); break;
    case 7: valence_torsion(s); break;
    case 8: add_dbond_forces(s); break;
  }
}
void orc_get_forces(void* hh, double* f) {  // f = -fCd xyz, all N
  System& s = ((OrcHandle*)hh)->md.sys;
  for (int i = 0; i < s.N; i++) for (int t = 0; t < 3; t++) f[3 * i + t] = -s.fCd[4 * i + t];
}
void orc_get_cddelta(void* hh, double* c) {
  System& s = ((OrcHandle*)hh)->md.sys;
  for (int i = 0; i < s.N; i++) c[i] = s.fCd[4 * i + 3];
}
void orc_get_energies(void* hh, double* e13, double* virial6) {
  System& s = ((OrcHandle*)hh)->md.sys;
  const Energies& e = s.en;
  double a[13] = {e.e_bond, e.e_ov, e.e_un, e.e_lp, e.e_ang, e.e_pen, e.e_coa, e.e_hb, e.e_tor, e.e_con, e.e_vdW, e.e_ele, e.e_pol};
  memcpy(e13, a, sizeof(a));
  if (virial6) memcpy(virial6, s.virial, 6 * sizeof(double));
}
int orc_num_bonds(void* hh) { return (int)((OrcHandle*)hh)->md.sys.bonds.size(); }
// per directed bond, CSR by atom: fields = nbr,sym | d,dvec3,BO,BO_s,BO_pi,BO_pi2,dBOp3,dlnpi3,dlnpi2_3,C1..3dbo,C1..4dbopi,C1..4dbopi2,Cdbo,Cdbopi,Cdbopi2 (31 doubles)
void orc_get_bonds(void* hh, int* b_start, int* b_end, int* nbr, int* sym, double* fields31) {
  System& s = ((OrcHandle*)hh)->md.sys;
  for (int i = 0; i < s.N; i++) { b_start[i] = s.b_start[i]; b_end[i] = s.b_end[i]; }
  for (size_t p = 0; p < s.bonds.size(); p++) {
    const Bond& b = s.bonds[p];
    nbr[p] = b.nbr; sym[p] = b.sym;
    double* o = fields31 + 31 * p;
    o[0] = b.d; o[1] = b.dvec[0]; o[2] = b.dvec[1]; o[3] = b.dvec[2];
    o[4] = b.BO; o[5] = b.BO_s; o[6] = b.BO_pi; o[7] = b.BO_pi2;
    for (int t = 0; t < 3; t++) { o[8 + t] = b.dBOp[t]; o[11 + t] = b.dln_BOp_pi[t]; o[14 + t] = b.dln_BOp_pi2[t]; }
    o[17] = b.C1dbo; o[18] = b.C2dbo; o[19] = b.C3dbo;
    o[20] = b.C1dbopi; o[21] = b.C2dbopi; o[22] = b.C3dbopi; o[23] = b.C4dbopi;
    o[24] = b.C1dbopi2; o[25] = b.C2dbopi2; o[26] = b.C3dbopi2; o[27] = b.C4dbopi2;
    o[28] = b.Cdbo; o[29] = b.Cdbopi; o[30] = b.Cdbopi2;
  }
}
// per-atom workspace: 16 columns
void orc_get_workspace(void* hh, double* w16) {
  System& s = ((OrcHandle*)hh)->md.sys;
  for (int i = 0; i < s.N; i++) {
    double* o = w16 + 16 * i;
    o[0] = s.total_bo[i]; o[1] = s.Delta_boc[i]; o[2] = s.Deltap[i]; o[3] = s.Deltap_boc[i]; o[4] = s.Delta[i];
    o[5] = s.Delta_e[i]; o[6] = s.Delta_val[i]; o[7] = s.vlpex[i]; o[8] = s.nlp[i]; o[9] = s.Delta_lp[i];
    o[10] = s.Clp[i]; o[11] = s.dDelta_lp[i]; o[12] = s.nlp_temp[i]; o[13] = s.Delta_lp_temp[i];
    o[14] = s.dDelta_lp_temp[i]; o[15] = 0;
  }
}
// spline tables of the tabulated mode: returns n (= tabulate + 2); out = [nt*nt][5][n][4] (a,b,c,d)
int orc_lookup_tables(void* hh, double* out) {
  Params& P = ((OrcHandle*)hh)->md.sys.prm;
  if (P.tabulate <= 0) return 0;
  if (P.lookup.n != P.tabulate + 2) build_lookup_tables(P);
  if (out) memcpy(out, P.lookup.tables.data(), P.lookup.tables.size() * sizeof(SplineCoef));
  return P.lookup.n;
}
void orc_get_ddeltap_self(void* hh, double* d3) {
  System& s = ((OrcHandle*)hh)->md.sys;
  memcpy(d3, s.dDeltap_self.data(), (size_t)3 * s.N * sizeof(double));
}
int orc_num_hbonds(void* hh) { return (int)((OrcHandle*)hh)->md.sys.hbonds.size(); }
void orc_get_hbonds(void* hh, int* Hindex, int* hb_start, int* hb_end, int* nbr) {
  System& s = ((OrcHandle*)hh)->md.sys;
  for (int i = 0; i < s.N; i++) Hindex[i] = s.Hindex[i];
  for (size_t h = 0; h < s.hb_start.size(); h++) { hb_start[h] = s.hb_start[h]; hb_end[h] = s.hb_end[h]; }
  for (size_t p = 0; p < s.hbonds.size(); p++) nbr[p] = s.hbonds[p].nbr;
}

// ---- QEq ----
void orc_qeq_init(void* hh, double swa, double swb, double tol) {
  OrcHandle* h = (OrcHandle*)hh;
  h->md.qeq.init(h->md.sys.prm, swa, swb, tol);
}
void orc_qeq_set_hist(void* hh, const double* s_hist, const double* t_hist) {
  OrcHandle* h = (OrcHandle*)hh;
  int n = h->md.sys.n;
  h->md.qeq.s_hist.assign(s_hist, s_hist + (size_t)5 * n);
  h->md.qeq.t_hist.assign(t_hist, t_hist + (size_t)5 * n);
}
void orc_qeq_get_hist(void* hh, double* s_hist, double* t_hist) {
  OrcHandle* h = (OrcHandle*)hh;
  memcpy(s_hist, h->md.qeq.s_hist.data(), h->md.qeq.s_hist.size() * sizeof(double));
  memcpy(t_hist, h->md.qeq.t_hist.data(), h->md.qeq.t_hist.size() * sizeof(double));
}
void orc_qeq_pre_force(void* hh, const int* ghost_owner, int* matvecs2) {
  OrcHandle* h = (OrcHandle*)hh;
  System& s = h->md.sys;
  std::vector<int> go(ghost_owner, ghost_owner + (s.N - s.n));
  h->md.qeq.pre_force(s, go);
  if (matvecs2) { matvecs2[0] = h->md.qeq.matvecs_s; matvecs2[1] = h->md.qeq.matvecs_t; }
}
void orc_get_q(void* hh, double* q) { System& s = ((OrcHandle*)hh)->md.sys; memcpy(q, s.q.data(), s.N * sizeof(double)); }
void orc_qeq_get_st(void* hh, double* sv, double* tv) {
  OrcHandle* h = (OrcHandle*)hh;
  memcpy(sv, h->md.qeq.sv.data(), h->md.sys.N * sizeof(double));
  memcpy(tv, h->md.qeq.tv.data(), h->md.sys.N * sizeof(double));
}
long orc_qeq_get_H(void* hh, long* off, int* num, int* col, double* val) {
  QEq& q = ((OrcHandle*)hh)->md.qeq;
  if (off) {
    memcpy(off, q.H_off.data(), q.H_off.size() * sizeof(long));
    memcpy(num, q.H_num.data(), q.H_num.size() * sizeof(int));
    memcpy(col, q.H_j.data(), q.H_j.size() * sizeof(int));
    memcpy(val, q.H_val.data(), q.H_val.size() * sizeof(double));
  }
  return (long)q.H_j.size();
}

// ---- mini MD (LAMMPS-core stand-in) ----
// box6 = xprd,yprd,zprd,xy,xz,yz ; mass indexed by LAMMPS type (1-based, mass[0] unused)
void orc_md_init(void* hh, const double* box6, int nlocal, const double* x, const double* v, const int* ltype,
                 const int* tag, const double* mass, int ntypes, double dt, double skin, int every, int qeq_on,
                 double qeq_tol) {
  OrcHandle* h = (OrcHandle*)hh;
  MD& md = h->md;
  md.box.set(box6[0], box6[1], box6[2], box6[3], box6[4], box6[5]);
  md.nlocal = nlocal;
  md.sys.x.assign(x, x + (size_t)3 * nlocal);
  md.v.assign(v, v + (size_t)3 * nlocal);
  md.ltype.assign(ltype, ltype + nlocal);
  md.sys.type.resize(nlocal);
  for (int i = 0; i < nlocal; i++) md.sys.type[i] = md.sys.prm.map[ltype[i]];
  md.sys.tag.assign(tag, tag + nlocal);
  md.sys.q.assign(nlocal, 0.0);
  md.mass.assign(mass, mass + ntypes + 1);
  md.dt = dt; md.skin = skin; md.every = every; md.qeq_on = qeq_on != 0;
  const Params& P = md.sys.prm;
  double cutmax = std::max(P.nonb_cut, std::max(P.hbond_cut, 2 * P.bond_cut));  // pair_reaxc_sunway.cpp:410
  md.cutneigh = cutmax + skin;
  md.qeq.init(P, 0.0, 10.0, qeq_tol);
  md.qeq.s_hist.clear(); md.qeq.t_hist.clear();
  md.ntimestep = 0;
  md.setup();
}
void orc_md_run(void* hh, int nsteps) { ((OrcHandle*)hh)->md.run(nsteps); }
int orc_md_nall(void* hh) { return ((OrcHandle*)hh)->md.sys.N; }
void orc_md_get(void* hh, double* x, double* v, double* f, double* q, double* e13, double* pe_ke) {
  MD& md = ((OrcHandle*)hh)->md;
  int n = md.nlocal;
  if (x) memcpy(x, md.sys.x.data(), (size_t)3 * n * sizeof(double));
  if (v) memcpy(v, md.v.data(), (size_t)3 * n * sizeof(double));
  if (f) memcpy(f, md.f.data(), (size_t)3 * n * sizeof(double));
  if (q) memcpy(q, md.sys.q.data(), (size_t)n * sizeof(double));
  if (e13) orc_get_energies(hh, e13, nullptr);
  if (pe_ke) { pe_ke[0] = md.potential(); pe_ke[1] = md.kinetic(); }
}
void orc_md_get_ghosts(void* hh, double* xall, int* typeall, int* tagall, int* owner) {
  MD& md = ((OrcHandle*)hh)->md;
  System& s = md.sys;
  memcpy(xall, s.x.data(), (size_t)3 * s.N * sizeof(double));
  for (int i = 0; i < s.N; i++) { typeall[i] = i < md.nlocal ? md.ltype[i] : md.ltype[md.ghost_owner[i - md.nlocal]]; tagall[i] = s.tag[i]; }
  for (int g = 0; g < s.N - md.nlocal; g++) owner[g] = md.ghost_owner[g];
}
// ---- fix reax/c/bonds, fix reax/c/species on the MD state ----
static long copy_text(const std::string& t, char* out, long cap) {
  if (out && cap > 0) { long m = std::min((long)t.size(), cap - 1); memcpy(out, t.data(), m); out[m] = 0; }
  return (long)t.size();
}
long orc_md_bonds_text(void* hh, long ntimestep, char* out, long cap) {
  OrcHandle* h = (OrcHandle*)hh;
  h->text = bonds_text(h->md, ntimestep);
  return copy_text(h->text, out, cap);
}
void orc_md_species_init(void* hh, int nevery, int nrepeat, int nfreq, const double* bocut) {
  OrcHandle* h = (OrcHandle*)hh;
  int nt = (int)h->md.mass.size() - 1;
  h->species.init(h->md, nevery, nrepeat, nfreq, std::vector<double>(bocut, bocut + (size_t)(nt + 1) * (nt + 1)));
}
// post_integrate hook of timestep `step`; returns 1 when molecules were found (an output step), -1 on error
int orc_md_species_step(void* hh, long step) {
  OrcHandle* h = (OrcHandle*)hh;
  bool f = h->species.post_integrate(h->md, step);
  if (!h->species.error.empty()) return -1;
  return f ? 1 : 0;
}
int orc_md_species_nmole(void* hh) { return ((OrcHandle*)hh)->species.Nmole; }
void orc_md_species_get(void* hh, int* composition, int* cluster_of_local) {
  OrcHandle* h = (OrcHandle*)hh;
  const SpeciesFix& S = h->species;
  if (composition) memcpy(composition, S.composition.data(), S.composition.size() * sizeof(int));
  if (cluster_of_local) for (int i = 0; i < h->md.nlocal; i++) cluster_of_local[i] = (int)S.clusterID[i];
}
// raw inputs of FindMolecule: tmpid [nlocal][12] (local index of the partner, 0 = none) and the averaged abo columns
void orc_md_species_raw(void* hh, int* tmpid, double* avg) {
  OrcHandle* h = (OrcHandle*)hh;
  const SpeciesFix& S = h->species;
  const size_t m = (size_t)h->md.nlocal * MAXSPECBOND;
  for (size_t k = 0; k < m; k++) { tmpid[k] = S.tmpid[k]; avg[k] = S.array[k]; }
}
long orc_md_species_text(void* hh, long ntimestep, char* out, long cap) {
  OrcHandle* h = (OrcHandle*)hh;
  h->text = h->species.formulas_text(ntimestep);
  return copy_text(h->text, out, cap);
}

// `position` output of fix reax/c/species at the last output step; box6 = boxlo[3], boxhi[3]; avg_qxyz (optional) receives
// the averaged q, x, y, z columns [nlocal][4] as they were BEFORE WritePos shifted them
long orc_md_species_pos_text(void* hh, long ntimestep, const double* box6, double* avg_qxyz, char* out, long cap) {
  OrcHandle* h = (OrcHandle*)hh;
  if (avg_qxyz) memcpy(avg_qxyz, h->species.qxyz.data(), h->species.qxyz.size() * sizeof(double));
  h->text = h->species.pos_text(h->md, ntimestep, box6);
  return copy_text(h->text, out, cap);
}

// OpenMP threads of this library: set > 0 sets the count (independent of OMP_NUM_THREADS, which launchers such as torchrun
// force to 1); returns the number of threads a parallel region actually gets (1 for the serial build).
int orc_omp_threads(int set) {
#ifdef _OPENMP
  if (set > 0) omp_set_num_threads(set);
  int got = 1;
#pragma omp parallel
  {
#pragma omp master
    got = omp_get_num_threads();
  }
  return got;
#else
  (void)set;
  return 1;
#endif
}

// fix qeq/reax <param file>: chi / eta / gamma per ELEMENT index (the tests map LAMMPS types 1:1 onto elements), as
// FixQEqReaxSunway::pertype_parameters + init_shielding would hold them (fix_qeq_reax_sunway.cpp:198-245, 440-454)
void orc_qeq_override(void* hh, const double* chi, const double* eta, const double* gamma) {
  QEq& q = ((OrcHandle*)hh)->md.qeq;
  const int nt = (int)q.chi.size();
  for (int i = 0; i < nt; i++) { q.chi[i] = chi[i]; q.eta[i] = eta[i]; q.gamma[i] = gamma[i]; }
  for (int i = 0; i < nt; i++)
    for (int j = 0; j < nt; j++) q.shld[(size_t)i * nt + j] = pow(q.gamma[i] * q.gamma[j], -1.5);
}

int orc_md_matvecs(void* hh, int which) { QEq& q = ((OrcHandle*)hh)->md.qeq; return which ? q.matvecs_t : q.matvecs_s; }

}  // extern "C"
