// extern "C" layer of librxb200.so: thin, exception -> status translation only (see include/rxb200.h).
#include <cuda_profiler_api.h>

#include <cstring>
#include <string>

#include "../../include/rxb200.h"
#include "rxb_system.h"

using rxb::System;

struct rxb_handle {
  System* sys = nullptr;
  int lgvdw = 0, enobonds = 1;
  std::string control;
};

namespace {
thread_local std::string g_err;

template <class F>
int guard(F&& fn) {
  try {
    fn();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  } catch (...) {
    g_err = "unknown error";
    return -1;
  }
}

template <class T>
void d2h(T* dst, const T* src, size_t n, cudaStream_t st) {
  if (n == 0) return;
  RXB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, st));
  RXB_SYNC(st);
}

// fp64 FMA throughput of this GPU, measured: 8 independent dependent-FMA chains per thread (enough ILP to cover the DFMA
// latency at full occupancy), nothing but DFMA in the loop.  The roofline denominator of the fp64-issue-bound kernels.
__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double a, double b, double* __restrict__ out) {
  double r[8];
#pragma unroll
  for (int k = 0; k < 8; k++) r[k] = 1.0 + 1e-3 * (threadIdx.x + k);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) r[k] = fma(r[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += r[k];
  if (s == 123.456) out[0] = s;   // never true: keeps the chains alive
}
}  // namespace

extern "C" {

double rxb_measure_fp64_tflops(int cuda_device) {
  double best = -1.0;
  guard([&] {
    RXB_CUDA(cudaSetDevice(cuda_device));
    double* out = nullptr;
    RXB_CUDA(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    RXB_CUDA(cudaEventCreate(&e0)); RXB_CUDA(cudaEventCreate(&e1));
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    for (int rep = 0; rep < 6; rep++) {
      RXB_CUDA(cudaEventRecord(e0));
      k_dfma_peak<<<blocks, threads>>>(iters, 0.999999, 1e-7, out);
      RXB_CUDA(cudaEventRecord(e1));
      RXB_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      RXB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  });
  return best;
}

const char* rxb_last_error(void) { return g_err.c_str(); }

int rxb_create(int cuda_device, rxb_handle** out) {
  return guard([&] {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) throw std::runtime_error("rxb_create: no CUDA device available (this path has no CPU fallback)");
    rxb_handle* h = new rxb_handle();
    h->sys = new System(cuda_device);
    *out = h;
  });
}

void rxb_destroy(rxb_handle* h) {
  if (!h) return;
  delete h->sys;
  delete h;
}

int rxb_pair_settings(rxb_handle* h, const char* control_file, int lgvdw, int enobonds) {
  return guard([&] {
    auto& ff = h->sys->ff;
    ff.ctl.lgflag = lgvdw;       // defaults: pair_reaxc_sunway.cpp:235-241
    ff.ctl.enobondsflag = enobonds;
    std::string e = ff.load_control(control_file);
    if (!e.empty()) throw std::runtime_error(e);
  });
}

int rxb_pair_coeff(rxb_handle* h, const char* ffield_file, int ntypes, const char* const* elements) {
  return guard([&] {
    auto& ff = h->sys->ff;
    std::string e = ff.load_ffield(ffield_file);
    if (e.empty()) e = ff.set_elements(ntypes, elements);
    if (!e.empty()) throw std::runtime_error(e);
    h->sys->upload_params();
  });
}

int rxb_pair_extract(rxb_handle* h, const char* name, double* out, int ntypes) {
  return guard([&] {
    const auto& ff = h->sys->ff;
    for (int t = 0; t <= ntypes; t++) {
      out[t] = 0.0;
      if (t >= 1 && t < (int)ff.map.size() && ff.map[t] >= 0) {
        const rxb::AtomPar& a = ff.atom[ff.map[t]];
        if (!strcmp(name, "chi")) out[t] = a.chi;
        else if (!strcmp(name, "eta")) out[t] = a.eta;
        else if (!strcmp(name, "gamma")) out[t] = a.gamma;
        else throw std::runtime_error(std::string("rxb_pair_extract: unknown quantity ") + name);
      }
    }
  });
}

int rxb_fix_qeq(rxb_handle* h, double swa, double swb, double tolerance, int max_iter) {
  return guard([&] {
    if (swb < 0) throw std::runtime_error("Fix qeq/reax has negative upper Taper radius cutoff");
    h->sys->qeq_swa = swa; h->sys->qeq_swb = swb; h->sys->qeq_tol = tolerance;
    if (max_iter > 0) h->sys->qeq_imax = max_iter;
  });
}

int rxb_neighbor_skin(rxb_handle* h, double skin) { return guard([&] { h->sys->skin = skin; }); }

long rxb_params_dump(rxb_handle* h, double* out, long cap) {
  std::vector<double> v = h->sys->params_dump();
  if (out) for (long i = 0; i < (long)v.size() && i < cap; i++) out[i] = v[i];
  return (long)v.size();
}

int rxb_set_atoms(rxb_handle* h, int nlocal, int nghost, const double* x, const int* type, const int* tag, const double* q,
                  const int* ghost_owner) {
  return guard([&] { h->sys->set_atoms(nlocal, nghost, x, type, tag, q, ghost_owner); });
}
static void check_nall(const System& s, int nall, const char* who) {
  if (nall != s.N)
    throw std::runtime_error(std::string(who) + ": caller passes nall = " + std::to_string(nall) + " but the device holds " +
                             std::to_string(s.N) + " atoms (rxb_set_atoms + rxb_neigh_build must follow every borders/exchange)");
}
int rxb_set_positions(rxb_handle* h, int nall, const double* x) {
  return guard([&] { check_nall(*h->sys, nall, "rxb_set_positions"); h->sys->set_positions(x); });
}
int rxb_set_charges(rxb_handle* h, const double* q) { return guard([&] { h->sys->set_charges(q); }); }
int rxb_neigh_build(rxb_handle* h) { return guard([&] { h->sys->build_neighbors(); }); }

int rxb_qeq_pre_force(rxb_handle* h, int* matvecs2) {
  return guard([&] {
    // matvecs2 == NULL: nobody needs the iteration counts now, so the solve is enqueued without a host round trip and
    // settled inside rxb_pair_compute (rxb_qeq_matvecs reports the counts afterwards)
    h->sys->plugin_qeq_pre_force(matvecs2 != nullptr);
    if (matvecs2) { matvecs2[0] = h->sys->matvecs_s; matvecs2[1] = h->sys->matvecs_t; }
  });
}
int rxb_qeq_matvecs(rxb_handle* h, int* matvecs2) {
  return guard([&] {
    h->sys->qeq_settle_now();
    matvecs2[0] = h->sys->matvecs_s; matvecs2[1] = h->sys->matvecs_t;
  });
}
int rxb_get_counters(rxb_handle* h, long long* out4) {
  return guard([&] {
    System& s = *h->sys;
    unsigned long long a = 0;
    d2h(&a, s.spmv_active_d.p, (size_t)1, s.stream());
    out4[0] = (long long)a; out4[1] = s.qeq_replays; out4[2] = s.qeq_iters_total; out4[3] = s.kernel_launches;
  });
}
long long rxb_host_sync_count(void) { return host_sync_counter(); }
int rxb_debug_set_caps(rxb_handle* h, int row_cap, int strong_cap, int cap_bonds, int cap_ang, int cap_tor, int cap_hb) {
  return guard([&] { h->sys->debug_set_caps(row_cap, strong_cap, cap_bonds, cap_ang, cap_tor, cap_hb); });
}
int rxb_debug_get_caps(rxb_handle* h, int* out6) {
  return guard([&] { h->sys->debug_get_caps(out6); });
}
int rxb_spec_atom_abo(rxb_handle* h, double* abo12) { return guard([&] { h->sys->spec_atom_abo(abo12); }); }
int rxb_fix_qeq_params(rxb_handle* h, int ntypes, const double* chi, const double* eta, const double* gamma) {
  return guard([&] { h->sys->qeq_set_type_params(ntypes, chi, eta, gamma); });
}
int rxb_set_h_exact(rxb_handle* h, int on) { return guard([&] { h->sys->h_exact_request = on != 0; }); }
int rxb_qeq_set_history(rxb_handle* h, const double* s_hist, const double* t_hist) {
  return guard([&] { h->sys->qeq_set_history(s_hist, t_hist); });
}
int rxb_qeq_get_history(rxb_handle* h, double* s_hist, double* t_hist) {
  return guard([&] { h->sys->qeq_get_history(s_hist, t_hist); });
}
int rxb_get_charges(rxb_handle* h, double* q) { return guard([&] { h->sys->get_charges(q); }); }

static void fill_pvector(const double* e, double* pvector, double* eng2) {
  using namespace rxb;
  if (pvector) {  // pair_reaxc_sunway.cpp:657-670
    pvector[0] = e[E_BOND]; pvector[1] = e[E_OV] + e[E_UN]; pvector[2] = e[E_LP]; pvector[3] = 0.0;
    pvector[4] = e[E_ANG]; pvector[5] = e[E_PEN]; pvector[6] = e[E_COA]; pvector[7] = e[E_HB];
    pvector[8] = e[E_TOR]; pvector[9] = e[E_CON]; pvector[10] = e[E_VDW]; pvector[11] = e[E_ELE];
    pvector[12] = 0.0; pvector[13] = e[E_POL];
  }
  if (eng2) {  // evdwl / ecoul split, pair_reaxc_sunway.cpp:636-650
    eng2[0] = e[E_BOND] + e[E_OV] + e[E_UN] + e[E_LP] + e[E_ANG] + e[E_PEN] + e[E_COA] + e[E_HB] + e[E_TOR] + e[E_CON] + e[E_VDW];
    eng2[1] = e[E_ELE] + e[E_POL];
  }
}

int rxb_pair_compute(rxb_handle* h, int nall, int eflag, int vflag, double* f_out, double* pvector, double* eng2,
                     double* virial6) {
  return guard([&] {
    System& s = *h->sys;
    check_nall(s, nall, "rxb_pair_compute");
    s.plugin_compute(eflag != 0, vflag != 0);
    if (f_out) s.get_forces(f_out);
    fill_pvector(s.energies, pvector, eng2);
    if (virial6) memcpy(virial6, s.virial, 6 * sizeof(double));
  });
}

int rxb_md_setup(rxb_handle* h, const double* box6, int nlocal, const double* x, const double* v, const int* type,
                 const int* tag, const double* mass, int ntypes, double dt, int reneigh_every, int thermo_every, int qeq_on) {
  return guard([&] {
    h->sys->qeq_on = qeq_on != 0;
    h->sys->md_thermo = thermo_every;
    h->sys->md_setup(box6, nlocal, x, v, type, tag, mass, ntypes, dt, reneigh_every);
  });
}
int rxb_md_run(rxb_handle* h, int nsteps) { return guard([&] { h->sys->md_run(nsteps); }); }
double rxb_md_last_run_ms(rxb_handle* h) { return h->sys->last_run_ms; }
int rxb_md_get(rxb_handle* h, double* x, double* v, double* f, double* q) { return guard([&] { h->sys->md_get(x, v, f, q); }); }
int rxb_md_get_tags(rxb_handle* h, int* tags) {
  return guard([&] { System& s = *h->sys; d2h(tags, s.tag.p, (size_t)s.n, s.stream()); });
}
int rxb_md_thermo(rxb_handle* h, double* pvector, double* pe, double* ke) {
  return guard([&] {
    System& s = *h->sys;
    double eng2[2];
    fill_pvector(s.energies, pvector, eng2);
    if (pe) *pe = eng2[0] + eng2[1];
    if (ke) *ke = s.md_kinetic();
  });
}

int rxb_get_counts(rxb_handle* h, long long* c) {
  return guard([&] {
    System& s = *h->sys;
    c[0] = s.n; c[1] = s.N; c[2] = s.vl.nnz; c[3] = s.bc.nnz; c[4] = s.num_bonds;
    long long far = 0;
    if (s.n > 0 && s.far_num.n >= (size_t)s.n) {
      std::vector<int> num(s.n);
      d2h(num.data(), s.far_num.p, s.n, s.stream());
      for (int v : num) far += v;
    }
    c[5] = far; c[6] = s.kernel_launches; c[7] = s.qeq_iters_total;
  });
}

int rxb_get_neighbors(rxb_handle* h, int which, long long* off, int* idx) {
  return guard([&] {
    System& s = *h->sys;
    if (which == 0) { s.export_verlet(off, idx); return; }   // the Verlet list lives in S space: translated on the host
    rxb::Csr& c = s.bc;
    // the device rows sit at a fixed stride: hand out a compact CSR
    std::vector<int> cnt(c.nrows), raw((size_t)std::max<long long>(c.slots, 1));
    d2h(cnt.data(), c.cnt.p, (size_t)c.nrows, s.stream());
    d2h(raw.data(), c.idx.p, (size_t)c.slots, s.stream());
    long long w = 0;
    for (int i = 0; i < c.nrows; i++) {
      off[i] = w;
      memcpy(idx + w, raw.data() + (size_t)i * c.stride, (size_t)cnt[i] * sizeof(int));
      w += cnt[i];
    }
    off[c.nrows] = w;
  });
}

int rxb_get_bonds(rxb_handle* h, int* b_start, int* b_cnt, int* nbr, int* sym, double* fld) {
  return guard([&] {
    System& s = *h->sys;
    const size_t N = s.N, nb = s.num_bonds;
    d2h(b_start, s.b_start.p, N, s.stream());
    d2h(b_cnt, s.b_cnt.p, N, s.stream());
    d2h(nbr, s.b_nbr.p, nb, s.stream());
    d2h(sym, s.b_sym.p, nb, s.stream());
    std::vector<double4> geo(nb), bo(nb), der(nb), c1(nb), c2(nb), c3(nb);
    std::vector<double> cd(nb), cdpi(nb), cdpi2(nb);
    d2h(geo.data(), s.b_geo.p, nb, s.stream()); d2h(bo.data(), s.b_bo.p, nb, s.stream());
    d2h(der.data(), s.b_der.p, nb, s.stream()); d2h(c1.data(), s.b_c1.p, nb, s.stream());
    d2h(c2.data(), s.b_c2.p, nb, s.stream()); d2h(c3.data(), s.b_c3.p, nb, s.stream());
    d2h(cd.data(), s.b_Cdbo.p, nb, s.stream()); d2h(cdpi.data(), s.b_Cdbopi.p, nb, s.stream());
    d2h(cdpi2.data(), s.b_Cdbopi2.p, nb, s.stream());
    // effective coefficients: the per-centre angle sums (CEval5/CEval6) are applied to every bond of the centre inside
    // K-dbond; fold them in here so that the export equals the reference's Cdbo/Cdbopi/Cdbopi2 arrays
    std::vector<double2> s56(N);
    std::vector<int> bs(N), bcn(N);
    d2h(s56.data(), s.sum56.p, N, s.stream());
    d2h(bs.data(), s.b_start.p, N, s.stream()); d2h(bcn.data(), s.b_cnt.p, N, s.stream());
    for (size_t i = 0; i < N; i++)
      for (int p = bs[i]; p < bs[i] + bcn[i]; p++) {
        const double b = bo[p].x, b3 = b * b * b;
        cd[p] += s56[i].y * (b3 * b3 * b); cdpi[p] += s56[i].x; cdpi2[p] += s56[i].x;
      }
    for (size_t p = 0; p < nb; p++) {
      double* o = fld + 31 * p;
      const double dv[3] = {geo[p].y, geo[p].z, geo[p].w};
      o[0] = geo[p].x; o[1] = dv[0]; o[2] = dv[1]; o[3] = dv[2];
      o[4] = bo[p].x; o[5] = bo[p].y; o[6] = bo[p].z; o[7] = bo[p].w;
      for (int t = 0; t < 3; t++) { o[8 + t] = der[p].x * dv[t]; o[11 + t] = der[p].y * dv[t]; o[14 + t] = der[p].z * dv[t]; }
      o[17] = c1[p].x; o[18] = c1[p].y; o[19] = c1[p].z;
      o[20] = c1[p].w; o[21] = c2[p].x; o[22] = c2[p].y; o[23] = c2[p].z;
      o[24] = c2[p].w; o[25] = c3[p].x; o[26] = c3[p].y; o[27] = c3[p].z;
      o[28] = cd[p]; o[29] = cdpi[p]; o[30] = cdpi2[p];
    }
  });
}

int rxb_get_hbond_pairs(rxb_handle* h, int* n_out, int* pairs2, int cap) {
  return guard([&] {
    System& s = *h->sys;
    int cnt[4] = {0, 0, 0, 0};
    d2h(cnt, s.it_count.p, (size_t)4, s.stream());
    const int m = cnt[2] < s.cap_hb ? cnt[2] : s.cap_hb;
    if (n_out) *n_out = m;
    if (pairs2 && m > 0) {
      std::vector<int4> items((size_t)m);
      d2h(items.data(), s.it_hb.p, (size_t)m, s.stream());
      for (int k = 0; k < m && k < cap; k++) { pairs2[2 * k] = items[k].x; pairs2[2 * k + 1] = items[k].y; }
    }
  });
}

int rxb_get_workspace(rxb_handle* h, double* w16) {
  return guard([&] {
    System& s = *h->sys;
    const size_t N = s.N;
    std::vector<double> a(N);
    std::vector<double2> dp(N);
    auto col = [&](const double* src, int c) { d2h(a.data(), src, N, s.stream()); for (size_t i = 0; i < N; i++) w16[16 * i + c] = a[i]; };
    memset(w16, 0, N * 16 * sizeof(double));
    col(s.total_bo.p, 0); col(s.Delta_boc.p, 1);
    d2h(dp.data(), s.Deltap.p, N, s.stream());
    for (size_t i = 0; i < N; i++) { w16[16 * i + 2] = dp[i].x; w16[16 * i + 3] = dp[i].y; }
    col(s.Delta.p, 4); col(s.Delta_val.p, 6); col(s.vlpex.p, 7); col(s.nlp.p, 8); col(s.Delta_lp.p, 9);
    col(s.dDelta_lp.p, 10); col(s.dDelta_lp.p, 11); col(s.Delta_lp_temp.p, 13); col(s.CdDelta.p, 15);
  });
}

int rxb_get_far(rxb_handle* h, int* num, int* idx, double* val) {
  return guard([&] {
    // compact like rxb_get_neighbors(0): the entries of atom i start at the compact Verlet offset of atom i
    h->sys->export_far(num, idx, val);
  });
}

int rxb_profile(rxb_handle* h, int enable, double* ms9) {
  return guard([&] {
    System& s = *h->sys;
    s.resolve_timers();
    if (ms9) for (int k = 0; k < rxb::StepTimers::NUM; k++) { ms9[k] = s.timers.ms[k]; ms9[rxb::StepTimers::NUM + k] = (double)s.timers.calls[k]; }
    if (enable >= 0) { s.profile = enable != 0; if (enable) s.timers = rxb::StepTimers(); }
  });
}

long rxb_parse_dump(const char* control_file, const char* ffield_file, int ntypes, const char* const* elements, int lgvdw,
                    int enobonds, double* out, long cap) {
  long count = -1;
  int rc = guard([&] {
    rxb::ForceField ff;
    ff.ctl.lgflag = lgvdw;
    ff.ctl.enobondsflag = enobonds;
    std::string e = ff.load_control(control_file);
    if (e.empty()) e = ff.load_ffield(ffield_file);
    if (e.empty()) e = ff.set_elements(ntypes, elements);
    if (!e.empty()) throw std::runtime_error(e);
    std::vector<double> v = ff.dump();
    if (out) for (long i = 0; i < (long)v.size() && i < cap; i++) out[i] = v[i];
    count = (long)v.size();
  });
  return rc == 0 ? count : -1;
}

long rxb_lookup_dump(const char* control_file, const char* ffield_file, int ntypes, const char* const* elements, int* n_out,
                     double* out, long cap) {
  long count = -1;
  int rc = guard([&] {
    rxb::ForceField ff;
    std::string e = ff.load_control(control_file);
    if (e.empty()) e = ff.load_ffield(ffield_file);
    if (e.empty()) e = ff.set_elements(ntypes, elements);
    if (!e.empty()) throw std::runtime_error(e);
    ff.derive();
    if (ff.ctl.tabulate <= 0) { count = 0; if (n_out) *n_out = 0; return; }
    int n = 0; double dx = 0;
    std::vector<double> v = ff.lookup_tables(&n, &dx);
    if (n_out) *n_out = n;
    if (out) for (long i = 0; i < (long)v.size() && i < cap; i++) out[i] = v[i];
    count = (long)v.size();
  });
  return rc == 0 ? count : -1;
}

int rxb_dist_unique_id(char* out128) { return guard([&] { System::dist_unique_id(out128); }); }
int rxb_dist_init(rxb_handle* h, int rank, int world, const char* id128, int px, int py, int pz) {
  return guard([&] { h->sys->dist_init(rank, world, id128, px, py, pz); });
}

int rxb_dist_set_p2p(rxb_handle* h, int on) { return guard([&] { h->sys->dist_set_p2p(on != 0); }); }

int rxb_comm_init(rxb_handle* h, int rank, int world, const char* id128) {
  return guard([&] { h->sys->comm_init(rank, world, id128); });
}
int rxb_comm_set_ghosts(rxb_handle* h, int nghost, const int* owner_rank, const int* owner_index) {
  return guard([&] { h->sys->comm_set_ghosts(nghost, owner_rank, owner_index); });
}

int rxb_bond_table(rxb_handle* h, double bo_cut, int* nlocal, int* nentries, int* max_per_atom) {
  return guard([&] {
    auto t = h->sys->bond_table_build(bo_cut);
    if (nlocal) *nlocal = t.n;
    if (nentries) *nentries = t.entries;
    if (max_per_atom) *max_per_atom = t.max_nb;
  });
}
int rxb_bond_table_get(rxb_handle* h, int* tag, int* type, int* off, int* nbr_tag, double* bo, double* abo, double* nlp,
                       double* q) {
  return guard([&] { h->sys->bond_table_get(tag, type, off, nbr_tag, bo, abo, nlp, q); });
}

int rxb_species_config(rxb_handle* h, int nevery, int nrepeat, int nfreq, int ntypes, const double* bocut, long natoms,
                       long ntimestep_now, int* reneighbor_reset) {
  return guard([&] {
    int r = h->sys->species_config(nevery, nrepeat, nfreq, ntypes, bocut, natoms, ntimestep_now);
    if (reneighbor_reset) *reneighbor_reset = r;
  });
}
int rxb_species_step(rxb_handle* h, long ntimestep, int* found) {
  return guard([&] { bool f = h->sys->species_step(ntimestep); if (found) *found = f ? 1 : 0; });
}
int rxb_species_result(rxb_handle* h, int* nmole, int* composition, long cap) {
  return guard([&] {
    const auto& S = h->sys->species;
    if (nmole) *nmole = S.nmole;
    if (composition) for (long i = 0; i < (long)S.composition.size() && i < cap; i++) composition[i] = S.composition[i];
  });
}
int rxb_species_cluster(rxb_handle* h, int* cluster_of_local) {
  return guard([&] { h->sys->species_get_cluster(cluster_of_local); });
}
int rxb_species_avg_qxyz(rxb_handle* h, double* qxyz4) {
  return guard([&] { h->sys->species_avg_qxyz(qxyz4); });
}
int rxb_species_log_size(rxb_handle* h) { return (int)h->sys->species_log.size(); }
int rxb_species_log_get(rxb_handle* h, int k, long* step, int* nmole, int* composition, long cap) {
  return guard([&] {
    const auto& L = h->sys->species_log;
    if (k < 0 || k >= (int)L.size()) throw std::runtime_error("rxb_species_log_get: no such record");
    if (step) *step = L[k].step;
    if (nmole) *nmole = L[k].nmole;
    if (composition) for (long i = 0; i < (long)L[k].composition.size() && i < cap; i++) composition[i] = L[k].composition[i];
  });
}

int rxb_get_cutoffs(rxb_handle* h, double* out3) {
  return guard([&] {
    const System& s = *h->sys;
    out3[0] = s.cutneigh(); out3[1] = s.bond_reach() + s.skin; out3[2] = s.bond_reach();
  });
}

int rxb_host_register(void* p, size_t bytes) {
  return guard([&] { RXB_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable)); });
}
int rxb_host_unregister(void* p) { return guard([&] { RXB_CUDA(cudaHostUnregister(p)); }); }

int rxb_get_h_format(rxb_handle* h, int* bytes_per_entry, char* name, int cap) {
  return guard([&] {
    const char* nm = h->sys->h_format_name();
    if (bytes_per_entry) *bytes_per_entry = h->sys->h_bytes_per_entry();
    if (name && cap > 0) { strncpy(name, nm, cap - 1); name[cap - 1] = 0; }
  });
}

int rxb_profiler_range(int start) {
  return guard([&] { if (start) RXB_CUDA(cudaProfilerStart()); else RXB_CUDA(cudaProfilerStop()); });
}

}  // extern "C"
