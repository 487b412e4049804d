// Analysis outputs that read the resident bond list: the connection table of `fix reax/c/bonds` and the molecule /
// species census of `fix reax/c/species`.  Both run only on output / sampling steps and copy
// to the host only what is written to the file.
//
//   f1  FixReaxCBondsSunway::FindBond + PassBuffer (fix_reaxc_bonds_sunway.cpp:187-260): per local atom the neighbours
//       with BO > bg_cut in bond-row order, plus abo = total bond order (bo_dboc[i][0]), nlp, q.
//       Here: count kernel -> exclusive scan -> fill kernel; one compact CSR instead of [nmax][MAXREAXBOND] arrays
//       (the reference's fixed 12 columns silently overflow; this table has no per-atom limit).
//   f2  PairReaxCSunway::FindBond (pair_reaxc_sunway.cpp:1170-1198): tmpid/tmpbo = bonds with j >= i and BO >= 0.10,
//       at most MAXSPECBOND 12 per atom; averaged slot by slot by `fix ave/atom nevery nrepeat nfreq`
//       (fix_reaxc_species_sunway.cpp:377-425); FindMolecule (:498-566) = connected components over the averaged
//       bond orders > BOCut[itype][jtype], cluster ID = smallest atom ID of the component; SortMolecule (:570-648)
//       renumbers 1..Nmole in ascending ID order; FindSpecies (:652-717) counts atoms per type and molecule.
//       Here: the reference's iterated min-label sweeps + halo exchange become one lock-free union-find over an edge
//       list in atom-ID space (hook the larger root under the smaller, so the root IS the smallest ID), all-gathered
//       across ranks when the run is decomposed — the fixed point is the same labelling.
#include <cub/cub.cuh>

#include <algorithm>

#include "rxb_system.h"

namespace rxb {

namespace {

constexpr int kMaxSpecBond = 12;   // MAXSPECBOND, reaxc_defs_sunway.h:124
constexpr double kSpecBoMin = 0.10;  // pair_reaxc_sunway.cpp:1178

inline int nblk(long n, int t = 256) { return (int)((n + t - 1) / t); }

// ---------------------------------------------------------------------------------------------- f1: bond table
template <bool FILL>
__global__ void k_bond_table(int n, double cut, const int* __restrict__ b_start, const int* __restrict__ b_cnt,
                             const int* __restrict__ b_nbr, const double4* __restrict__ b_bo,
                             const int* __restrict__ tag, int* __restrict__ cnt, const int* __restrict__ off,
                             int* __restrict__ out_tag, double* __restrict__ out_bo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = b_start[i], e = s + b_cnt[i];
  int k = FILL ? off[i] : 0;
  for (int p = s; p < e; p++) {
    const double bo = b_bo[p].x;
    if (bo > cut) {
      if (FILL) { out_tag[k] = tag[b_nbr[p]]; out_bo[k] = bo; }
      k++;
    }
  }
  if (!FILL) cnt[i] = k;
  if (!FILL && i == 0) cnt[n] = 0;
}

__global__ void k_max_int(int n, const int* __restrict__ v, int* __restrict__ out) {
  int m = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, v[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__global__ void k_gather_q(int n, const double4* __restrict__ xq, double* __restrict__ q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) q[i] = xq[i].w;
}

// ---------------------------------------------------------------------------------------------- f2: species
// one sample of compute SPEC/ATOM's abo columns: slot k of atom i holds the k-th bond (row order) with nbr >= i, BO >= 0.10
__global__ void k_spec_sample(int n, const int* __restrict__ b_start, const int* __restrict__ b_cnt,
                              const int* __restrict__ b_nbr, const double4* __restrict__ b_bo, int* __restrict__ ids,
                              double* __restrict__ acc, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = b_start[i], e = s + b_cnt[i];
  int k = 0;
  for (int p = s; p < e; p++) {
    const int j = b_nbr[p];
    if (j < i) continue;
    const double bo = b_bo[p].x;
    if (bo >= kSpecBoMin) {
      if (k < kMaxSpecBond) { ids[i * kMaxSpecBond + k] = j; acc[i * kMaxSpecBond + k] += bo; }
      k++;
    }
  }
  if (k > kMaxSpecBond) atomicMax(err, k);
  for (; k < kMaxSpecBond; k++) ids[i * kMaxSpecBond + k] = 0;   // tmpid is zeroed before every FindBond (:783-790)
}

// one sample of compute SPEC/ATOM's q, x, y, z columns (compute_spec_atom_sunway.cpp:142-170), summed over the window
__global__ void k_spec_qxyz(int n, const double4* __restrict__ xq, double* __restrict__ acc4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = xq[i];
  acc4[4 * i] += p.w; acc4[4 * i + 1] += p.x; acc4[4 * i + 2] += p.y; acc4[4 * i + 3] += p.z;
}

// edges of the molecule graph in atom-ID space; COUNT pass sizes the list, FILL pass writes it
template <bool FILL>
__global__ void k_spec_edges(int n, int ntypes, double nrepeat, const int* __restrict__ ids,
                             const double* __restrict__ acc, const int* __restrict__ ltype, const int* __restrict__ tag,
                             const double* __restrict__ bocut, int* __restrict__ cursor, int* __restrict__ edges) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int it = ltype[i];
  int mine = 0;
  int ej[kMaxSpecBond];
  for (int k = 0; k < kMaxSpecBond; k++) {
    const int j = ids[i * kMaxSpecBond + k];
    if (j == 0 || j < i) continue;                    // fix_reaxc_species_sunway.cpp:533
    const double bo = acc[i * kMaxSpecBond + k] / nrepeat;   // fix ave/atom divides the sum
    if (bo > bocut[it * (ntypes + 1) + ltype[j]]) ej[mine++] = tag[j];
  }
  if (mine == 0) return;
  const int at = atomicAdd(cursor, mine);
  if (FILL) {
    const int ti = tag[i];
    for (int k = 0; k < mine; k++) { edges[2 * (at + k)] = ti; edges[2 * (at + k) + 1] = ej[k]; }
  }
}

__global__ void k_iota(int m, int* __restrict__ parent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) parent[i] = i;
}

__device__ __forceinline__ int uf_find(int* parent, int x) {
  while (true) {
    const int p = parent[x];
    if (p == x) return x;
    const int gp = parent[p];
    if (gp != p) parent[x] = gp;   // path halving; racing writers only ever store ancestors
    x = p;
  }
}

__global__ void k_union(int nedges, const int* __restrict__ edges, int* __restrict__ parent) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nedges) return;
  int a = edges[2 * e], b = edges[2 * e + 1];
  if (a < 0) return;   // padding of the gathered list
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) break;
    if (a < b) { const int t = a; a = b; b = t; }
    if (atomicCAS(&parent[a], a, b) == a) break;   // larger root hooks under the smaller one
  }
}

__global__ void k_flatten(int m, int* __restrict__ parent, int* __restrict__ root_flag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  int x = t;
  while (parent[x] != x) x = parent[x];
  parent[t] = x;   // roots never change in this kernel, so concurrent flattening is safe
  root_flag[t] = (t > 0 && x == t) ? 1 : 0;
  if (t == 0) root_flag[m] = 0;
}

__global__ void k_composition(int n, int ntypes, const int* __restrict__ tag, const int* __restrict__ ltype,
                              const int* __restrict__ parent, const int* __restrict__ molidx, int* __restrict__ comp,
                              int* __restrict__ cluster) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int m = molidx[parent[tag[i]]];
  cluster[i] = m + 1;
  atomicAdd(&comp[m * ntypes + ltype[i] - 1], 1);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
System::BondTable System::bond_table_build(double bo_cut) {
  if (bo_cut < 0) bo_cut = ff.ctl.bg_cut;
  BondTable t;
  t.n = n;
  bt_cnt.resize((size_t)n + 1);
  bt_off.resize((size_t)n + 1);
  sp_misc.resize(4);
  RXB_CUDA(cudaMemsetAsync(sp_misc.p, 0, 4 * sizeof(int), st_));
  if (n == 0) return t;
  k_bond_table<false><<<nblk(n), 256, 0, st_>>>(n, bo_cut, b_start.p, b_cnt.p, b_nbr.p, b_bo.p, tag.p, bt_cnt.p, nullptr,
                                               nullptr, nullptr);
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, bt_cnt.p, bt_off.p, n + 1, st_);
  scan_temp.resize(need + 16);
  cub::DeviceScan::ExclusiveSum(scan_temp.p, need, bt_cnt.p, bt_off.p, n + 1, st_);
  k_max_int<<<std::min(nblk(n), 1024), 256, 0, st_>>>(n, bt_cnt.p, sp_misc.p);
  int got[2] = {0, 0};
  RXB_CUDA(cudaMemcpyAsync(&got[0], bt_off.p + n, sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaMemcpyAsync(&got[1], sp_misc.p, sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
  t.entries = got[0];
  t.max_nb = got[1];
  bt_tag.resize((size_t)std::max(t.entries, 1));
  bt_bo.resize((size_t)std::max(t.entries, 1));
  k_bond_table<true><<<nblk(n), 256, 0, st_>>>(n, bo_cut, b_start.p, b_cnt.p, b_nbr.p, b_bo.p, tag.p, nullptr, bt_off.p,
                                              bt_tag.p, bt_bo.p);
  RXB_CUDA(cudaGetLastError());
  kernel_launches += 3;
  bt_last_ = t;
  return t;
}

void System::bond_table_get(int* tag_out, int* type_out, int* off_out, int* nbr_tag, double* bo, double* abo,
                            double* nlp_out, double* q_out) {
  const BondTable& t = bt_last_;
  if (t.n != n) throw std::runtime_error("bond table is stale: call rxb_bond_table first");
  if (n == 0) { if (off_out) off_out[0] = 0; return; }
  auto d2h = [&](void* dst, const void* src, size_t bytes) {
    if (dst && bytes) RXB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st_));
  };
  d2h(tag_out, tag.p, (size_t)n * sizeof(int));
  d2h(type_out, ltype_d.p, (size_t)n * sizeof(int));
  d2h(off_out, bt_off.p, ((size_t)n + 1) * sizeof(int));
  d2h(nbr_tag, bt_tag.p, (size_t)t.entries * sizeof(int));
  d2h(bo, bt_bo.p, (size_t)t.entries * sizeof(double));
  d2h(abo, total_bo.p, (size_t)n * sizeof(double));
  d2h(nlp_out, nlp.p, (size_t)n * sizeof(double));
  if (q_out) {
    x_stage.resize(std::max(x_stage.n, (size_t)n));
    k_gather_q<<<nblk(n), 256, 0, st_>>>(n, xq.p, x_stage.p);
    d2h(q_out, x_stage.p, (size_t)n * sizeof(double));
  }
  RXB_SYNC(st_);
}

// ---------------------------------------------------------------------------------------------------------------------
int System::species_config(int nevery, int nrepeat, int nfreq, int ntypes, const double* bocut, long natoms_total, long now) {
  if (now < 0) now = ntimestep;
  if (nevery <= 0 || nrepeat <= 0 || nfreq <= 0 || nfreq % nevery || (long)nrepeat * nevery > nfreq)
    throw std::runtime_error("Illegal fix reax/c/species command");   // fix_reaxc_species_sunway.cpp:76-79
  // neighbour lists (here: atom order and bond-row slots) must stay unchanged while bond orders are averaged (:81-101)
  int reset = 0;
  if (nevery * nrepeat != 1 && (nfreq % md_every != 0 || md_every < nevery * nrepeat)) {
    int ne = nevery * nrepeat;
    while (nfreq % ne != 0 && ne <= nfreq / 2) ne++;
    if (nfreq % ne != 0) ne = nfreq;
    md_every = ne;
    reset = 1;
  }
  Species& S = species;
  S.on = true; S.nevery = nevery; S.nrepeat = nrepeat; S.nfreq = nfreq; S.ntypes = ntypes; S.natoms = natoms_total;
  S.irepeat = 0;
  S.nvalid_out = now + nfreq;                                   // :318-319
  // FixAveAtom::nextvalid (LAMMPS core, stock semantics): last nrepeat samples, nevery apart, ending on a multiple of nfreq
  long nv = (now / nfreq) * nfreq + nfreq;
  if (nv - nfreq == now && nrepeat == 1) nv = now;
  else nv -= (long)(nrepeat - 1) * nevery;
  if (nv < now) nv += nfreq;
  S.nvalid_ave = nv;
  sp_bocut.resize((size_t)(ntypes + 1) * (ntypes + 1));
  RXB_CUDA(cudaMemcpyAsync(sp_bocut.p, bocut, sp_bocut.n * sizeof(double), cudaMemcpyHostToDevice, st_));
  RXB_SYNC(st_);
  species_log.clear();
  return reset;
}

// compute SPEC/ATOM, the abo01..abo12 columns on their own (compute_spec_atom_sunway.cpp:142-170 reading
// PairReaxCSunway::tmpbo, pair_reaxc_sunway.cpp:1170-1198): ONE sample of the bond orders (>= 0.10) of each local atom's
// bonds to partners of higher index, in bond-row order, zero padded.  abo[nlocal][12].
void System::spec_atom_abo(double* abo_host) {
  RXB_CUDA(cudaSetDevice(device_));
  const size_t m = (size_t)std::max(n, 1) * kMaxSpecBond;
  DBuf<int> ids;
  DBuf<double> acc;
  ids.resize(m); acc.resize(m);
  RXB_CUDA(cudaMemsetAsync(acc.p, 0, m * sizeof(double), st_));
  sp_misc.resize(4);
  RXB_CUDA(cudaMemsetAsync(sp_misc.p, 0, 4 * sizeof(int), st_));
  if (n > 0) k_spec_sample<<<nblk(n), 256, 0, st_>>>(n, b_start.p, b_cnt.p, b_nbr.p, b_bo.p, ids.p, acc.p, sp_misc.p);
  int err = 0;
  RXB_CUDA(cudaMemcpyAsync(&err, sp_misc.p, sizeof(int), cudaMemcpyDeviceToHost, st_));
  if (n > 0) RXB_CUDA(cudaMemcpyAsync(abo_host, acc.p, (size_t)n * kMaxSpecBond * sizeof(double), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
  kernel_launches += 1;
  if (err > kMaxSpecBond) throw std::runtime_error("Increase MAXSPECBOND in reaxc_defs_sunway.h");   // pair_reaxc_sunway.cpp:1194
}

void System::species_sample() {
  Species& S = species;
  const size_t m = (size_t)std::max(n, 1) * kMaxSpecBond;
  if (S.irepeat == 0) {
    sp_id.resize(m); sp_acc.resize(m);
    RXB_CUDA(cudaMemsetAsync(sp_acc.p, 0, m * sizeof(double), st_));
    sp_qxyz.resize((size_t)4 * std::max(n, 1));
    RXB_CUDA(cudaMemsetAsync(sp_qxyz.p, 0, (size_t)4 * std::max(n, 1) * sizeof(double), st_));
    sp_n_ = n;
  }
  if (n != sp_n_) throw std::runtime_error("fix reax/c/species: atoms migrated inside an averaging window");
  sp_misc.resize(4);
  RXB_CUDA(cudaMemsetAsync(sp_misc.p, 0, 4 * sizeof(int), st_));
  if (n > 0)
    k_spec_sample<<<nblk(n), 256, 0, st_>>>(n, b_start.p, b_cnt.p, b_nbr.p, b_bo.p, sp_id.p, sp_acc.p, sp_misc.p);
  if (n > 0) k_spec_qxyz<<<nblk(n), 256, 0, st_>>>(n, xq.p, sp_qxyz.p);
  int err = 0;
  RXB_CUDA(cudaMemcpyAsync(&err, sp_misc.p, sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
  kernel_launches += 2;
  if (err > kMaxSpecBond) throw std::runtime_error("Increase MAXSPECBOND in reaxc_defs_sunway.h");   // pair_reaxc_sunway.cpp:1194
}

// the fix ave/atom result for the q, x, y, z columns: sums of the last complete window / nrepeat (`position` keyword)
void System::species_avg_qxyz(double* out4) {
  RXB_CUDA(cudaSetDevice(device_));
  if (!species.on || sp_qxyz.n < (size_t)4 * sp_n_ || sp_n_ != n)
    throw std::runtime_error("rxb_species_avg_qxyz: no complete averaging window for the current atoms");
  if (species.irepeat != 0) throw std::runtime_error("rxb_species_avg_qxyz: called inside an averaging window");
  RXB_CUDA(cudaMemcpyAsync(out4, sp_qxyz.p, (size_t)4 * n * sizeof(double), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
  for (size_t k = 0; k < (size_t)4 * n; k++) out4[k] /= species.nrepeat;
}

bool System::species_step(long step) {
  Species& S = species;
  if (!S.on) return false;
  if (step == S.nvalid_ave) {
    species_sample();
    S.irepeat++;
    if (S.irepeat < S.nrepeat) S.nvalid_ave += S.nevery;
    else { S.irepeat = 0; S.nvalid_ave = step + S.nfreq - (long)(S.nrepeat - 1) * S.nevery; }
  }
  if (step != S.nvalid_out) return false;
  species_find();
  S.nvalid_out += S.nfreq;
  return true;
}

void System::species_find() {
  Species& S = species;
  if (sp_id.n == 0 || sp_n_ != n) throw std::runtime_error("fix reax/c/species: no averaged bond orders to analyse");
  const int W = dist_world();
  const int T = S.ntypes;
  const int M = (int)S.natoms + 1;   // atom IDs 1..natoms (consecutive, as fix reax/c/bonds also requires)
  const double inv = (double)S.nrepeat;
  sp_misc.resize(4);
  RXB_CUDA(cudaMemsetAsync(sp_misc.p, 0, 4 * sizeof(int), st_));
  if (n > 0)
    k_spec_edges<false><<<nblk(n), 256, 0, st_>>>(n, T, inv, sp_id.p, sp_acc.p, ltype_d.p, tag.p, sp_bocut.p, sp_misc.p, nullptr);
  // edge counts of every rank
  DBuf<int>& cnts = sp_flag;   // reused below once the counts are on the host
  cnts.resize((size_t)std::max(W, M + 1));
  dist_allgather_int(sp_misc.p, cnts.p, 1);
  std::vector<int> hc(W);
  RXB_CUDA(cudaMemcpyAsync(hc.data(), cnts.p, W * sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
  const int mine = hc[dist_rank()];
  const int chunk = std::max(1, *std::max_element(hc.begin(), hc.end()));
  sp_edges.resize((size_t)2 * chunk);
  RXB_CUDA(cudaMemsetAsync(sp_edges.p, 0xff, (size_t)2 * chunk * sizeof(int), st_));   // -1 padding
  RXB_CUDA(cudaMemsetAsync(sp_misc.p, 0, sizeof(int), st_));
  if (n > 0 && mine > 0)
    k_spec_edges<true><<<nblk(n), 256, 0, st_>>>(n, T, inv, sp_id.p, sp_acc.p, ltype_d.p, tag.p, sp_bocut.p, sp_misc.p, sp_edges.p);
  const int* all = sp_edges.p;
  if (W > 1) {
    sp_edges_all.resize((size_t)2 * chunk * W);
    dist_allgather_int(sp_edges.p, sp_edges_all.p, (size_t)2 * chunk);
    all = sp_edges_all.p;
  }
  sp_parent.resize((size_t)M);
  sp_molidx.resize((size_t)M + 1);
  k_iota<<<nblk(M), 256, 0, st_>>>(M, sp_parent.p);
  k_union<<<nblk((long)chunk * W), 256, 0, st_>>>(chunk * W, all, sp_parent.p);
  k_flatten<<<nblk(M), 256, 0, st_>>>(M, sp_parent.p, sp_flag.p);
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, sp_flag.p, sp_molidx.p, M + 1, st_);
  scan_temp.resize(need + 16);
  cub::DeviceScan::ExclusiveSum(scan_temp.p, need, sp_flag.p, sp_molidx.p, M + 1, st_);
  int nmole = 0;
  RXB_CUDA(cudaMemcpyAsync(&nmole, sp_molidx.p + M, sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
  S.nmole = nmole;
  sp_comp.resize((size_t)std::max(1, nmole) * T + (size_t)std::max(n, 1));
  int* cluster = sp_comp.p + (size_t)std::max(1, nmole) * T;
  RXB_CUDA(cudaMemsetAsync(sp_comp.p, 0, (size_t)std::max(1, nmole) * T * sizeof(int), st_));
  if (n > 0) k_composition<<<nblk(n), 256, 0, st_>>>(n, T, tag.p, ltype_d.p, sp_parent.p, sp_molidx.p, sp_comp.p, cluster);
  if (W > 1) dist_allreduce_int(sp_comp.p, (size_t)nmole * T);
  S.composition.assign((size_t)nmole * T, 0);
  if (nmole > 0)
    RXB_CUDA(cudaMemcpyAsync(S.composition.data(), sp_comp.p, (size_t)nmole * T * sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
  RXB_CUDA(cudaGetLastError());
  kernel_launches += 6;
}

void System::species_get_cluster(int* cluster_of_local) {
  if (species.nmole == 0 || n == 0) return;
  const int* cluster = sp_comp.p + (size_t)std::max(1, species.nmole) * species.ntypes;
  RXB_CUDA(cudaMemcpyAsync(cluster_of_local, cluster, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
}

}  // namespace rxb
