// System: device-resident state + launch sequence of one timestep (see rxb_system.h).
//
// Step order follows the reference's Verlet/Compute_Forces order:
//   fix nve initial -> [reneighbour | ghost forward] -> qeq pre_force -> pair compute
//   (Reset -> bond list -> nonbonded -> BO -> multi-body -> hbonds -> valence/torsion -> dBond) -> reverse -> nve final
// /root/reference/pair_reaxc_sunway.cpp:541-793, reaxc_forces_sunway.cpp:1297-1365, fix_nve_sw64.c:43-170.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <unordered_map>

#include <nvtx3/nvToolsExt.h>

#include "rxb_system.h"

namespace rxb {

namespace {

constexpr double kFtm2v = 1.0 / 48.88821291 / 48.88821291;  // LAMMPS units real
constexpr double kMvv2e = 48.88821291 * 48.88821291;

__global__ void k_set_xyz(int N, const double* __restrict__ x, double4* __restrict__ xq) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double4 p = xq[i];
  p.x = x[3 * i]; p.y = x[3 * i + 1]; p.z = x[3 * i + 2];
  xq[i] = p;
}
// set_atoms: (x[3], q, LAMMPS type) -> xq record + force-field element index (write_reax_atoms_and_pack, pair_reaxc_sw64.c:114-190)
__global__ void k_pack_atoms(int N, const double* __restrict__ x, const double* __restrict__ q, const int* __restrict__ ltype,
                             const int* __restrict__ map, int nmap, double4* __restrict__ xq, int* __restrict__ type) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  xq[i] = make_double4(x[3 * i], x[3 * i + 1], x[3 * i + 2], q ? q[i] : 0.0);
  const int lt = ltype[i];
  type[i] = (lt >= 1 && lt < nmap) ? map[lt] : -1;
}

__global__ void k_set_q(int N, const double* __restrict__ q, double4* __restrict__ xq) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) xq[i].w = q[i];
}
__global__ void k_get_q(int N, const double4* __restrict__ xq, double* __restrict__ q) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) q[i] = xq[i].w;
}
__global__ void k_get_xyz(int N, const double4* __restrict__ xq, double* __restrict__ x) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double4 p = xq[i];
  x[3 * i] = p.x; x[3 * i + 1] = p.y; x[3 * i + 2] = p.z;
}

// fp32 shadow of the positions + (xb != null) the largest squared displacement since the neighbour build.  Squared
// distances are non-negative doubles, whose bit patterns order like unsigned integers: one atomicMax per warp.
// One thread per sorted position k (atom s2a[k]): the S-order copies xs (fp32 shadow) / xqs (exact x,y,z,q) are written
// coalesced, the atom-order shadow xf is scattered; xb = S-order positions at the last build.
__global__ void k_shadow(int N, const int* __restrict__ s2a, const double4* __restrict__ xq, const int* __restrict__ type,
                         double ox, double oy, double oz, float4* __restrict__ xf, float4* __restrict__ xs,
                         double4* __restrict__ xqs, const double4* __restrict__ xb, double* __restrict__ disp2) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (k < N) {
    const int i = s2a[k];
    const double4 p = xq[i];
    const float4 sh = make_float4((float)(p.x - ox), (float)(p.y - oy), (float)(p.z - oz), __int_as_float(type[i]));
    xf[i] = sh; xs[k] = sh; xqs[k] = p;
    if (xb) {
      const double4 b = xb[k];
      const double dx = p.x - b.x, dy = p.y - b.y, dz = p.z - b.z;
      d2 = dx * dx + dy * dy + dz * dz;
      if (!(d2 >= 0.0)) d2 = 1e300;   // NaN positions: never trust the inner block
    }
  }
  if (xb) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
    if ((threadIdx.x & 31) == 0 && d2 > 0.0)
      atomicMax(reinterpret_cast<unsigned long long*>(disp2), (unsigned long long)__double_as_longlong(d2));
  }
}

// S-space maps of a fresh cell sort: inverse permutation, element per sorted position, "is a local atom" flags
__global__ void k_sorted_maps(int N, int n, const int* __restrict__ s2a, const int* __restrict__ type,
                              const int* __restrict__ ltype, int* __restrict__ a2s, int* __restrict__ type_s,
                              int* __restrict__ ltype_s, long long* __restrict__ row_flag) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > N) return;
  if (k == N) { row_flag[N] = 0; return; }
  const int i = s2a[k];
  a2s[i] = k;
  type_s[k] = type[i];
  ltype_s[k] = ltype[i];
  row_flag[k] = i < n ? 1 : 0;
}
__global__ void k_sorted_rows(int N, const int* __restrict__ s2a, const long long* __restrict__ row_flag,
                              const long long* __restrict__ row_scan, int* __restrict__ rowpos, int* __restrict__ row_atom) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N || !row_flag[k]) return;
  const long long r = row_scan[k];
  rowpos[r] = k;
  row_atom[r] = s2a[k];
}
__global__ void k_sorted_ghosts(int n, int nghost, const int* __restrict__ a2s, const int* __restrict__ owner,
                                int* __restrict__ gs_pos, int* __restrict__ gs_own) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  gs_pos[g] = a2s[n + g];
  const int o = owner ? owner[g] : -1;       // multi-GPU runs: ghosts are filled by the boundary exchange instead
  gs_own[g] = o >= 0 ? a2s[o] : -1;
}

// Periodic images of each local row, as a CSR list of sorted positions: lets the CG sweep store a row's new search-direction
// value into its images itself (single-rank forward_comm_fix fused into the producer; no separate ghost-copy launch)
__global__ void k_row_of_atom(int n, const int* __restrict__ row_atom, int* __restrict__ row_of_atom) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) row_of_atom[row_atom[r]] = r;
}
__global__ void k_img_count(int nghost, int n, const int* __restrict__ owner, const int* __restrict__ row_of_atom, int* __restrict__ cnt) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  const int o = owner[g];
  if (o >= 0 && o < n) atomicAdd(&cnt[row_of_atom[o]], 1);
}
__global__ void k_img_fill(int nghost, int n, const int* __restrict__ owner, const int* __restrict__ row_of_atom,
                           const int* __restrict__ off, int* __restrict__ cursor, const int* __restrict__ gs_pos, int* __restrict__ img_pos) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  const int o = owner[g];
  if (o < 0 || o >= n) return;
  const int r = row_of_atom[o];
  img_pos[off[r] + atomicAdd(&cursor[r], 1)] = gs_pos[g];
}

// virial_fdotr over all atoms, pair_reaxc_sunway.cpp:674-702
__global__ void k_fdotr(int N, const double4* __restrict__ xq, const double* __restrict__ f, double* __restrict__ virial) {
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const double4 p = xq[i];
    const double fx = f[3 * i], fy = f[3 * i + 1], fz = f[3 * i + 2];
    v[0] += fx * p.x; v[1] += fy * p.y; v[2] += fz * p.z;
    v[3] += fy * p.x; v[4] += fz * p.x; v[5] += fz * p.y;
  }
  for (int k = 0; k < 6; k++) {
    const double s = warp_sum(v[k]);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(&virial[k], s);
  }
}

// ---- mini LAMMPS core on the device ----
struct BoxD { double h[6], h_inv[6]; };

__device__ __forceinline__ void x2lamda(const BoxD& b, double x, double y, double z, double* l) {
  l[0] = b.h_inv[0] * x + b.h_inv[5] * y + b.h_inv[4] * z;
  l[1] = b.h_inv[1] * y + b.h_inv[3] * z;
  l[2] = b.h_inv[2] * z;
}
__device__ __forceinline__ void shift_vec(const BoxD& b, int sx, int sy, int sz, double* d) {
  d[0] = sx * b.h[0] + sy * b.h[5] + sz * b.h[4];
  d[1] = sy * b.h[1] + sz * b.h[3];
  d[2] = sz * b.h[2];
}

__global__ void k_remap(int n, BoxD b, double4* __restrict__ xq) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = xq[i];
  double l[3];
  x2lamda(b, p.x, p.y, p.z, l);
  const int s0 = (int)floor(l[0]), s1 = (int)floor(l[1]), s2 = (int)floor(l[2]);
  if (s0 | s1 | s2) {
    double d[3];
    shift_vec(b, s0, s1, s2, d);
    p.x -= d[0]; p.y -= d[1]; p.z -= d[2];
    xq[i] = p;
  }
}

// periodic-image ghosts out to cutghost (LAMMPS Comm::borders semantics, single rank)
template <bool FILL>
__global__ void k_ghosts(int n, BoxD b, double cg0, double cg1, double cg2, int m0, int m1, int m2,
                         double4* __restrict__ xq, int* __restrict__ type, int* __restrict__ tag, int* __restrict__ ltype,
                         long long* __restrict__ count, const long long* __restrict__ off, int* __restrict__ owner,
                         int* __restrict__ shift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = xq[i];
  double l[3];
  x2lamda(b, p.x, p.y, p.z, l);
  int c = 0;
  long long w = FILL ? off[i] : 0;
  for (int sz = -m2; sz <= m2; sz++) {
    const double l2 = l[2] + sz;
    if (!(l2 >= -cg2 && l2 < 1.0 + cg2)) continue;
    for (int sy = -m1; sy <= m1; sy++) {
      const double l1 = l[1] + sy;
      if (!(l1 >= -cg1 && l1 < 1.0 + cg1)) continue;
      for (int sx = -m0; sx <= m0; sx++) {
        if (!sx && !sy && !sz) continue;
        const double l0 = l[0] + sx;
        if (!(l0 >= -cg0 && l0 < 1.0 + cg0)) continue;
        if (FILL) {
          double d[3];
          shift_vec(b, sx, sy, sz, d);
          const long long g = n + w;
          xq[g] = make_double4(p.x + d[0], p.y + d[1], p.z + d[2], p.w);
          type[g] = type[i]; tag[g] = tag[i]; ltype[g] = ltype[i];
          owner[w] = i;
          shift[3 * w] = sx; shift[3 * w + 1] = sy; shift[3 * w + 2] = sz;
          w++;
        } else {
          c++;
        }
      }
    }
  }
  if (!FILL) count[i] = c;
}

__global__ void k_forward_x(int n, int nghost, BoxD b, const int* __restrict__ owner, const int* __restrict__ shift,
                            double4* __restrict__ xq) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  const int o = owner[g];
  if (o < 0) return;
  double d[3];
  shift_vec(b, shift[3 * g], shift[3 * g + 1], shift[3 * g + 2], d);
  const double4 p = xq[o];
  xq[n + g] = make_double4(p.x + d[0], p.y + d[1], p.z + d[2], p.w);
}

__global__ void k_reverse_f(int n, int nghost, const int* __restrict__ owner, double* __restrict__ f) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  const int o = owner[g];
  if (o < 0) return;
  const double fx = f[3 * (n + g)], fy = f[3 * (n + g) + 1], fz = f[3 * (n + g) + 2];
  if (fx != 0.0) atomicAdd(&f[3 * o], fx);
  if (fy != 0.0) atomicAdd(&f[3 * o + 1], fy);
  if (fz != 0.0) atomicAdd(&f[3 * o + 2], fz);
}

// fix nve, fix_nve_sw64.c:43-99 (per-type mass branch)
__global__ void k_nve_initial(int n, double dtf, double dtv, const int* __restrict__ ltype, const double* __restrict__ mass,
                              const double* __restrict__ f, double* __restrict__ vel, double4* __restrict__ xq) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double dtfm = dtf / mass[ltype[i]];
  double4 p = xq[i];
  double vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
  vx += dtfm * f[3 * i]; vy += dtfm * f[3 * i + 1]; vz += dtfm * f[3 * i + 2];
  p.x += dtv * vx; p.y += dtv * vy; p.z += dtv * vz;
  vel[3 * i] = vx; vel[3 * i + 1] = vy; vel[3 * i + 2] = vz;
  xq[i] = p;
}
__global__ void k_nve_final(int n, double dtf, const int* __restrict__ ltype, const double* __restrict__ mass,
                            const double* __restrict__ f, double* __restrict__ vel) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double dtfm = dtf / mass[ltype[i]];
  vel[3 * i] += dtfm * f[3 * i]; vel[3 * i + 1] += dtfm * f[3 * i + 1]; vel[3 * i + 2] += dtfm * f[3 * i + 2];
}
__global__ void k_kinetic(int n, const int* __restrict__ ltype, const double* __restrict__ mass, const double* __restrict__ vel,
                          double* __restrict__ out) {
  double ke = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    ke += mass[ltype[i]] * (vel[3 * i] * vel[3 * i] + vel[3 * i + 1] * vel[3 * i + 1] + vel[3 * i + 2] * vel[3 * i + 2]);
  ke = warp_sum(ke);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, ke);
}

inline int nblk(long n, int t = 256) { return (int)((n + t - 1) / t); }

}  // namespace

void Box::set(double xprd, double yprd, double zprd, double xy, double xz, double yz) {
  h[0] = xprd; h[1] = yprd; h[2] = zprd; h[3] = yz; h[4] = xz; h[5] = xy;
  h_inv[0] = 1.0 / h[0]; h_inv[1] = 1.0 / h[1]; h_inv[2] = 1.0 / h[2];
  h_inv[3] = -h[3] / (h[1] * h[2]);
  h_inv[4] = (h[3] * h[5] - h[1] * h[4]) / (h[0] * h[1] * h[2]);
  h_inv[5] = -h[5] / (h[0] * h[1]);
}

System::System(int device) : device_(device) {
  RXB_CUDA(cudaSetDevice(device));
  RXB_CUDA(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;  // numerically lower = higher priority
    RXB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    // Where the bonded chain runs (RXB_CHAIN_MODE, A/B in profiles/r02_chain_mode_ab.txt): 2 (default) = from the start of
    // the step on a LOW-priority stream, so it fills what the bandwidth-bound CG solve leaves idle instead of displacing
    // SpMV CTAs; 0 = the same on a high-priority stream (round 1: 0.2 ms/step slower); 1 = after the CG solve, beside the
    // nonbonded kernel (no gain).
    chain_mode_ = getenv("RXB_CHAIN_MODE") ? atoi(getenv("RXB_CHAIN_MODE")) : 2;
    RXB_CUDA(cudaStreamCreateWithPriority(&st2_, cudaStreamNonBlocking, chain_mode_ == 2 ? lo : hi));
  }
  RXB_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
  RXB_CUDA(cudaEventCreateWithFlags(&ev_far_, cudaEventDisableTiming));
  RXB_CUDA(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
  b_cursor.resize(1); overflow.resize(1); en_d.resize(E_NUM); virial_d.resize(6); need_row_d.resize(2);
  RXB_CUDA(cudaMemset(overflow.p, 0, sizeof(int)));
  RXB_CUDA(cudaMemset(need_row_d.p, 0, 2 * sizeof(int)));
  spmv_active_d.resize(1);
  RXB_CUDA(cudaMemset(spmv_active_d.p, 0, sizeof(unsigned long long)));
}

System::~System() {
  cudaSetDevice(device_);
  dist_destroy();
  if (h_pin_) cudaFreeHost(h_pin_);
  for (cudaEvent_t e : ev_pool_) cudaEventDestroy(e);
  if (st_) cudaStreamDestroy(st_);
  if (st2_) cudaStreamDestroy(st2_);
  if (ev_fork_) { cudaEventDestroy(ev_fork_); cudaEventDestroy(ev_far_); cudaEventDestroy(ev_join_); }
}

double* System::pin(size_t doubles) {
  if (doubles > h_pin_cap_) {
    if (h_pin_) cudaFreeHost(h_pin_);
    h_pin_cap_ = doubles + doubles / 4 + 1024;
    RXB_CUDA(cudaMallocHost(&h_pin_, h_pin_cap_ * sizeof(double)));
  }
  return h_pin_;
}

// NVTX ranges (RXB_NVTX=1) named after the reference's GPTL regions, so that an nsys / ncu timeline of this path reads like
// the reference's own profile (gptl.h regions in pair_reaxc_sunway.cpp, reaxc_forces_sunway.cpp, fix_qeq_reax_sunway.cpp)
static const char* const kNvtxNames[StepTimers::NUM] = {
    "full_build_sunway", "compute H Full c", "CG (sparse matvec in for)", "reaxc init forces noqeq", "reaxc bond orders c",
    "reaxc bonded forces", "reaxc vdw coul full c", "reaxc total force", "slave sparse matvec c", "reaxc hydrogen bonds c",
    "reaxc valence angles c + torsion angles c", "reaxc atom energy and bonds c", "enumerate bonded work lists",
    "slave sparse matvec c (boundary rows)"};
static const bool g_nvtx = getenv("RXB_NVTX") && atoi(getenv("RXB_NVTX")) != 0;

int System::tick(int which, cudaStream_t st) {
  if (g_nvtx) nvtxRangePushA(kNvtxNames[which]);
  if (!profile) return g_nvtx ? -2 : -1;
  if (!st) st = st_;
  if (ev_used_ + 2 > ev_pool_.size()) {
    for (int k = 0; k < 64; k++) { cudaEvent_t e; RXB_CUDA(cudaEventCreate(&e)); ev_pool_.push_back(e); }
  }
  const int a = (int)ev_used_++, b = (int)ev_used_++;
  RXB_CUDA(cudaEventRecord(ev_pool_[a], st));
  ev_pending_.push_back(Pending{which, a, b});
  return (int)ev_pending_.size() - 1;
}
void System::tock(int id, cudaStream_t st) {
  if (g_nvtx) nvtxRangePop();
  if (id < 0) return;
  RXB_CUDA(cudaEventRecord(ev_pool_[ev_pending_[id].b], st ? st : st_));
}
void System::resolve_timers() {
  if (ev_pending_.empty()) return;
  RXB_SYNC(st_);
  for (const Pending& p : ev_pending_) {
    float ms = 0;
    RXB_CUDA(cudaEventElapsedTime(&ms, ev_pool_[p.a], ev_pool_[p.b]));
    timers.ms[p.which] += ms;
    timers.calls[p.which]++;
  }
  ev_pending_.clear();
  ev_used_ = 0;
}

void System::upload_params() {
  RXB_CUDA(cudaSetDevice(device_));
  ff.derive();
  const int nt = ff.nt;
  size_t o_gp = 0;
  size_t o_atom = o_gp + ((ff.gp.size() * sizeof(double) + 31) / 32) * 32;
  size_t o_pair = o_atom + ((ff.atom.size() * sizeof(AtomPar) + 31) / 32) * 32;
  size_t o_angle = o_pair + ((ff.pair.size() * sizeof(PairPar) + 31) / 32) * 32;
  size_t o_tors = o_angle + ((ff.angle.size() * sizeof(AngleSet) + 31) / 32) * 32;
  size_t o_hb = o_tors + ((ff.tors.size() * sizeof(TorsPar) + 31) / 32) * 32;
  size_t total = o_hb + ff.hb.size() * sizeof(HbPar) + 64;
  std::vector<char> blob(total, 0);
  memcpy(blob.data() + o_gp, ff.gp.data(), ff.gp.size() * sizeof(double));
  memcpy(blob.data() + o_atom, ff.atom.data(), ff.atom.size() * sizeof(AtomPar));
  memcpy(blob.data() + o_pair, ff.pair.data(), ff.pair.size() * sizeof(PairPar));
  memcpy(blob.data() + o_angle, ff.angle.data(), ff.angle.size() * sizeof(AngleSet));
  memcpy(blob.data() + o_tors, ff.tors.data(), ff.tors.size() * sizeof(TorsPar));
  memcpy(blob.data() + o_hb, ff.hb.data(), ff.hb.size() * sizeof(HbPar));
  param_blob_.resize(total);
  RXB_CUDA(cudaMemcpy(param_blob_.p, blob.data(), total, cudaMemcpyHostToDevice));
  dp_.nt = nt;
  dp_.ctl = ff.ctl;
  dp_.gp = reinterpret_cast<const double*>(param_blob_.p + o_gp);
  dp_.atom = reinterpret_cast<const AtomPar*>(param_blob_.p + o_atom);
  dp_.pair = reinterpret_cast<const PairPar*>(param_blob_.p + o_pair);
  dp_.angle = reinterpret_cast<const AngleSet*>(param_blob_.p + o_angle);
  dp_.tors = reinterpret_cast<const TorsPar*>(param_blob_.p + o_tors);
  dp_.hb = reinterpret_cast<const HbPar*>(param_blob_.p + o_hb);
  dp_.lut = nullptr; dp_.lut_n = 0; dp_.lut_dx = dp_.lut_inv_dx = 0.0;
  if (ff.ctl.tabulate > 0) {
    int ln = 0; double ldx = 0;
    const std::vector<double> t = ff.lookup_tables(&ln, &ldx);
    lut_d.resize(t.size() / 4);
    RXB_CUDA(cudaMemcpy(lut_d.p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
    dp_.lut = lut_d.p; dp_.lut_n = ln; dp_.lut_dx = ldx; dp_.lut_inv_dx = ff.ctl.tabulate / ff.ctl.nonb_cut;
  }
  // fix qeq/reax shielding: shld = (gamma_i gamma_j)^-1.5 from Pair::extract("gamma"), fix_qeq_reax_sunway.cpp:440-454
  std::vector<double> sh((size_t)nt * nt);
  for (int i = 0; i < nt; i++)
    for (int j = 0; j < nt; j++) sh[(size_t)i * nt + j] = pow(ff.atom[i].gamma * ff.atom[j].gamma, -1.5);
  shld_d.resize(sh.size());
  RXB_CUDA(cudaMemcpy(shld_d.p, sh.data(), sh.size() * sizeof(double), cudaMemcpyHostToDevice));
}

// Largest distance at which any element pair of the force field can still have BO' >= bo_cut (PairPar::d_bond_max,
// never more than the control file's bond_cut): no bond of BOp_single lies beyond it.
double System::bond_reach() const {
  double r = 0.0;
  for (const PairPar& p : ff.pair) r = std::max(r, p.d_bond_max);
  return (r > 0.0 && r < ff.ctl.bond_cut) ? r : ff.ctl.bond_cut;
}

double System::cutneigh() const {
  const Control& c = ff.ctl;
  const double cutmax = std::max(c.nonb_cut, std::max(c.hbond_cut, 2 * c.bond_cut));  // pair_reaxc_sunway.cpp:410
  return std::max(cutmax, qeq_swb) + skin;
}

void System::ensure_atom_capacity() {
  const size_t NN = N;
  xq.resize_keep(NN); type.resize_keep(NN); tag.resize_keep(NN); ltype_d.resize_keep(NN);
  f.resize(3 * NN); CdDelta.resize(NN);
  b_start.resize(NN); b_cnt.resize(NN);
  total_bop.resize(NN); Deltap.resize(NN); dDeltap_self.resize(3 * NN); total_bo.resize(NN); Delta_boc.resize(NN);
  Delta.resize(NN); Delta_val.resize(NN); vlpex.resize(NN); nlp.resize(NN); Delta_lp.resize(NN); dDelta_lp.resize(NN);
  Delta_lp_temp.resize(NN);
  far_num.resize(n > 0 ? n : 1);
  vt_sbo.resize(n > 0 ? n : 1); sum56.resize(NN > 0 ? NN : 1); it_count.resize(4);
  if (cap_ang < n * 6 + 1024) { cap_ang = n * 6 + 1024; it_ang.resize(cap_ang); }
  if (cap_tor < n * 16 + 1024) { cap_tor = n * 16 + 1024; it_tor.resize(cap_tor); }
  if (cap_hb < n * 20 + 1024) { cap_hb = n * 20 + 1024; it_hb.resize(cap_hb); }
  if (cap_bonds < (int)std::min<size_t>(NN * 14 + 1024, 2000000000)) ensure_bond_capacity((int)std::min<size_t>(NN * 14 + 1024, 2000000000));
}

void System::ensure_bond_capacity(int cap) {
  if (cap <= cap_bonds) return;
  const size_t c = cap;
  b_nbr.resize(c); b_sym.resize(c); b_owner.resize(c); b_geo.resize(c); b_bo.resize(c); b_der.resize(c); b_c1.resize(c); b_c2.resize(c);
  b_c3.resize(c); b_Cdbo.resize(c); b_Cdbopi.resize(c); b_Cdbopi2.resize(c);
  cap_bonds = cap;
}

// true when the CUDA runtime knows `p` as page-locked host memory (cudaHostRegister / cudaMallocHost): such buffers are
// copied directly, pageable ones go through the pinned staging buffer
static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

void System::h2d(void* dst, const void* src, size_t bytes, size_t stage_off_doubles) {
  if (!bytes) return;
  if (is_pinned(src)) {
    RXB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st_));
  } else {
    char* pp = (char*)h_pin_ + stage_off_doubles * sizeof(double);
    memcpy(pp, src, bytes);
    RXB_CUDA(cudaMemcpyAsync(dst, pp, bytes, cudaMemcpyHostToDevice, st_));
  }
}

void System::set_atoms(int nlocal, int nghost, const double* x, const int* ltype, const int* tg, const double* q,
                       const int* owner_in) {
  RXB_CUDA(cudaSetDevice(device_));
  if (chain_inflight_) cancel_inflight();
  positions_changed();
  s2a.n = 0; x_build.n = 0;              // the sorted space and the lists belong to the previous atom set
  n = nlocal; N = nlocal + nghost;
  ensure_atom_capacity();
  std::vector<int> hown;
  if (!owner_in && nghost > 0) {  // single-rank LAMMPS: the ghost's owner is the local atom with the same tag (atom->map)
    hown.assign(nghost, -1);
    std::unordered_map<int, int> by_tag;
    by_tag.reserve(nlocal * 2);
    for (int i = 0; i < nlocal; i++) by_tag.emplace(tg[i], i);
    for (int g = 0; g < nghost; g++) {
      auto it = by_tag.find(tg[nlocal + g]);
      hown[g] = it == by_tag.end() ? -1 : it->second;
    }
    owner_in = hown.data();
  }
  ghost_owner.resize(std::max(nghost, 1));
  map_d.resize(std::max<size_t>(ff.map.size(), 1));
  RXB_CUDA(cudaMemcpyAsync(map_d.p, ff.map.data(), ff.map.size() * sizeof(int), cudaMemcpyHostToDevice, st_));
  // raw arrays up (x 24 B, q 8 B, type/tag 4 B per atom), packed into xq/type on the device
  const size_t NN = N;
  pin(5 * NN + 64);
  x_stage.resize(4 * NN + 16);
  h2d(x_stage.p, x, 3 * NN * sizeof(double), 0);
  if (q) h2d(x_stage.p + 3 * NN, q, NN * sizeof(double), 3 * NN);
  h2d(ltype_d.p, ltype, NN * sizeof(int), 4 * NN);
  h2d(tag.p, tg, NN * sizeof(int), 4 * NN + NN / 2 + 1);
  if (nghost > 0) RXB_CUDA(cudaMemcpyAsync(ghost_owner.p, owner_in, (size_t)nghost * sizeof(int), cudaMemcpyHostToDevice, st_));
  if (N > 0)
    k_pack_atoms<<<nblk(N), 256, 0, st_>>>(N, x_stage.p, q ? x_stage.p + 3 * NN : nullptr, ltype_d.p, map_d.p, (int)ff.map.size(),
                                           xq.p, type.p);
  kernel_launches++;
  RXB_SYNC(st_);
}

// x_host must stay unchanged until the next synchronising call on this handle (rxb_qeq_pre_force / rxb_pair_compute)
// when it is page-locked memory; pageable memory is staged before returning.
void System::set_positions(const double* x_host) {
  RXB_CUDA(cudaSetDevice(device_));
  if (chain_inflight_) cancel_inflight();
  positions_changed();
  pin((size_t)3 * N);
  x_stage.resize((size_t)3 * N);
  h2d(x_stage.p, x_host, (size_t)3 * N * sizeof(double), 0);
  k_set_xyz<<<nblk(N), 256, 0, st_>>>(N, x_stage.p, xq.p);
  kernel_launches++;
}

void System::set_charges(const double* q_host) {
  RXB_CUDA(cudaSetDevice(device_));
  positions_changed();                   // xqs carries q: refresh the S-order copy
  double* pp = pin((size_t)N);
  memcpy(pp, q_host, (size_t)N * sizeof(double));
  x_stage.resize((size_t)3 * N);
  RXB_CUDA(cudaMemcpyAsync(x_stage.p, pp, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, st_));
  k_set_q<<<nblk(N), 256, 0, st_>>>(N, x_stage.p, xq.p);
  RXB_SYNC(st_);
  kernel_launches++;
}

void System::get_forces(double* f_host) {
  RXB_CUDA(cudaSetDevice(device_));
  const size_t bytes = (size_t)3 * N * sizeof(double);
  if (is_pinned(f_host)) {
    RXB_CUDA(cudaMemcpyAsync(f_host, f.p, bytes, cudaMemcpyDeviceToHost, st_));
    RXB_SYNC(st_);
    return;
  }
  double* pp = pin((size_t)3 * N);
  RXB_CUDA(cudaMemcpyAsync(pp, f.p, bytes, cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
  memcpy(f_host, pp, bytes);
}

void System::get_charges(double* q_host) {
  RXB_CUDA(cudaSetDevice(device_));
  x_stage.resize((size_t)3 * N);
  k_get_q<<<nblk(N), 256, 0, st_>>>(N, xq.p, x_stage.p);
  double* pp = pin((size_t)N);
  RXB_CUDA(cudaMemcpyAsync(pp, x_stage.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
  memcpy(q_host, pp, (size_t)N * sizeof(double));
  kernel_launches++;
}

BondedWork System::bonded_work() {
  BondedWork W{};
  W.ang = it_ang.p; W.tor = it_tor.p; W.hb = it_hb.p;
  W.n_ang = it_count.p; W.n_tor = it_count.p + 1; W.n_hb = it_count.p + 2;
  W.cap_ang = cap_ang; W.cap_tor = cap_tor; W.cap_hb = cap_hb;
  W.sbo = vt_sbo.p; W.sum56 = sum56.p;
  return W;
}

void System::update_shadow(cudaStream_t st) {
  if (N == 0 || shadow_valid_) return;        // once per set of positions / charges (reset by positions_changed())
  if (s2a.n != (size_t)N)
    throw std::runtime_error("rxb: no neighbour build for the current atom set (rxb_neigh_build must follow rxb_set_atoms)");
  xf.resize((size_t)N); xs.resize((size_t)N); xqs.resize((size_t)N);
  disp2_d.resize(1);
  // displacement since the build is only meaningful for the atom set the lists were built for
  const bool track = x_build.n == (size_t)N && vl.cut_in > 0.0;
  if (track) RXB_CUDA(cudaMemsetAsync(disp2_d.p, 0, sizeof(double), st));
  else { static const double huge = 1e300; RXB_CUDA(cudaMemcpyAsync(disp2_d.p, &huge, sizeof(double), cudaMemcpyHostToDevice, st)); }
  k_shadow<<<nblk(N), 256, 0, st>>>(N, s2a.p, xq.p, type.p, cells_a_.origin[0], cells_a_.origin[1], cells_a_.origin[2], xf.p,
                                    xs.p, xqs.p, track ? x_build.p : nullptr, disp2_d.p);
  kernel_launches++;
  shadow_valid_ = true;
}

// S space of a fresh cell sort (cells_a_): permutation and its inverse, rows = local atoms in sorted order, ghost maps
void System::build_sorted_space() {
  const size_t NN = std::max(N, 1);
  s2a.resize(NN); a2s.resize(NN); type_s.resize(NN); row_flag.resize(NN + 1); row_scan.resize(NN + 1);
  rowpos.resize(std::max(n, 1)); row_atom.resize(std::max(n, 1));
  s2a.n = (size_t)N;
  if (N == 0) return;
  RXB_CUDA(cudaMemcpyAsync(s2a.p, cells_a_.sorted_idx.p, (size_t)N * sizeof(int), cudaMemcpyDeviceToDevice, st_));
  ltype_s.resize(NN);
  k_sorted_maps<<<nblk(N + 1), 256, 0, st_>>>(N, n, s2a.p, type.p, ltype_d.p, a2s.p, type_s.p, ltype_s.p, row_flag.p);
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, row_flag.p, row_scan.p, N + 1, st_);
  scan_temp.resize(need + 16);
  cub::DeviceScan::ExclusiveSum(scan_temp.p, need, row_flag.p, row_scan.p, N + 1, st_);
  k_sorted_rows<<<nblk(N), 256, 0, st_>>>(N, s2a.p, row_flag.p, row_scan.p, rowpos.p, row_atom.p);
  const int nghost = N - n;
  gs_pos.resize(std::max(nghost, 1)); gs_own.resize(std::max(nghost, 1));
  if (nghost > 0)
    k_sorted_ghosts<<<nblk(nghost), 256, 0, st_>>>(n, nghost, a2s.p, dist_ ? nullptr : ghost_owner.p, gs_pos.p, gs_own.p);
  kernel_launches += 4;
  // image lists per row (single-rank runs; RXB_FUSE_FWD=0 keeps the separate ghost-copy kernel in the CG loop)
  static const bool fuse_fwd = !(getenv("RXB_FUSE_FWD") && atoi(getenv("RXB_FUSE_FWD")) == 0);
  img_valid_ = false;
  if (!dist_ && fuse_fwd && nghost > 0 && n > 0) {
    row_of_atom.resize(n); img_off.resize(n + 1); img_cur.resize(n + 1); img_pos.resize(nghost);
    RXB_CUDA(cudaMemsetAsync(img_cur.p, 0, (size_t)(n + 1) * sizeof(int), st_));
    k_row_of_atom<<<nblk(n), 256, 0, st_>>>(n, row_atom.p, row_of_atom.p);
    k_img_count<<<nblk(nghost), 256, 0, st_>>>(nghost, n, ghost_owner.p, row_of_atom.p, img_cur.p);
    size_t need2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need2, img_cur.p, img_off.p, n + 1, st_);
    scan_temp.resize(need2 + 16);
    cub::DeviceScan::ExclusiveSum(scan_temp.p, need2, img_cur.p, img_off.p, n + 1, st_);
    RXB_CUDA(cudaMemsetAsync(img_cur.p, 0, (size_t)(n + 1) * sizeof(int), st_));
    k_img_fill<<<nblk(nghost), 256, 0, st_>>>(nghost, n, ghost_owner.p, row_of_atom.p, img_off.p, img_cur.p, gs_pos.p, img_pos.p);
    kernel_launches += 4;
    img_valid_ = true;
  }
}

// fix qeq/reax <param file>: chi, eta, gamma per LAMMPS type (index 1..ntypes; ntypes = 0 returns to the pair style's values)
void System::qeq_set_type_params(int ntypes, const double* chi, const double* eta, const double* gamma) {
  RXB_CUDA(cudaSetDevice(device_));
  qeq_chi_lt.clear(); qeq_eta_lt.clear(); qeq_gamma_lt.clear();
  if (ntypes <= 0 || !chi || !eta || !gamma) return;
  const int nlt = ntypes + 1;
  qeq_chi_lt.assign(chi, chi + nlt); qeq_eta_lt.assign(eta, eta + nlt); qeq_gamma_lt.assign(gamma, gamma + nlt);
  std::vector<double> blob((size_t)2 * nlt + (size_t)nlt * nlt, 0.0);
  for (int t = 1; t < nlt; t++) { blob[t] = chi[t]; blob[nlt + t] = eta[t]; }
  for (int i = 1; i < nlt; i++)
    for (int j = 1; j < nlt; j++) blob[2 * nlt + (size_t)i * nlt + j] = pow(gamma[i] * gamma[j], -1.5);   // init_shielding :440-454
  qeq_lt_d.resize(blob.size());
  RXB_CUDA(cudaMemcpy(qeq_lt_d.p, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice));
}

// H entry format for the current settings (see rxb_dev.cuh); (re)allocates the far-list storage of the chosen format
void System::choose_h_format() {
  double shld_min = 1e300;
  for (int i = 0; i < ff.nt; i++)
    for (int j = 0; j < ff.nt; j++) shld_min = std::min(shld_min, pow(ff.atom[i].gamma * ff.atom[j].gamma, -1.5));
  for (size_t i = 1; i < qeq_gamma_lt.size(); i++)           // fix qeq/reax <param file>: its gammas define the bound
    for (size_t j = 1; j < qeq_gamma_lt.size(); j++) shld_min = std::min(shld_min, pow(qeq_gamma_lt[i] * qeq_gamma_lt[j], -1.5));
  const double bound = 14.4 / cbrt(shld_min > 0 ? shld_min : 1e-300);          // Tap in [0,1] when the taper starts at 0
  const bool packed = !h_exact_request && qeq_swa == 0.0 && N < (1 << 22) && bound > 0 && bound < 1e6 && std::isfinite(bound);
  h_packed_ = packed;
  const size_t slots = (size_t)std::max<long long>(vl.slots, 1);
  if (packed) {
    int shift = 0;
    while (shift < 60 && bound * 1.0001 * ldexp(1.0, shift + 1) < ldexp(1.0, kHColShift)) shift++;
    h_quant_ = ldexp(1.0, shift);
    hpk.resize(slots);
  } else {
    h_quant_ = 1.0;
    far_idx.resize(slots); H_val.resize(slots);
  }
}

DevView System::view() {
  DevView v{};
  v.n = n; v.N = N; v.cap_bonds = cap_bonds;
  xf.resize((size_t)(N > 0 ? N : 1));
  v.xf = xf.p;
  {
    const double far = std::max(ff.ctl.nonb_cut, last_swb_ > 0 ? last_swb_ : qeq_swb);
    v.far_band = cells_a_.fp32_band(far);
    v.bond_band = cells_a_.fp32_band(ff.ctl.bond_cut);
  }
  v.xq = xq.p; v.type = type.p; v.tag = tag.p; v.f = f.p; v.CdDelta = CdDelta.p;
  v.vl_off = vl.off.p; v.vl_idx = vl.idx.p; v.vl_cnt = vl.cnt.p; v.vl_stride = vl.stride;
  v.s2a = s2a.p; v.rowpos = rowpos.p; v.row_atom = row_atom.p; v.xs = xs.p; v.xqs = xqs.p; v.type_s = type_s.p;
  v.ltype_s = ltype_s.p;
  const bool par_file = !qeq_gamma_lt.empty();
  v.nlt = par_file ? (int)qeq_gamma_lt.size() : 0;
  v.shld_lt = par_file ? qeq_lt_d.p + 2 * v.nlt : nullptr;
  v.chi_lt = par_file ? qeq_lt_d.p : nullptr;
  v.eta_lt = par_file ? qeq_lt_d.p + v.nlt : nullptr;
  disp2_d.resize(1);
  v.vl_cnt_in = vl.cnt_in.p; v.disp2 = disp2_d.p; v.vl_cut_in = vl.cut_in;
  v.bc_off = bc.off.p; v.bc_idx = bc.idx.p; v.bc_cnt = bc.cnt.p;
  v.hc_off = nullptr; v.hc_idx = nullptr;
  v.far_num = far_num.p; v.far_idx = far_idx.p; v.H_val = H_val.p;
  v.hpk = h_packed_ ? hpk.p : nullptr; v.h_quant = h_quant_;
  v.b_start = b_start.p; v.b_cnt = b_cnt.p; v.b_cursor = b_cursor.p; v.overflow = overflow.p;
  v.row_cap = row_cap_; v.strong_cap = strong_cap_; v.need_row = need_row_d.p;
  v.b_nbr = b_nbr.p; v.b_sym = b_sym.p; v.b_owner = b_owner.p; v.b_geo = b_geo.p; v.b_bo = b_bo.p; v.b_der = b_der.p;
  v.b_c1 = b_c1.p; v.b_c2 = b_c2.p; v.b_c3 = b_c3.p;
  v.b_Cdbo = b_Cdbo.p; v.b_Cdbopi = b_Cdbopi.p; v.b_Cdbopi2 = b_Cdbopi2.p;
  v.total_bop = total_bop.p; v.Deltap = Deltap.p; v.dDeltap_self = dDeltap_self.p; v.total_bo = total_bo.p;
  v.Delta_boc = Delta_boc.p; v.Delta = Delta.p; v.Delta_val = Delta_val.p; v.vlpex = vlpex.p; v.nlp = nlp.p;
  v.Delta_lp = Delta_lp.p; v.dDelta_lp = dDelta_lp.p; v.Delta_lp_temp = Delta_lp_temp.p;
  v.en = en_d.p; v.virial = virial_d.p;
  return v;
}

void System::build_neighbors() {
  RXB_CUDA(cudaSetDevice(device_));
  const int t_NEIGH = tick(StepTimers::NEIGH);
  const double cn = cutneigh();
  // Verlet list for local rows: bins of cn/2, +-2 cells
  // bins of cn / reach, +-reach cells (RXB_VL_REACH, default 2): finer y/z bins make the sorted order - and with it every
  // gather of the long-range kernels - more local and trim the candidate volume, at the price of more, shorter runs
  static const int vl_reach = getenv("RXB_VL_REACH") ? std::max(1, atoi(getenv("RXB_VL_REACH"))) : 2;
  cells_a_.bin(xq.p, N, cn / vl_reach, vl_reach, st_);
  build_sorted_space();
  // rows partitioned at far cut-off + kInnerSkin: between rebuilds the per-step far-list sweep reads only that block while
  // no atom has moved more than kInnerSkin / 2 (checked on the device every step)
  const double far = std::max(ff.ctl.nonb_cut, qeq_swb);
  cells_a_.build(xq.p, n, cn, far + kInnerSkin, vl, st_, rowpos.p);
  // S-order copies of the positions; they are also the reference positions of the displacement tracking (zero now)
  x_build.resize((size_t)std::max(N, 1));
  x_build.n = 0;
  shadow_valid_ = false;
  update_shadow(st_);
  if (N > 0) {
    RXB_CUDA(cudaMemcpyAsync(x_build.p, xqs.p, (size_t)N * sizeof(double4), cudaMemcpyDeviceToDevice, st_));
    RXB_CUDA(cudaMemsetAsync(disp2_d.p, 0, sizeof(double), st_));
  }
  x_build.n = (size_t)N;
  if (dist_) { dist_sorted_maps(); if (!dist_external()) dist_classify_rows(); }
  // bond candidates for all rows (ghosts too): (reach of the longest possible bond <= bond_cut) + skin
  const double cb = bond_reach() + skin;
  cells_b_.bin(xq.p, N, cb / 2.0, 2, st_);
  cells_b_.build(xq.p, N, cb, 0.0, bc, st_);
  choose_h_format();
  kernel_launches += 12;
  tock(t_NEIGH);
}

void System::step_forces(bool eflag, bool vflag) {
  DevView v = view();
  RXB_CUDA(cudaMemsetAsync(f.p, 0, (size_t)3 * N * sizeof(double), st_));
  RXB_CUDA(cudaMemsetAsync(CdDelta.p, 0, (size_t)N * sizeof(double), st_));
  RXB_CUDA(cudaMemsetAsync(en_d.p, 0, E_NUM * sizeof(double), st_));
  RXB_CUDA(cudaMemsetAsync(virial_d.p, 0, 6 * sizeof(double), st_));
  const int t_BONDS = tick(StepTimers::BONDS);
  launch_bond_list(*this, v, dp_, st_);
  tock(t_BONDS);
  const int t_NONB = tick(StepTimers::NONB);
  launch_nonbonded(*this, v, dp_, eflag || vflag, st_);
  tock(t_NONB);
  const int t_BO = tick(StepTimers::BO);
  launch_bond_orders(*this, v, dp_, st_);
  tock(t_BO);
  const int t_BONDED = tick(StepTimers::BONDED);
  launch_bonded(*this, v, dp_, st_);
  tock(t_BONDED);
  const int t_DBOND = tick(StepTimers::DBOND);
  launch_dbond(*this, v, dp_, st_);
  tock(t_DBOND);
  if (vflag) { k_fdotr<<<148 * 4, 256, 0, st_>>>(N, xq.p, f.p, virial_d.p); kernel_launches++; }
}

// End-of-force-phase status: bond cursor, overflow bits and work-list counts of this rank (h, wk) plus, in need_, the
// maxima over all ranks (one small ncclAllReduce(max) in multi-GPU runs), energies / virial on ev steps.  ONE host
// synchronisation.  Every replay decision is taken on need_ / overflow_flag, which are identical on all ranks.
namespace {
// slots: 0 bond cursor, 1..3 n_ang n_tor n_hb, 4 bond arrays too small, 5 fatal per-atom bits, 6..8 work list too small.
// Slots 4..8 are 0/1 (or a bit mask whose any-non-zero matters), so the maximum over ranks is the OR over ranks.
__global__ void k_gather_status(const int* __restrict__ cursor, const int* __restrict__ overflow, const int* __restrict__ counts,
                                const int* __restrict__ need_row, int cap_ang, int cap_tor, int cap_hb,
                                const int* __restrict__ peer_err, int* __restrict__ out) {
  if (threadIdx.x == 0) {
    const int ov = overflow[0];
    out[0] = cursor[0]; out[1] = counts[0]; out[2] = counts[1]; out[3] = counts[2];
    out[4] = (ov & 2) ? 1 : 0; out[5] = ov & ~2;
    out[6] = counts[0] > cap_ang; out[7] = counts[1] > cap_tor; out[8] = counts[2] > cap_hb;
    out[9] = need_row[0]; out[10] = need_row[1];       // longest bond row / strong list that did not fit its staging
    out[11] = peer_err ? peer_err[0] : 0;              // a wait of the peer exchange timed out (a peer rank is gone)
    for (int k = 0; k < 11; k++) out[16 + k] = out[k];
  }
}
}  // namespace

void System::read_step_status(bool ev, int* h, int* wk) {
  status_d_.resize(32);
  k_gather_status<<<1, 32, 0, st_>>>(b_cursor.p, overflow.p, it_count.p, need_row_d.p, cap_ang, cap_tor, cap_hb,
                                     dist_ ? dist_peer_err_ptr() : nullptr, status_d_.p);
  kernel_launches++;
  if (dist_) dist_allreduce_max_int(status_d_.p + 16, 11);
  int host[32];
  RXB_CUDA(cudaMemcpyAsync(host, status_d_.p, 32 * sizeof(int), cudaMemcpyDeviceToHost, st_));
  // (host-planned halo: every rank reports its own partial sums, the host reduces them like the reference's MPI ranks)
  if (ev && dist_ && !dist_external()) { dist_allreduce(en_d.p, E_NUM); dist_allreduce(virial_d.p, 6); }
  if (ev) {
    RXB_CUDA(cudaMemcpyAsync(energies, en_d.p, E_NUM * sizeof(double), cudaMemcpyDeviceToHost, st_));
    RXB_CUDA(cudaMemcpyAsync(virial, virial_d.p, 6 * sizeof(double), cudaMemcpyDeviceToHost, st_));
  }
  RXB_SYNC(st_);
  if (host[11]) throw std::runtime_error("rxb dist: peer-memory exchange timed out (a peer rank stopped responding)");
  h[0] = host[0]; h[1] = host[5] | (host[4] ? 2 : 0);
  wk[0] = host[1]; wk[1] = host[2]; wk[2] = host[3]; wk[3] = 0;
  for (int k = 0; k < 11; k++) need_[k] = host[16 + k];
  overflow_flag = need_[5] | (need_[4] ? 2 : 0);      // over all ranks
}

void System::compute(bool eflag, bool vflag) {
  RXB_CUDA(cudaSetDevice(device_));
  update_shadow(st_);
  if (!qeq_ran_this_step_) {
    // pair style without fix qeq/reax this step (checkqeq no): the far list is still needed
    choose_h_format();
    DevView v = view();
    double Tap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    launch_far_and_H(*this, v, dp_, Tap, shld_d.p, 0.0, st_);
    memcpy(last_tap_, Tap, sizeof(last_tap_)); last_swb_ = 0.0;
  }
  qeq_ran_this_step_ = false;
  for (int attempt = 0; attempt < 6; attempt++) {
    step_forces(eflag, vflag);
    int h[2], wk[4];
    read_step_status(eflag || vflag, h, wk);
    num_bonds = h[0];
    num_ang = wk[0]; num_tor = wk[1]; num_hb = wk[2];
    // a QEq solve enqueued without polling that had not converged yet is continued now: new charges, replay
    const bool q_changed = qeq_settle();
    // (multi-GPU: overflow_flag and the needed capacities below are the maxima over all ranks, so every rank replays the
    // same number of times and the collectives inside the loop stay matched)
    const bool lists_fit = !(need_[6] | need_[7] | need_[8]);
    const bool staging_fits = !(overflow_flag & (1 | 8));
    if (!(overflow_flag & 2) && lists_fit && staging_fits && !q_changed) { overflow_flag = 0; break; }
    if (attempt == 5) { overflow_flag |= 64; break; }           // still not fitting after five replays: reported below
    if (!staging_fits) grow_staging();
    // a list did not fit (on some rank): grow to the largest need of any rank and replay the force computation of this
    // step (positions are unchanged)
    RXB_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), st_));
    h[0] = need_[0]; wk[0] = need_[1]; wk[1] = need_[2]; wk[2] = need_[3];
    if (overflow_flag & 2) ensure_bond_capacity((int)std::min<long long>((long long)h[0] + h[0] / 4 + 1024, 2000000000LL));
    if (wk[0] > cap_ang) { cap_ang = wk[0] + wk[0] / 4 + 1024; it_ang.resize(cap_ang); }
    if (wk[1] > cap_tor) { cap_tor = wk[1] + wk[1] / 4 + 1024; it_tor.resize(cap_tor); }
    if (wk[2] > cap_hb) {   // the candidate list is produced by K-farH: grow it and rebuild the far list of this step
      cap_hb = wk[2] + wk[2] / 4 + 1024; it_hb.resize(cap_hb);
      DevView v2 = view();
      launch_far_and_H(*this, v2, dp_, last_tap_, shld_d.p, last_swb_, st_);
    }
    overflow_flag &= ~2;
  }
  if (overflow_flag)
    throw std::runtime_error("rxb: the force phase did not fit its lists after 6 grow-and-replay attempts (overflow bits " +
                             std::to_string(overflow_flag) + ")");
}

// A bond row / strong-bond list outgrew its shared-memory staging (overflow bits 1 / 8): size it for the largest need seen
// on any rank (+ slack, doubled at least) before the replay.  Bounded by the 227 KB of shared memory a CTA can have
// (~600 bonds per atom, an order of magnitude beyond any physical density).
void System::grow_staging() {
  if (overflow_flag & 1) {
    row_cap_ = std::max(2 * row_cap_, (need_[9] + 8 + 31) / 32 * 32);
    if (row_cap_ > 576) throw std::runtime_error("rxb: more than 576 bonds on one atom (" + std::to_string(need_[9]) + "): unphysical input");
  }
  if (overflow_flag & 8) {
    strong_cap_ = std::max(2 * strong_cap_, need_[10] + 8);
    if (strong_cap_ > 1536) throw std::runtime_error("rxb: more than 1536 strong bonds on one atom: unphysical input");
  }
  RXB_CUDA(cudaMemsetAsync(need_row_d.p, 0, 2 * sizeof(int), st_));
}

// ---------------------------------------------------------------------------------------------------------------
void System::md_setup(const double* box6, int nlocal, const double* x, const double* v, const int* ltype, const int* tg,
                      const double* mass_by_type, int ntypes, double dt, int every) {
  RXB_CUDA(cudaSetDevice(device_));
  if (dist_external()) throw std::runtime_error("rxb_md_setup: a handle in host-planned halo mode (rxb_comm_init) is driven through the plugin calls");
  box.set(box6[0], box6[1], box6[2], box6[3], box6[4], box6[5]);
  md_dt = dt; md_every = every; md_ago = 0; ntimestep = 0;
  std::vector<double> q0(nlocal, 0.0);
  set_atoms(nlocal, 0, x, ltype, tg, q0.data(), nullptr);
  v_d.resize((size_t)3 * nlocal);
  RXB_CUDA(cudaMemcpy(v_d.p, v, (size_t)3 * nlocal * sizeof(double), cudaMemcpyHostToDevice));
  mass_d.resize(ntypes + 1);
  RXB_CUDA(cudaMemcpy(mass_d.p, mass_by_type, (size_t)(ntypes + 1) * sizeof(double), cudaMemcpyHostToDevice));
  qeq_reset_history();
  if (dist_) dist_exchange(); else md_make_ghosts();
  build_neighbors();
  md_force();
}

void System::md_make_ghosts() {
  BoxD b;
  memcpy(b.h, box.h, sizeof(b.h));
  memcpy(b.h_inv, box.h_inv, sizeof(b.h_inv));
  const double cut = cutneigh();
  const double cg0 = cut * sqrt(b.h_inv[0] * b.h_inv[0] + b.h_inv[5] * b.h_inv[5] + b.h_inv[4] * b.h_inv[4]);
  const double cg1 = cut * sqrt(b.h_inv[1] * b.h_inv[1] + b.h_inv[3] * b.h_inv[3]);
  const double cg2 = cut * b.h_inv[2];
  const int m0 = (int)ceil(cg0), m1 = (int)ceil(cg1), m2 = (int)ceil(cg2);
  k_remap<<<nblk(n), 256, 0, st_>>>(n, b, xq.p);
  gcount.resize(n + 1); goff.resize(n + 1);
  k_ghosts<false><<<nblk(n), 256, 0, st_>>>(n, b, cg0, cg1, cg2, m0, m1, m2, xq.p, type.p, tag.p, ltype_d.p, gcount.p, nullptr,
                                           nullptr, nullptr);
  RXB_CUDA(cudaMemsetAsync(gcount.p + n, 0, sizeof(long long), st_));
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, gcount.p, goff.p, n + 1, st_);
  scan_temp.resize(need + 16);
  cub::DeviceScan::ExclusiveSum(scan_temp.p, need, gcount.p, goff.p, n + 1, st_);
  long long nghost = 0;
  RXB_CUDA(cudaMemcpyAsync(&nghost, goff.p + n, sizeof(long long), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
  N = n + (int)nghost;
  ensure_atom_capacity();
  ghost_owner.resize(std::max<size_t>(nghost, 1));
  ghost_shift.resize(std::max<size_t>(3 * nghost, 3));
  k_ghosts<true><<<nblk(n), 256, 0, st_>>>(n, b, cg0, cg1, cg2, m0, m1, m2, xq.p, type.p, tag.p, ltype_d.p, nullptr, goff.p,
                                          ghost_owner.p, ghost_shift.p);
  kernel_launches += 5;
  RXB_CUDA(cudaGetLastError());
}

// Two-stream step: the bond list -> bond orders -> multi-body -> angle/torsion/hbond chain does not
// depend on this step's charges, so it runs on st2_ while the latency/HBM-bound CG solve runs on st_.  Only the hydrogen
// bond enumeration needs this step's far list (event after K-farH); K-nb needs q; K-dbond joins both.
// RXB_OVERLAP_DEBUG=1: time the two streams of the overlapped step with CUDA events (development aid)
namespace {
struct OverlapDbg {
  bool on = getenv("RXB_OVERLAP_DEBUG") != nullptr;
  cudaEvent_t e[8] = {};
  double acc[8] = {};
  long steps = 0;
  void init() { if (!e[0]) for (auto& x : e) cudaEventCreate(&x); }
} g_odbg;
}  // namespace

void System::after_far_hook() {
  if (!hook_after_far_) return;
  hook_after_far_ = false;
  DevView v = view();
  RXB_CUDA(cudaEventRecord(ev_far_, st_));
  if (g_odbg.on) cudaEventRecord(g_odbg.e[2], st_);    // end of K-farH
  RXB_CUDA(cudaStreamWaitEvent(st2_, ev_far_, 0));
  launch_bonded_part2(*this, v, dp_, st2_);
  RXB_CUDA(cudaEventRecord(ev_join_, st2_));
  if (g_odbg.on) cudaEventRecord(g_odbg.e[4], st2_);   // end of the bonded chain
}

// The step's force evaluation in two halves so that the bond list -> bond order -> bonded chain (second stream) overlaps
// the bandwidth-bound QEq solve.  The resident run calls both back to back; the plugin path calls the front half from
// fix qeq/reax's pre_force hook and the back half from the pair style's compute hook.
void System::overlapped_front(bool wait_for_convergence) {
  if (chain_inflight_) cancel_inflight();
  DevView v = view();
  if (g_odbg.on) { g_odbg.init(); cudaEventRecord(g_odbg.e[0], st_); }
  update_shadow(st_);
  RXB_CUDA(cudaMemsetAsync(f.p, 0, (size_t)3 * N * sizeof(double), st_));
  RXB_CUDA(cudaMemsetAsync(CdDelta.p, 0, (size_t)N * sizeof(double), st_));
  RXB_CUDA(cudaMemsetAsync(en_d.p, 0, E_NUM * sizeof(double), st_));
  RXB_CUDA(cudaMemsetAsync(virial_d.p, 0, 6 * sizeof(double), st_));
  if (chain_mode_ != 1) {
    RXB_CUDA(cudaEventRecord(ev_fork_, st_));
    RXB_CUDA(cudaStreamWaitEvent(st2_, ev_fork_, 0));
    launch_bond_list(*this, v, dp_, st2_);
    launch_bond_orders(*this, v, dp_, st2_);
    launch_bonded_part1(*this, v, dp_, st2_);
    if (g_odbg.on) cudaEventRecord(g_odbg.e[1], st2_);   // end of bond list + BO + multi
    hook_after_far_ = true;
  }
  chain_inflight_ = true;
  qeq_pre_force(wait_for_convergence);   // K-farH, then after_far_hook() enqueues the rest of the chain on st2_, then CG
  if (g_odbg.on) cudaEventRecord(g_odbg.e[3], st_);    // end of CG
}

void System::cancel_inflight() {         // the atoms change before the pair style consumed the chain: just join the streams
  RXB_CUDA(cudaStreamWaitEvent(st_, ev_join_, 0));
  chain_inflight_ = false;
}

void System::overlapped_back(bool eflag, bool vflag) {
  const bool ev = eflag || vflag;
  DevView v = view();
  chain_inflight_ = false;
  qeq_ran_this_step_ = false;
  update_shadow(st_);                    // no-op unless rxb_set_charges replaced the charges after the QEq hook (xqs carries q)
  if (chain_mode_ == 1) {                // the whole chain beside the nonbonded kernel instead of beside the CG solve
    RXB_CUDA(cudaEventRecord(ev_fork_, st_));
    RXB_CUDA(cudaStreamWaitEvent(st2_, ev_fork_, 0));
    launch_bond_list(*this, v, dp_, st2_);
    launch_bond_orders(*this, v, dp_, st2_);
    launch_bonded_part1(*this, v, dp_, st2_);
    launch_bonded_part2(*this, v, dp_, st2_);
    RXB_CUDA(cudaEventRecord(ev_join_, st2_));
  }
  launch_nonbonded(*this, v, dp_, ev, st_);
  if (g_odbg.on) cudaEventRecord(g_odbg.e[5], st_);    // end of nonbonded
  RXB_CUDA(cudaStreamWaitEvent(st_, ev_join_, 0));
  launch_dbond(*this, v, dp_, st_);
  if (g_odbg.on) cudaEventRecord(g_odbg.e[6], st_);    // end of dbond
  if (vflag) { k_fdotr<<<148 * 4, 256, 0, st_>>>(N, xq.p, f.p, virial_d.p); kernel_launches++; }
  int h[2], wk[4];
  read_step_status(ev, h, wk);
  if (g_odbg.on) {
    cudaStreamSynchronize(st2_);
    float ms;
    for (int k = 1; k <= 6; k++) { cudaEventElapsedTime(&ms, g_odbg.e[0], g_odbg.e[k]); g_odbg.acc[k] += ms; }
    if (++g_odbg.steps % 20 == 0) {
      const double s = (double)g_odbg.steps;
      fprintf(stderr, "overlap dbg (ms after step start): chain1 %.3f  farH %.3f  CG %.3f  chain2 %.3f  nonbonded %.3f  dbond %.3f\n",
              g_odbg.acc[1] / s, g_odbg.acc[2] / s, g_odbg.acc[3] / s, g_odbg.acc[4] / s, g_odbg.acc[5] / s, g_odbg.acc[6] / s);
      for (double& a : g_odbg.acc) a = 0.0;
      g_odbg.steps = 0;
    }
  }
  num_bonds = h[0];
  num_ang = wk[0]; num_tor = wk[1]; num_hb = wk[2];
  const bool q_changed = qeq_settle();   // a solve enqueued without polling that had to be continued: new charges
  if (q_changed || overflow_flag || need_[6] || need_[7] || need_[8]) {
    qeq_ran_this_step_ = true;           // far list and charges of this step are valid: replay only the force phase
    compute(eflag, vflag);               // sequential path grows the arrays and replays
  }
}

void System::md_force_overlapped(bool ev) {
  overlapped_front(getenv("RXB_QEQ_WAIT") != nullptr);
  overlapped_back(ev, ev);
}

// plugin path (C ABI): fix qeq/reax pre_force, then pair compute
void System::plugin_qeq_pre_force(bool wait_for_convergence) {
  if (overlap && !profile && (!dist_ || dist_external()) && n > 0) overlapped_front(wait_for_convergence);
  else { if (chain_inflight_) cancel_inflight(); qeq_pre_force(wait_for_convergence); }
}
void System::plugin_compute(bool eflag, bool vflag) {
  if (chain_inflight_) overlapped_back(eflag, vflag);
  else compute(eflag, vflag);
}

void System::md_force() {
  const bool ev = md_thermo > 0 && (ntimestep % md_thermo == 0);
  if (overlap && qeq_on && !profile) {
    md_force_overlapped(ev);
  } else {
    if (qeq_on) qeq_pre_force(getenv("RXB_QEQ_WAIT") != nullptr);
    compute(ev, ev);
  }
  const int nghost = N - n;
  if (dist_) dist_reverse_f();
  else if (nghost > 0) { k_reverse_f<<<nblk(nghost), 256, 0, st_>>>(n, nghost, ghost_owner.p, f.p); kernel_launches++; }
}

void System::md_run(int nsteps) {
  RXB_CUDA(cudaSetDevice(device_));
  BoxD b;
  memcpy(b.h, box.h, sizeof(b.h));
  memcpy(b.h_inv, box.h_inv, sizeof(b.h_inv));
  const double dtv = md_dt, dtf = 0.5 * md_dt * kFtm2v;
  if (!run_ev_[0]) { RXB_CUDA(cudaEventCreate(&run_ev_[0])); RXB_CUDA(cudaEventCreate(&run_ev_[1])); }
  RXB_CUDA(cudaEventRecord(run_ev_[0], st_));
  for (int s = 0; s < nsteps; s++) {
    ntimestep++;
    k_nve_initial<<<nblk(n), 256, 0, st_>>>(n, dtf, dtv, ltype_d.p, mass_d.p, f.p, v_d.p, xq.p);
    positions_changed();
    if (species.on && species_step(ntimestep))   // post_integrate: reads the bond list of the previous force evaluation
      species_log.push_back({ntimestep, species.nmole, species.composition});
    md_ago++;
    if (md_ago % md_every == 0) {
      if (dist_) dist_exchange(); else md_make_ghosts();
      build_neighbors();
      md_ago = 0;
    } else if (dist_) {
      dist_forward_xq();
    } else if (N > n) {
      k_forward_x<<<nblk(N - n), 256, 0, st_>>>(n, N - n, b, ghost_owner.p, ghost_shift.p, xq.p);
    }
    md_force();
    k_nve_final<<<nblk(n), 256, 0, st_>>>(n, dtf, ltype_d.p, mass_d.p, f.p, v_d.p);
    kernel_launches += 3;
  }
  RXB_CUDA(cudaEventRecord(run_ev_[1], st_));
  ++::host_sync_counter();               // the wait at the end of the run
  RXB_CUDA(cudaEventSynchronize(run_ev_[1]));
  float ms = 0;
  RXB_CUDA(cudaEventElapsedTime(&ms, run_ev_[0], run_ev_[1]));
  last_run_ms = ms;
}

void System::md_get(double* x, double* v, double* fo, double* q) {
  RXB_CUDA(cudaSetDevice(device_));
  x_stage.resize((size_t)3 * N);
  if (x) {
    k_get_xyz<<<nblk(n), 256, 0, st_>>>(n, xq.p, x_stage.p);
    RXB_CUDA(cudaMemcpyAsync(x, x_stage.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, st_));
    RXB_SYNC(st_);
  }
  if (v) RXB_CUDA(cudaMemcpy(v, v_d.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost));
  if (fo) RXB_CUDA(cudaMemcpy(fo, f.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost));
  if (q) {
    k_get_q<<<nblk(n), 256, 0, st_>>>(n, xq.p, x_stage.p);
    RXB_CUDA(cudaMemcpyAsync(q, x_stage.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st_));
    RXB_SYNC(st_);
  }
}

double System::md_kinetic() {
  RXB_CUDA(cudaSetDevice(device_));
  RXB_CUDA(cudaMemsetAsync(virial_d.p, 0, sizeof(double), st_));
  k_kinetic<<<148 * 2, 256, 0, st_>>>(n, ltype_d.p, mass_d.p, v_d.p, virial_d.p);
  if (dist_) dist_allreduce(virial_d.p, 1);
  double ke = 0;
  RXB_CUDA(cudaMemcpyAsync(&ke, virial_d.p, sizeof(double), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
  return 0.5 * kMvv2e * ke;
}

// ---- introspection: the S-space lists translated into the caller's index space on the host (parity tests only) ----
namespace {
template <class T>
std::vector<T> fetch(const T* dev, size_t count, cudaStream_t st) {
  std::vector<T> h(std::max<size_t>(count, 1));
  if (count) {
    RXB_CUDA(cudaMemcpyAsync(h.data(), dev, count * sizeof(T), cudaMemcpyDeviceToHost, st));
    RXB_SYNC(st);
  }
  return h;
}
}  // namespace

// compact CSR over the local atoms (row i = atom i), columns = atom indices in the order the device row holds them
void System::export_verlet(long long* off, int* idx) {
  RXB_CUDA(cudaSetDevice(device_));
  const std::vector<int> cnt = fetch(vl.cnt.p, (size_t)n, st_), ra = fetch(row_atom.p, (size_t)n, st_),
                         sa = fetch(s2a.p, (size_t)N, st_), raw = fetch(vl.idx.p, (size_t)vl.slots, st_);
  std::vector<int> row_of(std::max(n, 1));
  for (int r = 0; r < n; r++) row_of[ra[r]] = r;
  long long w = 0;
  for (int i = 0; i < n; i++) {
    const int r = row_of[i];
    off[i] = w;
    const int* src = raw.data() + (size_t)r * vl.stride;
    for (int k = 0; k < cnt[r]; k++) idx[w + k] = sa[src[k]];
    w += cnt[r];
  }
  off[n] = w;
}

// far list / H: num[i] entries for atom i, written at the compact Verlet offset of atom i (the layout rxb_get_far documents)
void System::export_far(int* num, int* idx, double* val) {
  RXB_CUDA(cudaSetDevice(device_));
  const std::vector<int> cnt = fetch(vl.cnt.p, (size_t)n, st_), ra = fetch(row_atom.p, (size_t)n, st_),
                         sa = fetch(s2a.p, (size_t)N, st_), fn = fetch(far_num.p, (size_t)n, st_);
  std::vector<unsigned long long> pk;
  std::vector<int> fi;
  std::vector<double> hv;
  if (h_packed_) pk = fetch(hpk.p, (size_t)vl.slots, st_);
  else { fi = fetch(far_idx.p, (size_t)vl.slots, st_); hv = fetch(H_val.p, (size_t)vl.slots, st_); }
  std::vector<int> row_of(std::max(n, 1));
  for (int r = 0; r < n; r++) row_of[ra[r]] = r;
  long long w = 0;
  for (int i = 0; i < n; i++) {
    const int r = row_of[i];
    num[i] = fn[r];
    const size_t base = (size_t)r * vl.stride;
    for (int k = 0; k < fn[r]; k++) {
      if (h_packed_) {
        const unsigned long long e = pk[base + k];
        idx[w + k] = sa[(int)(e >> kHColShift)];
        val[w + k] = (double)(long long)(e & kHValMask) / h_quant_;
      } else {
        idx[w + k] = sa[fi[base + k]];
        val[w + k] = hv[base + k];
      }
    }
    w += cnt[r];
  }
}

}  // namespace rxb
