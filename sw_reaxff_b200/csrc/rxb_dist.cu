// Multi-GPU: LAMMPS-style spatial domain decomposition, one process per GPU, NCCL over NVLink/NVSwitch.
//
// What the reference gets from the (absent) LAMMPS core over MPI — exchange/borders every reneighbouring step,
// forward_comm of x every step, reverse_comm of f, forward_comm_fix of the CG search direction and MPI_Allreduce of the
// dot products every CG iteration (fix_qeq_reax_sunway.cpp:1043-1132) — is done here on the device:
//   * bricks in lamda space (px x py x pz), ghost shell = cutneigh;
//   * reneighbouring ("exchange + borders"): every rank all-gathers the 144-byte migration records of all atoms
//     (x, v, q, s_hist, t_hist, tag, type), then selects its new local atoms and its ghost images from the gathered set
//     with two count/scan/fill kernels.  One collective instead of 6-direction staged swaps: on NVSwitch every peer is one
//     hop at full bandwidth, and the gathered set makes migration of s_hist/t_hist/v trivial;
//   * forward (x,q), CG halo (d as double2), reverse (f): peer-to-peer boundary exchange.  At every exchange each rank
//     lists, per peer, the local atoms that peer holds as ghosts (dist_build_plan); a step then packs those atoms,
//     issues one grouped ncclSend/ncclRecv per peer pair and unpacks (+ image shift) — only the boundary layer moves
//     (8-24 B per ghost).  rxb_dist_set_p2p(0) falls back to whole-slab all-gathers / reduce-scatter for comparison;
//   * CG dots / sums / energies: ncclAllReduce on the device scalars, stream-ordered, no host sync.
// Parity-tested against the single-GPU path (positions, forces, charges, energies, species; incl. migration).  What is
// left on the table at N > 1 is NCCL launch latency per CG iteration (halo + all-reduce are serial with the SpMV).
#include <cub/cub.cuh>
#include <nccl.h>

#include <cmath>

#include "rxb_system.h"

#define RXB_NCCL(call)                                                                                    \
  do {                                                                                                    \
    ncclResult_t r_ = (call);                                                                             \
    if (r_ != ncclSuccess)                                                                                \
      throw std::runtime_error(std::string("NCCL error: ") + ncclGetErrorString(r_) + " at " __FILE__ ":" + \
                               std::to_string(__LINE__));                                                 \
  } while (0)

namespace rxb {

constexpr int kRec = 18;  // doubles per migration record: x3 v3 q s_hist5 t_hist5 (tag,type)

struct Dist {
  int rank = 0, world = 1;
  int grid[3] = {1, 1, 1}, coord[3] = {0, 0, 0};
  double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  ncclComm_t comm = nullptr;
  int chunk = 0;                 // slots per rank in the gathered arrays
  std::vector<int> counts;       // local atoms per rank at the last exchange
  DBuf<int> counts_d;
  DBuf<double> rec_send, rec_all;        // [chunk][kRec], [world*chunk][kRec]
  DBuf<double> rec2_send, rec2_all;      // compact post-migration records [..][5]
  DBuf<double4> xq_all;                  // [world*chunk]
  DBuf<double2> d_all;                   // [world*chunk]
  DBuf<double> f_all, f_recv;            // [world*chunk][3], [chunk][3]
  DBuf<int> gsrc;                        // per ghost: source slot in the gathered arrays
  DBuf<long long> flag, off;             // selection scans
  DBuf<char> temp;
  // peer-to-peer boundary exchange plan (rebuilt at every exchange): ghosts are ordered by source slot, hence grouped by
  // source rank; rank r sends me exactly the local atoms I listed, in my ghost order, so receives land in place
  bool p2p = true;
  std::vector<int> need_from, send_to, goff, soff;   // per peer: ghosts I need / atoms I send, and their offsets
  int nsend = 0;
  DBuf<int> greq, sendlist, cnt_d, cnt_all_d;
  DBuf<double> sendbuf, recvbuf;
  DBuf<double> dots_all;                 // [world][4]: every rank's partial CG dot products (dist_forward2_dots)
  DBuf<int> send_s, self_s;              // sorted positions of the atoms I send / of the sources of my own periodic images
  DBuf<double> recv2;                    // staging of received double2 ghost values, ghost order
  // ---- peer-memory exchange (NVLink P2P stores, no NCCL in the CG iteration): see "peer exchange" below
  bool peer_ok = false;                  // every rank mapped every other rank's window
  bool peer_plan_ok = false;             // ... and the current plan fits the halo capacity on every rank
  bool peer_use = true;                  // RXB_PEER=0 keeps the NCCL path (A/B, debugging)
  char* win = nullptr;                   // my window (cudaMalloc): flags | dots | halo
  std::vector<char*> win_of;             // mapped window of every rank (win_of[rank] = win)
  size_t cap_g = 0;                      // ghost entries (double2) per parity the halo region holds
  unsigned long long seq = 0;            // exchange sequence number (flags are monotonic)
  DBuf<int> dst_off_d, soff_d;           // per peer: where my values land in its halo; my send offsets (W + 1)
  DBuf<unsigned int> done_d;             // CTA completion counter of the push kernel
  DBuf<int> peer_err_d;                  // set when a wait timed out (a peer died): checked with the end-of-step status
};

namespace {

struct BoxD { double h[6], h_inv[6]; };
struct Brick { double lo[3], hi[3], cg[3]; int m[3]; };

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
__device__ __forceinline__ bool in_brick(const Brick& k, const double* l) {
  return l[0] >= k.lo[0] && l[0] < k.hi[0] && l[1] >= k.lo[1] && l[1] < k.hi[1] && l[2] >= k.lo[2] && l[2] < k.hi[2];
}

__global__ void k_wrap_pack(int n, BoxD b, double4* __restrict__ xq, const double* __restrict__ vel,
                            const double* __restrict__ s_hist, const double* __restrict__ t_hist,
                            const int* __restrict__ tag, const int* __restrict__ ltype, double* __restrict__ rec) {
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
  }
  double* r = rec + (size_t)kRec * i;
  r[0] = p.x; r[1] = p.y; r[2] = p.z;
  r[3] = vel[3 * i]; r[4] = vel[3 * i + 1]; r[5] = vel[3 * i + 2];
  r[6] = p.w;
  for (int k = 0; k < 5; k++) { r[7 + k] = s_hist[5 * (size_t)i + k]; r[12 + k] = t_hist[5 * (size_t)i + k]; }
  r[17] = __longlong_as_double(((long long)tag[i] << 32) | (unsigned int)ltype[i]);
}

__global__ void k_pack_compact(int n, const double4* __restrict__ xq, const int* __restrict__ tag, const int* __restrict__ ltype,
                               double* __restrict__ rec) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = xq[i];
  double* r = rec + (size_t)5 * i;
  r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = p.w;
  r[4] = __longlong_as_double(((long long)tag[i] << 32) | (unsigned int)ltype[i]);
}

__device__ __forceinline__ bool slot_valid(int s, int chunk, const int* counts) { return (s % chunk) < counts[s / chunk]; }

__global__ void k_flag_locals(int nslots, int chunk, const int* __restrict__ counts, const double* __restrict__ rec, BoxD b,
                              Brick k, long long* __restrict__ flag) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nslots) return;
  long long f = 0;
  if (s < nslots && slot_valid(s, chunk, counts)) {
    const double* r = rec + (size_t)kRec * s;
    double l[3];
    x2lamda(b, r[0], r[1], r[2], l);
    f = in_brick(k, l) ? 1 : 0;
  }
  flag[s] = f;
}

__global__ void k_fill_locals(int nslots, const long long* __restrict__ flag, const long long* __restrict__ off,
                              const double* __restrict__ rec, const int* __restrict__ map, int maplen, double4* __restrict__ xq,
                              double* __restrict__ vel, double* __restrict__ s_hist, double* __restrict__ t_hist,
                              int* __restrict__ tag, int* __restrict__ ltype, int* __restrict__ type) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots || !flag[s]) return;
  const long long i = off[s];
  const double* r = rec + (size_t)kRec * s;
  xq[i] = make_double4(r[0], r[1], r[2], r[6]);
  vel[3 * i] = r[3]; vel[3 * i + 1] = r[4]; vel[3 * i + 2] = r[5];
  for (int k = 0; k < 5; k++) { s_hist[5 * i + k] = r[7 + k]; t_hist[5 * i + k] = r[12 + k]; }
  const long long tt = __double_as_longlong(r[17]);
  const int tg = (int)(tt >> 32), lt = (int)(tt & 0xffffffffLL);
  tag[i] = tg; ltype[i] = lt;
  type[i] = (lt >= 1 && lt < maplen) ? map[lt] : -1;
}

// STRIDE doubles per record, charge at QOFF, packed (tag,type) at TOFF
template <bool FILL, int STRIDE, int QOFF, int TOFF>
__global__ void k_ghosts_dist(int nslots, int chunk, const int* __restrict__ counts, const double* __restrict__ rec, BoxD b,
                              Brick k, int n, const int* __restrict__ map, int maplen, long long* __restrict__ count,
                              const long long* __restrict__ off, double4* __restrict__ xq, int* __restrict__ tag,
                              int* __restrict__ ltype, int* __restrict__ type, int* __restrict__ gsrc, int* __restrict__ shift) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nslots) return;
  if (s == nslots || !slot_valid(s, chunk, counts)) { if (!FILL) count[s] = 0; return; }
  const double* r = rec + (size_t)STRIDE * s;
  double l[3];
  x2lamda(b, r[0], r[1], r[2], l);
  const bool mine = in_brick(k, l);
  long long c = 0, w = FILL ? off[s] : 0;
  for (int sz = -k.m[2]; sz <= k.m[2]; sz++) {
    const double l2 = l[2] + sz;
    if (!(l2 >= k.lo[2] - k.cg[2] && l2 < k.hi[2] + k.cg[2])) continue;
    for (int sy = -k.m[1]; sy <= k.m[1]; sy++) {
      const double l1 = l[1] + sy;
      if (!(l1 >= k.lo[1] - k.cg[1] && l1 < k.hi[1] + k.cg[1])) continue;
      for (int sx = -k.m[0]; sx <= k.m[0]; sx++) {
        if (mine && !sx && !sy && !sz) continue;
        const double l0 = l[0] + sx;
        if (!(l0 >= k.lo[0] - k.cg[0] && l0 < k.hi[0] + k.cg[0])) continue;
        if (FILL) {
          double d[3];
          shift_vec(b, sx, sy, sz, d);
          const long long g = n + w;
          xq[g] = make_double4(r[0] + d[0], r[1] + d[1], r[2] + d[2], r[QOFF]);
          const long long tt = __double_as_longlong(r[TOFF]);
          const int lt = (int)(tt & 0xffffffffLL);
          tag[g] = (int)(tt >> 32); ltype[g] = lt;
          type[g] = (lt >= 1 && lt < maplen) ? map[lt] : -1;
          gsrc[w] = s;
          shift[3 * w] = sx; shift[3 * w + 1] = sy; shift[3 * w + 2] = sz;
          w++;
        } else {
          c++;
        }
      }
    }
  }
  if (!FILL) count[s] = c;
}

__global__ void k_ghost_x_from_all(int n, int nghost, BoxD b, const int* __restrict__ gsrc, const int* __restrict__ shift,
                                   const double4* __restrict__ xq_all, double4* __restrict__ xq) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  double d[3];
  shift_vec(b, shift[3 * g], shift[3 * g + 1], shift[3 * g + 2], d);
  const double4 p = xq_all[gsrc[g]];
  xq[n + g] = make_double4(p.x + d[0], p.y + d[1], p.z + d[2], p.w);
}
__global__ void k_ghost_d_from_all(int n, int nghost, const int* __restrict__ gsrc, const double2* __restrict__ d_all,
                                   double2* __restrict__ vec) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < nghost) vec[n + g] = d_all[gsrc[g]];
}
__global__ void k_scatter_f(int n, int nghost, int my_off, const int* __restrict__ gsrc, const double* __restrict__ f,
                            double* __restrict__ f_all) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n + nghost) return;
  const long long dst = i < n ? (long long)my_off + i : gsrc[i - n];
  const double fx = f[3 * i], fy = f[3 * i + 1], fz = f[3 * i + 2];
  if (fx != 0.0) atomicAdd(&f_all[3 * dst], fx);
  if (fy != 0.0) atomicAdd(&f_all[3 * dst + 1], fy);
  if (fz != 0.0) atomicAdd(&f_all[3 * dst + 2], fz);
}

inline int nblk(long n, int t = 256) { return (int)((n + t - 1) / t); }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
void System::dist_unique_id(char* out128) {
  ncclUniqueId id;
  RXB_NCCL(ncclGetUniqueId(&id));
  static_assert(sizeof(id) <= 128, "ncclUniqueId larger than 128 bytes");
  memset(out128, 0, 128);
  memcpy(out128, &id, sizeof(id));
}

void System::dist_init(int rank, int world, const char* id128, int px, int py, int pz) {
  RXB_CUDA(cudaSetDevice(device_));
  if (px * py * pz != world) throw std::runtime_error("rxb_dist_init: processor grid does not match the world size");
  dist_ = new Dist();
  Dist& D = *dist_;
  D.rank = rank; D.world = world;
  D.grid[0] = px; D.grid[1] = py; D.grid[2] = pz;
  D.coord[0] = rank % px; D.coord[1] = (rank / px) % py; D.coord[2] = rank / (px * py);
  for (int t = 0; t < 3; t++) {
    D.lo[t] = (double)D.coord[t] / D.grid[t];
    D.hi[t] = (D.coord[t] == D.grid[t] - 1) ? 1.0 : (double)(D.coord[t] + 1) / D.grid[t];
  }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  RXB_NCCL(ncclCommInitRank(&D.comm, world, id, rank));
  D.counts.assign(world, 0);
  D.counts_d.resize(world);
  dist_peer_setup();
}

void System::dist_destroy() {
  if (!dist_) return;
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  for (int r = 0; r < (int)dist_->win_of.size(); r++)
    if (r != dist_->rank && dist_->win_of[r]) cudaIpcCloseMemHandle(dist_->win_of[r]);
  if (dist_->comm) {                     // nobody may unmap or free a window a peer could still be storing into
    DBuf<int> bar;
    bar.resize(1);
    cudaMemsetAsync(bar.p, 0, sizeof(int), st_);
    ncclAllReduce(bar.p, bar.p, 1, ncclInt, ncclSum, dist_->comm, st_);
    cudaStreamSynchronize(st_);
  }
  if (dist_->win) cudaFree(dist_->win);
  if (dist_->comm) ncclCommDestroy(dist_->comm);
  delete dist_;
  dist_ = nullptr;
}

int System::dist_world() const { return dist_ ? dist_->world : 1; }
size_t System::slab() const { return dist_ ? (size_t)dist_->chunk : 0; }

void System::dist_allreduce(double* dev_ptr, int count) {
  if (!dist_) return;
  RXB_NCCL(ncclAllReduce(dev_ptr, dev_ptr, count, ncclDouble, ncclSum, dist_->comm, st_));
}

int System::dist_rank() const { return dist_ ? dist_->rank : 0; }
void System::dist_allreduce_int(int* dev_ptr, size_t count) {
  if (!dist_) return;
  RXB_NCCL(ncclAllReduce(dev_ptr, dev_ptr, count, ncclInt, ncclSum, dist_->comm, st_));
}
void System::dist_allreduce_max_int(int* dev_ptr, size_t count) {
  if (!dist_) return;
  RXB_NCCL(ncclAllReduce(dev_ptr, dev_ptr, count, ncclInt, ncclMax, dist_->comm, st_));
}
void System::dist_allgather_int(const int* send, int* recv, size_t count_per_rank) {
  if (!dist_) {
    if (send != recv) RXB_CUDA(cudaMemcpyAsync(recv, send, count_per_rank * sizeof(int), cudaMemcpyDeviceToDevice, st_));
    return;
  }
  RXB_NCCL(ncclAllGather(send, recv, count_per_rank, ncclInt, dist_->comm, st_));
}

// exchange + borders
void System::dist_exchange() {
  Dist& D = *dist_;
  BoxD b;
  memcpy(b.h, box.h, sizeof(b.h));
  memcpy(b.h_inv, box.h_inv, sizeof(b.h_inv));
  // 1. counts
  int my = n;
  RXB_CUDA(cudaMemcpyAsync(D.counts_d.p + D.rank, &my, sizeof(int), cudaMemcpyHostToDevice, st_));
  RXB_NCCL(ncclAllGather(D.counts_d.p + D.rank, D.counts_d.p, 1, ncclInt, D.comm, st_));
  RXB_CUDA(cudaMemcpyAsync(D.counts.data(), D.counts_d.p, D.world * sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaStreamSynchronize(st_));
  int mx = 0;
  long long total = 0;
  for (int c : D.counts) { mx = std::max(mx, c); total += c; }
  // slots per rank: generous, so that the post-migration local counts fit as well (growth triggers a re-gather next time)
  int want = std::max(mx, (int)((total / D.world) + (total / D.world) / 4 + 64));
  if (want > D.chunk) D.chunk = want + want / 8;
  int chunk = D.chunk;
  int nslots = chunk * D.world;
  // 2. migration records of the wrapped local atoms
  D.rec_send.resize((size_t)chunk * kRec);
  D.rec_all.resize((size_t)nslots * kRec);
  k_wrap_pack<<<nblk(n), 256, 0, st_>>>(n, b, xq.p, v_d.p, q_s_hist.p, q_t_hist.p, tag.p, ltype_d.p, D.rec_send.p);
  RXB_NCCL(ncclAllGather(D.rec_send.p, D.rec_all.p, (size_t)chunk * kRec, ncclDouble, D.comm, st_));
  // 3. brick and shell in lamda space
  Brick k;
  const double cut = cutneigh();
  k.cg[0] = cut * sqrt(b.h_inv[0] * b.h_inv[0] + b.h_inv[5] * b.h_inv[5] + b.h_inv[4] * b.h_inv[4]);
  k.cg[1] = cut * sqrt(b.h_inv[1] * b.h_inv[1] + b.h_inv[3] * b.h_inv[3]);
  k.cg[2] = cut * b.h_inv[2];
  for (int t = 0; t < 3; t++) { k.lo[t] = D.lo[t]; k.hi[t] = D.hi[t]; k.m[t] = (int)ceil(k.cg[t]) + 1; }
  // 4. new local atoms: flag, scan, fill (stable => deterministic order)
  D.flag.resize(nslots + 1); D.off.resize(nslots + 1);
  k_flag_locals<<<nblk(nslots + 1), 256, 0, st_>>>(nslots, chunk, D.counts_d.p, D.rec_all.p, b, k, D.flag.p);
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, D.flag.p, D.off.p, nslots + 1, st_);
  D.temp.resize(need + 16);
  cub::DeviceScan::ExclusiveSum(D.temp.p, need, D.flag.p, D.off.p, nslots + 1, st_);
  long long newn = 0;
  RXB_CUDA(cudaMemcpyAsync(&newn, D.off.p + nslots, sizeof(long long), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaStreamSynchronize(st_));
  // (a brick may end up with more atoms than a slab slot holds: the slab is re-sized below, identically on every rank,
  // from the all-gathered post-migration counts - no rank-local abort that would leave the others in a collective)
  n = (int)newn;
  N = n;
  ensure_atom_capacity();
  v_d.resize_keep((size_t)3 * std::max(n, chunk));
  q_s_hist.resize((size_t)5 * std::max(n, chunk)); q_t_hist.resize((size_t)5 * std::max(n, chunk));
  q_s_hist.n = (size_t)5 * n; q_t_hist.n = (size_t)5 * n;
  if (map_d.n != ff.map.size()) {
    map_d.resize(ff.map.size());
    RXB_CUDA(cudaMemcpyAsync(map_d.p, ff.map.data(), ff.map.size() * sizeof(int), cudaMemcpyHostToDevice, st_));
  }
  xq.resize_keep(std::max((size_t)n, (size_t)chunk)); xq.n = n;
  tag.resize_keep(std::max((size_t)n, (size_t)chunk)); ltype_d.resize_keep(std::max((size_t)n, (size_t)chunk));
  type.resize_keep(std::max((size_t)n, (size_t)chunk));
  k_fill_locals<<<nblk(nslots), 256, 0, st_>>>(nslots, D.flag.p, D.off.p, D.rec_all.p, map_d.p, (int)ff.map.size(), xq.p, v_d.p,
                                              q_s_hist.p, q_t_hist.p, tag.p, ltype_d.p, type.p);
  // 5. post-migration layout: all-gather the new counts and a compact (x,y,z,q,tag|type) record of the NEW local atoms;
  //    ghosts are selected from that, so their source slots refer to the layout every later halo exchange uses and come
  //    out sorted by slot, i.e. grouped by source rank (what the peer-to-peer plan relies on)
  my = n;
  RXB_CUDA(cudaMemcpyAsync(D.counts_d.p + D.rank, &my, sizeof(int), cudaMemcpyHostToDevice, st_));
  RXB_NCCL(ncclAllGather(D.counts_d.p + D.rank, D.counts_d.p, 1, ncclInt, D.comm, st_));
  RXB_CUDA(cudaMemcpyAsync(D.counts.data(), D.counts_d.p, D.world * sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaStreamSynchronize(st_));
  {
    int mx2 = 0;
    for (int c : D.counts) mx2 = std::max(mx2, c);
    if (mx2 > D.chunk) {           // same numbers on every rank, hence the same new slab size everywhere
      D.chunk = mx2 + mx2 / 8 + 64;
      chunk = D.chunk; nslots = chunk * D.world;
      D.flag.resize(nslots + 1); D.off.resize(nslots + 1);
      cub::DeviceScan::ExclusiveSum(nullptr, need, D.flag.p, D.off.p, nslots + 1, st_);
      D.temp.resize(need + 16);
      xq.resize_keep((size_t)chunk); tag.resize_keep((size_t)chunk); ltype_d.resize_keep((size_t)chunk); type.resize_keep((size_t)chunk);
      xq.n = n;
      v_d.resize_keep((size_t)3 * chunk);
      { const size_t keep = q_s_hist.n; q_s_hist.resize_keep((size_t)5 * chunk); q_t_hist.resize_keep((size_t)5 * chunk); q_s_hist.n = keep; q_t_hist.n = keep; }
    }
  }
  D.rec2_send.resize((size_t)chunk * 5); D.rec2_all.resize((size_t)nslots * 5);
  k_pack_compact<<<nblk(n), 256, 0, st_>>>(n, xq.p, tag.p, ltype_d.p, D.rec2_send.p);
  RXB_NCCL(ncclAllGather(D.rec2_send.p, D.rec2_all.p, (size_t)chunk * 5, ncclDouble, D.comm, st_));
  k_ghosts_dist<false, 5, 3, 4><<<nblk(nslots + 1), 256, 0, st_>>>(nslots, chunk, D.counts_d.p, D.rec2_all.p, b, k, n, map_d.p,
                                                                  (int)ff.map.size(), D.flag.p, nullptr, nullptr, nullptr,
                                                                  nullptr, nullptr, nullptr, nullptr);
  cub::DeviceScan::ExclusiveSum(D.temp.p, need, D.flag.p, D.off.p, nslots + 1, st_);
  long long nghost = 0;
  RXB_CUDA(cudaMemcpyAsync(&nghost, D.off.p + nslots, sizeof(long long), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaStreamSynchronize(st_));
  N = n + (int)nghost;
  ensure_atom_capacity();
  xq.resize_keep(std::max((size_t)N, (size_t)chunk)); xq.n = N;
  D.gsrc.resize(std::max<size_t>(nghost, 1));
  ghost_shift.resize(std::max<size_t>(3 * nghost, 3));
  ghost_owner.resize(std::max<size_t>(nghost, 1));
  k_ghosts_dist<true, 5, 3, 4><<<nblk(nslots + 1), 256, 0, st_>>>(nslots, chunk, D.counts_d.p, D.rec2_all.p, b, k, n, map_d.p,
                                                                 (int)ff.map.size(), nullptr, D.off.p, xq.p, tag.p, ltype_d.p,
                                                                 type.p, D.gsrc.p, ghost_shift.p);
  dist_build_plan();
  D.xq_all.resize((size_t)nslots); D.d_all.resize((size_t)nslots);
  D.f_all.resize((size_t)3 * nslots); D.f_recv.resize((size_t)3 * chunk);
  kernel_launches += 6;
  RXB_CUDA(cudaGetLastError());
}

namespace {
__global__ void k_plan_count(int nghost, int chunk, const int* __restrict__ gsrc, int* __restrict__ greq, int* __restrict__ cnt) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  const int s = gsrc[g];
  greq[g] = s % chunk;
  atomicAdd(&cnt[s / chunk], 1);
}
template <int W>
__global__ void k_pack(int m, const int* __restrict__ list, const double* __restrict__ src, double* __restrict__ dst) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const int i = list[e];
#pragma unroll
  for (int t = 0; t < W; t++) dst[(size_t)W * e + t] = src[(size_t)W * i + t];
}
template <int W>
__global__ void k_self_ghosts(int n, int g0, int g1, const int* __restrict__ greq, double* __restrict__ vec) {
  int g = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= g1) return;
  const int i = greq[g];
#pragma unroll
  for (int t = 0; t < W; t++) vec[(size_t)W * (n + g) + t] = vec[(size_t)W * i + t];
}
__global__ void k_shift_ghosts(int n, int nghost, BoxD b, const int* __restrict__ shift, double4* __restrict__ xq) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  const int sx = shift[3 * g], sy = shift[3 * g + 1], sz = shift[3 * g + 2];
  if (!(sx | sy | sz)) return;
  double d[3];
  shift_vec(b, sx, sy, sz, d);
  double4 p = xq[n + g];
  p.x += d[0]; p.y += d[1]; p.z += d[2];
  xq[n + g] = p;
}
__global__ void k_unpack_add_f(int m, const int* __restrict__ list, const double* __restrict__ recv, double* __restrict__ f) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const int i = list[e];
  const double fx = recv[3 * (size_t)e], fy = recv[3 * (size_t)e + 1], fz = recv[3 * (size_t)e + 2];
  if (fx != 0.0) atomicAdd(&f[3 * i], fx);
  if (fy != 0.0) atomicAdd(&f[3 * i + 1], fy);
  if (fz != 0.0) atomicAdd(&f[3 * i + 2], fz);
}
__global__ void k_self_reverse_f(int n, int g0, int g1, const int* __restrict__ greq, double* __restrict__ f) {
  int g = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= g1) return;
  const int i = greq[g];
  for (int t = 0; t < 3; t++) { const double v = f[3 * (size_t)(n + g) + t]; if (v != 0.0) atomicAdd(&f[3 * i + t], v); }
}
}  // namespace

// Build the peer-to-peer plan: who needs which of my atoms (one small all-gather of counts + one grouped send/recv of
// index lists per reneighbouring).
void System::dist_build_plan() {
  Dist& D = *dist_;
  const int W = D.world, nghost = N - n;
  D.cnt_d.resize(W); D.cnt_all_d.resize((size_t)W * W);
  D.greq.resize(std::max(nghost, 1));
  RXB_CUDA(cudaMemsetAsync(D.cnt_d.p, 0, W * sizeof(int), st_));
  if (nghost > 0) k_plan_count<<<nblk(nghost), 256, 0, st_>>>(nghost, D.chunk, D.gsrc.p, D.greq.p, D.cnt_d.p);
  RXB_NCCL(ncclAllGather(D.cnt_d.p, D.cnt_all_d.p, W, ncclInt, D.comm, st_));
  std::vector<int> all((size_t)W * W);
  RXB_CUDA(cudaMemcpyAsync(all.data(), D.cnt_all_d.p, all.size() * sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaStreamSynchronize(st_));
  D.need_from.assign(W, 0); D.send_to.assign(W, 0); D.goff.assign(W + 1, 0); D.soff.assign(W + 1, 0);
  for (int r = 0; r < W; r++) {
    D.need_from[r] = all[(size_t)D.rank * W + r];      // row a = what rank a needs from each rank
    D.send_to[r] = (r == D.rank) ? 0 : all[(size_t)r * W + D.rank];
    D.goff[r + 1] = D.goff[r] + D.need_from[r];
    D.soff[r + 1] = D.soff[r] + D.send_to[r];
  }
  D.nsend = D.soff[W];
  // peer exchange: where my values land in each consumer's halo (its ghost offset for source = me), and whether every
  // rank's ghosts fit the halo windows (decided from the all-gathered table: the same answer on every rank)
  {
    std::vector<int> dst_off(W, 0);
    long long max_ghosts = 0;
    for (int p = 0; p < W; p++) {
      long long tot = 0;
      for (int r = 0; r < W; r++) {
        if (r == D.rank) dst_off[p] = (int)tot;
        tot += all[(size_t)p * W + r];
      }
      max_ghosts = std::max(max_ghosts, tot);
    }
    D.peer_plan_ok = D.peer_ok && max_ghosts <= (long long)D.cap_g;
    D.dst_off_d.resize(W); D.soff_d.resize(W + 1);
    RXB_CUDA(cudaMemcpyAsync(D.dst_off_d.p, dst_off.data(), W * sizeof(int), cudaMemcpyHostToDevice, st_));
    RXB_CUDA(cudaMemcpyAsync(D.soff_d.p, D.soff.data(), (W + 1) * sizeof(int), cudaMemcpyHostToDevice, st_));
    RXB_CUDA(cudaStreamSynchronize(st_));          // dst_off is a local that dies with this scope
  }
  D.sendlist.resize(std::max(D.nsend, 1));
  D.sendbuf.resize((size_t)4 * std::max(D.nsend, 1));
  D.recvbuf.resize((size_t)3 * std::max(D.nsend, 1));
  RXB_NCCL(ncclGroupStart());
  for (int r = 0; r < W; r++) {
    if (r == D.rank) continue;
    if (D.need_from[r] > 0) RXB_NCCL(ncclSend(D.greq.p + D.goff[r], D.need_from[r], ncclInt, r, D.comm, st_));
    if (D.send_to[r] > 0) RXB_NCCL(ncclRecv(D.sendlist.p + D.soff[r], D.send_to[r], ncclInt, r, D.comm, st_));
  }
  RXB_NCCL(ncclGroupEnd());
  kernel_launches++;
}

// forward: my atoms -> peers' ghost slots (width doubles per atom), then my own periodic images
template <int WD>
static void p2p_forward(System& s, Dist& D, double* vec, int n, cudaStream_t st) {
  const int W = D.world;
  if (D.nsend > 0) k_pack<WD><<<nblk(D.nsend), 256, 0, st>>>(D.nsend, D.sendlist.p, vec, D.sendbuf.p);
  RXB_NCCL(ncclGroupStart());
  for (int r = 0; r < W; r++) {
    if (r == D.rank) continue;
    if (D.send_to[r] > 0) RXB_NCCL(ncclSend(D.sendbuf.p + (size_t)WD * D.soff[r], (size_t)WD * D.send_to[r], ncclDouble, r, D.comm, st));
    if (D.need_from[r] > 0) RXB_NCCL(ncclRecv(vec + (size_t)WD * (n + D.goff[r]), (size_t)WD * D.need_from[r], ncclDouble, r, D.comm, st));
  }
  RXB_NCCL(ncclGroupEnd());
  const int g0 = D.goff[D.rank], g1 = D.goff[D.rank + 1];
  if (g1 > g0) k_self_ghosts<WD><<<nblk(g1 - g0), 256, 0, st>>>(n, g0, g1, D.greq.p, vec);
  s.kernel_launches += 2;
}

void System::dist_forward_xq() {
  Dist& D = *dist_;
  BoxD b;
  memcpy(b.h, box.h, sizeof(b.h));
  memcpy(b.h_inv, box.h_inv, sizeof(b.h_inv));
  const int nghost = N - n;
  if (D.p2p) {
    p2p_forward<4>(*this, D, reinterpret_cast<double*>(xq.p), n, st_);
    if (nghost > 0) k_shift_ghosts<<<nblk(nghost), 256, 0, st_>>>(n, nghost, b, ghost_shift.p, xq.p);
    kernel_launches++;
    return;
  }
  RXB_NCCL(ncclAllGather(xq.p, D.xq_all.p, (size_t)D.chunk * 4, ncclDouble, D.comm, st_));
  if (nghost > 0) k_ghost_x_from_all<<<nblk(nghost), 256, 0, st_>>>(n, nghost, b, D.gsrc.p, ghost_shift.p, D.xq_all.p, xq.p);
  kernel_launches++;
}

// ---- peer exchange -------------------------------------------------------------------------------------------------
// Each rank owns one window (cudaMalloc, exported through CUDA IPC, mapped by every other rank of the node):
//   [0, 2 KB)              flags[W]      : flags[r] = sequence number of the last exchange rank r completed towards me
//   [2 KB, + 2*W*8 doubles) dots[2][W][8] : partial sums from every rank, double-buffered by exchange parity
//   [halo_off, ...)        halo[2][cap_g] double2: incoming ghost values in my ghost order, double-buffered
// One exchange = k_peer_push (boundary values stored straight into the consumers' halo over NVLink, partial sums into their
// dots, then - after a system-scope fence by every thread and a last-CTA election - the sequence number into their flags)
// + k_peer_pull on the consumer (acquire-spin on its own flags, ghosts <- halo, partial sums added in rank order so all
// ranks hold bit-identical totals).  Remote stores, local loads; no NCCL call and no host involvement per CG iteration
// (the reference: MPI_Allreduce + forward_comm_fix per iteration, fix_qeq_reax_sunway.cpp:1108-1140).
// Double buffering is sufficient: a rank can only start exchange s+2 after pulling s+1, which needs every peer's push
// s+1, which each peer issues (stream order) after its own pull s - the last reader of the parity-s buffers.
namespace {
constexpr int kMaxPeers = 16;
constexpr size_t kFlagBytes = 2048;
constexpr int kDotSlots = 8;
struct PeerView {
  int W, me;
  char* base[kMaxPeers];
  size_t dots_off, halo_off, cap_g;
};
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(256)
k_peer_push(PeerView P, int par, unsigned long long seq, int nsend, const int* __restrict__ send_s, const int* __restrict__ soff,
            const int* __restrict__ dst_off, double2* vec, int g0, int g1, const int* __restrict__ gs_pos,
            const int* __restrict__ self_s, int ndots, const double* __restrict__ dots, unsigned int* __restrict__ done) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  // boundary values -> the consumers' halo (their ghost order)
  for (int e = tid; e < nsend; e += nth) {
    int p = 0;
    while (e >= soff[p + 1]) p++;                       // W <= 16: a short scan
    double2* halo = reinterpret_cast<double2*>(P.base[p] + P.halo_off) + (size_t)par * P.cap_g;
    halo[dst_off[p] + (e - soff[p])] = vec[send_s[e]];
  }
  // my own periodic images (local copy; sources are local atoms, destinations ghosts: disjoint)
  for (int g = g0 + tid; g < g1; g += nth) vec[gs_pos[g]] = vec[self_s[g - g0]];
  // partial sums -> every peer
  if (blockIdx.x == 0 && threadIdx.x < ndots * P.W) {
    const int p = threadIdx.x / ndots, k = threadIdx.x % ndots;
    if (p != P.me) {
      double* d = reinterpret_cast<double*>(P.base[p] + P.dots_off) + ((size_t)par * P.W + P.me) * kDotSlots;
      d[k] = dots[k];
    }
  }
  __threadfence_system();                               // every thread: its remote stores are visible system-wide ...
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(done, 1u);
    if (t == gridDim.x - 1) {                           // ... before the last CTA publishes the sequence number
      *done = 0;
      __threadfence_system();
      for (int p = 0; p < P.W; p++)
        if (p != P.me) st_release_sys(reinterpret_cast<unsigned long long*>(P.base[p]) + P.me, seq);
    }
  }
}

__global__ void __launch_bounds__(256)
k_peer_pull(PeerView P, int par, unsigned long long seq, int nghost, int g0, int g1, const int* __restrict__ gs_pos,
            double2* __restrict__ vec, int ndots, double* __restrict__ dots, int* __restrict__ err) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = 1;
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(P.base[P.me]);
    const unsigned long long t0 = globaltimer_ns();
    for (int r = 0; r < P.W && ok; r++) {
      if (r == P.me) continue;
      while (ld_acquire_sys(flags + r) < seq) {
        if (globaltimer_ns() - t0 > 10000000000ULL) { ok = 0; atomicExch(err, 1); break; }   // 10 s: a peer is gone
      }
    }
  }
  __syncthreads();
  if (!ok) return;
  const double2* halo = reinterpret_cast<const double2*>(P.base[P.me] + P.halo_off) + (size_t)par * P.cap_g;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < nghost; g += gridDim.x * blockDim.x)
    if (g < g0 || g >= g1) vec[gs_pos[g]] = halo[g];
  if (blockIdx.x == 0 && threadIdx.x < ndots) {
    const double* d = reinterpret_cast<const double*>(P.base[P.me] + P.dots_off) + (size_t)par * P.W * kDotSlots;
    const double mine = dots[threadIdx.x];
    double s = 0.0;
    for (int r = 0; r < P.W; r++) s += (r == P.me) ? mine : d[(size_t)r * kDotSlots + threadIdx.x];
    dots[threadIdx.x] = s;                              // the same numbers added in the same (rank) order on every rank
  }
}

PeerView peer_view(const Dist& D) {
  PeerView P{};
  P.W = D.world; P.me = D.rank;
  for (int r = 0; r < D.world; r++) P.base[r] = D.win_of[r];
  P.dots_off = kFlagBytes;
  P.halo_off = kFlagBytes + (((size_t)2 * D.world * kDotSlots * sizeof(double) + 255) / 256) * 256;
  P.cap_g = D.cap_g;
  return P;
}

// vec: S-space double2 vector whose ghosts are refreshed (null: pure reduction); dots[ndots]: partial sums -> totals
void peer_exchange(System& s, Dist& D, double2* vec, int nghost, const int* gs_pos, double* dots, int ndots, cudaStream_t st) {
  D.seq++;
  const int par = (int)(D.seq & 1);
  const PeerView P = peer_view(D);
  const int g0 = vec ? D.goff[D.rank] : 0, g1 = vec ? D.goff[D.rank + 1] : 0;
  const int nsend = vec ? D.nsend : 0;
  const int work = std::max(nsend, g1 - g0);
  k_peer_push<<<std::max(1, std::min(148, (work + 255) / 256)), 256, 0, st>>>(P, par, D.seq, nsend, D.send_s.p, D.soff_d.p,
                                                                           D.dst_off_d.p, vec, g0, g1, gs_pos, D.self_s.p,
                                                                           ndots, dots, D.done_d.p);
  const int ng = vec ? nghost : 0;
  k_peer_pull<<<std::max(1, std::min(296, (ng + 255) / 256)), 256, 0, st>>>(P, par, D.seq, ng, g0, g1, gs_pos, vec, ndots, dots,
                                                                         D.peer_err_d.p);
  s.kernel_launches += 2;
}
}  // namespace

// Map every rank's window (CUDA IPC; all ranks are processes on one node).  Any failure on any rank (no peer access,
// IPC disabled in the container, more than 16 ranks) leaves the NCCL send/recv path in place on ALL ranks.
void System::dist_peer_setup() {
  Dist& D = *dist_;
  const int W = D.world;
  const char* e = getenv("RXB_PEER");
  D.peer_use = e ? atoi(e) != 0 : true;
  D.peer_ok = false;
  D.done_d.resize(1); D.peer_err_d.resize(1);
  RXB_CUDA(cudaMemsetAsync(D.done_d.p, 0, sizeof(unsigned int), st_));
  RXB_CUDA(cudaMemsetAsync(D.peer_err_d.p, 0, sizeof(int), st_));
  int ok = (D.peer_use && W <= kMaxPeers) ? 1 : 0;
  const char* ce = getenv("RXB_PEER_CAP");
  D.cap_g = ce ? (size_t)atol(ce) : ((size_t)1 << 20);
  D.win_of.assign(W, nullptr);
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    const size_t bytes = peer_view(D).halo_off + (size_t)2 * D.cap_g * sizeof(double2);
    if (cudaMalloc(&D.win, bytes) != cudaSuccess) { cudaGetLastError(); D.win = nullptr; ok = 0; }
    else {
      RXB_CUDA(cudaMemsetAsync(D.win, 0, bytes, st_));
      if (cudaIpcGetMemHandle(&mine, D.win) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
  }
  // all-gather the handles (64 bytes each) through NCCL
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  DBuf<char> hs, ha;
  hs.resize(64); ha.resize((size_t)64 * W);
  RXB_CUDA(cudaMemcpyAsync(hs.p, &mine, 64, cudaMemcpyHostToDevice, st_));
  RXB_NCCL(ncclAllGather(hs.p, ha.p, 64, ncclChar, D.comm, st_));
  std::vector<cudaIpcMemHandle_t> all(W);
  RXB_CUDA(cudaMemcpyAsync(all.data(), ha.p, (size_t)64 * W, cudaMemcpyDeviceToHost, st_));
  // does every rank have a window?
  DBuf<int> flag;
  flag.resize(1);
  RXB_CUDA(cudaMemcpyAsync(flag.p, &ok, sizeof(int), cudaMemcpyHostToDevice, st_));
  RXB_NCCL(ncclAllReduce(flag.p, flag.p, 1, ncclInt, ncclMin, D.comm, st_));
  RXB_CUDA(cudaMemcpyAsync(&ok, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaStreamSynchronize(st_));
  if (ok) {
    for (int r = 0; r < W; r++) {
      if (r == D.rank) { D.win_of[r] = D.win; continue; }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      D.win_of[r] = (char*)p;
    }
  }
  RXB_CUDA(cudaMemcpyAsync(flag.p, &ok, sizeof(int), cudaMemcpyHostToDevice, st_));
  RXB_NCCL(ncclAllReduce(flag.p, flag.p, 1, ncclInt, ncclMin, D.comm, st_));
  RXB_CUDA(cudaMemcpyAsync(&ok, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaStreamSynchronize(st_));
  D.peer_ok = ok != 0;
  if (getenv("RXB_PEER_VERBOSE") && D.rank == 0)
    fprintf(stderr, "rxb dist: peer-memory exchange %s (%d ranks, halo capacity %zu ghosts)\n", D.peer_ok ? "ON" : "off (NCCL send/recv)",
            W, D.cap_g);
}

bool System::dist_peer_active() const { return dist_ && dist_->peer_ok && dist_->peer_plan_ok; }

// ---- S-space vectors (the CG search direction): locals and ghosts are interleaved in cell-sorted order, so the boundary
// values are packed through send_s (sorted positions of the atoms each peer needs), received into a staging buffer in ghost
// order and scattered to the ghosts' sorted positions by one kernel that also serves this rank's own periodic images.
namespace {
__global__ void k_map_idx(int m, const int* __restrict__ list, const int* __restrict__ a2s, int* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < m) out[e] = a2s[list[e]];
}
__global__ void k_unpack_ghosts2(int nghost, int g0, int g1, const int* __restrict__ gs_pos, const int* __restrict__ self_s,
                                 const double2* __restrict__ recv, double2* __restrict__ vec) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  vec[gs_pos[g]] = (g >= g0 && g < g1) ? vec[self_s[g - g0]] : recv[g];
}
}  // namespace

void System::dist_sorted_maps() {
  Dist& D = *dist_;
  const int nghost = N - n;
  D.send_s.resize(std::max(D.nsend, 1));
  if (D.nsend > 0) k_map_idx<<<nblk(D.nsend), 256, 0, st_>>>(D.nsend, D.sendlist.p, a2s.p, D.send_s.p);
  const int g0 = D.goff[D.rank], g1 = D.goff[D.rank + 1];
  D.self_s.resize(std::max(g1 - g0, 1));
  if (g1 > g0) k_map_idx<<<nblk(g1 - g0), 256, 0, st_>>>(g1 - g0, D.greq.p + g0, a2s.p, D.self_s.p);
  D.recv2.resize((size_t)2 * std::max(nghost, 1));
  kernel_launches += 2;
}

static void s_forward2(System& s, Dist& D, double2* vec, int n, int nghost, const int* gs_pos, double* dots, cudaStream_t st) {
  const int W = D.world;
  if (D.peer_ok && D.peer_plan_ok) { peer_exchange(s, D, vec, nghost, gs_pos, dots, dots ? 4 : 0, st); return; }
  double* v = reinterpret_cast<double*>(vec);
  if (D.nsend > 0) k_pack<2><<<nblk(D.nsend), 256, 0, st>>>(D.nsend, D.send_s.p, v, D.sendbuf.p);
  if (dots) D.dots_all.resize((size_t)4 * W);
  RXB_NCCL(ncclGroupStart());
  for (int r = 0; r < W; r++) {
    if (r == D.rank) continue;
    if (dots) {
      RXB_NCCL(ncclSend(dots, 4, ncclDouble, r, D.comm, st));
      RXB_NCCL(ncclRecv(D.dots_all.p + (size_t)4 * r, 4, ncclDouble, r, D.comm, st));
    }
    if (D.send_to[r] > 0) RXB_NCCL(ncclSend(D.sendbuf.p + (size_t)2 * D.soff[r], (size_t)2 * D.send_to[r], ncclDouble, r, D.comm, st));
    if (D.need_from[r] > 0) RXB_NCCL(ncclRecv(D.recv2.p + (size_t)2 * D.goff[r], (size_t)2 * D.need_from[r], ncclDouble, r, D.comm, st));
  }
  RXB_NCCL(ncclGroupEnd());
  if (nghost > 0)
    k_unpack_ghosts2<<<nblk(nghost), 256, 0, st>>>(nghost, D.goff[D.rank], D.goff[D.rank + 1], gs_pos, D.self_s.p,
                                                  reinterpret_cast<const double2*>(D.recv2.p), vec);
  s.kernel_launches += 2;
}

void System::dist_forward2(double2* vec) {
  s_forward2(*this, *dist_, vec, n, N - n, gs_pos.p, nullptr, st_);
}

// sum of the per-rank partials in rank order: every rank adds the same numbers in the same order, so all ranks hold the
// bit-identical result (and take the same convergence decisions)
__global__ void k_sum_dots(int world, int rank, const double* __restrict__ all, double* __restrict__ dots) {
  const int c = threadIdx.x;
  if (c >= 4) return;
  const double mine = dots[c];
  double s = 0.0;
  for (int r = 0; r < world; r++) s += (r == rank) ? mine : all[4 * r + c];
  dots[c] = s;
}

// One exchange per CG iteration instead of two: the 4 partial dot products of the sweep travel in the SAME grouped
// send/recv as the boundary values of d (to every rank, 32 bytes each), replacing the separate ncclAllReduce
// (the reference: MPI_Allreduce + comm->forward_comm_fix per iteration, fix_qeq_reax_sunway.cpp:1108-1140).
void System::dist_forward2_dots(double2* vec, double* dots) {
  Dist& D = *dist_;
  s_forward2(*this, D, vec, n, N - n, gs_pos.p, dots, st_);
  if (D.peer_ok && D.peer_plan_ok) return;         // the pull kernel has already formed the totals
  k_sum_dots<<<1, 32, 0, st_>>>(D.world, D.rank, D.dots_all.p, dots);
  kernel_launches++;
}

// sum of a few device scalars over all ranks, identical bits on every rank (count <= 8)
void System::dist_sum_small(double* dev_ptr, int count) {
  if (!dist_) return;
  Dist& D = *dist_;
  if (D.peer_ok && D.peer_plan_ok && count <= kDotSlots) { peer_exchange(*this, D, nullptr, 0, nullptr, dev_ptr, count, st_); return; }
  dist_allreduce(dev_ptr, count);
}

// a wait of the peer exchange timed out (a peer process died): reported with the end-of-step status
void System::dist_peer_check() {
  if (!dist_ || !dist_->peer_ok) return;
  int err = 0;
  RXB_CUDA(cudaMemcpyAsync(&err, dist_->peer_err_d.p, sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaStreamSynchronize(st_));
  if (err) throw std::runtime_error("rxb dist: peer-memory exchange timed out (a peer rank stopped responding)");
}

void System::dist_reverse_f() {
  Dist& D = *dist_;
  if (D.p2p) {
    const int W = D.world;
    RXB_NCCL(ncclGroupStart());
    for (int r = 0; r < W; r++) {
      if (r == D.rank) continue;
      if (D.need_from[r] > 0) RXB_NCCL(ncclSend(f.p + (size_t)3 * (n + D.goff[r]), (size_t)3 * D.need_from[r], ncclDouble, r, D.comm, st_));
      if (D.send_to[r] > 0) RXB_NCCL(ncclRecv(D.recvbuf.p + (size_t)3 * D.soff[r], (size_t)3 * D.send_to[r], ncclDouble, r, D.comm, st_));
    }
    RXB_NCCL(ncclGroupEnd());
    if (D.nsend > 0) k_unpack_add_f<<<nblk(D.nsend), 256, 0, st_>>>(D.nsend, D.sendlist.p, D.recvbuf.p, f.p);
    const int g0 = D.goff[D.rank], g1 = D.goff[D.rank + 1];
    if (g1 > g0) k_self_reverse_f<<<nblk(g1 - g0), 256, 0, st_>>>(n, g0, g1, D.greq.p, f.p);
    kernel_launches += 2;
    return;
  }
  const size_t nslots = (size_t)D.chunk * D.world;
  RXB_CUDA(cudaMemsetAsync(D.f_all.p, 0, 3 * nslots * sizeof(double), st_));
  k_scatter_f<<<nblk(N), 256, 0, st_>>>(n, N - n, D.rank * D.chunk, D.gsrc.p, f.p, D.f_all.p);
  RXB_NCCL(ncclReduceScatter(D.f_all.p, D.f_recv.p, (size_t)3 * D.chunk, ncclDouble, ncclSum, D.comm, st_));
  RXB_CUDA(cudaMemcpyAsync(f.p, D.f_recv.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToDevice, st_));
  kernel_launches++;
}

void System::dist_set_p2p(bool on) { if (dist_) dist_->p2p = on; }

}  // namespace rxb
