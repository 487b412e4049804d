// Multi-GPU: LAMMPS-style spatial domain decomposition, one process per GPU, NCCL over NVLink/NVSwitch.
//
// What the reference gets from the (absent) LAMMPS core over MPI — exchange/borders every reneighbouring step,
// forward_comm of x every step, reverse_comm of f, forward_comm_fix of the CG search direction and MPI_Allreduce of the
// dot products every CG iteration (fix_qeq_reax_sunway.cpp:1043-1132) — is done here on the device:
//   * bricks in lamda space (px x py x pz), ghost shell = cutneigh;
//   * reneighbouring ("exchange + borders"): atoms that left their brick travel as 144-byte records to their new owner
//     only; every rank then works out which of its atoms (and periodic images) lie in the ghost shell of which rank and
//     sends 48-byte records there (dist_exchange).  Received bytes per rank scale with the shell, not with the system;
//   * forward (x,q) and reverse (f) every step: grouped ncclSend/ncclRecv of the boundary values between the ranks that
//     share atoms, along the send lists the borders step produced;
//   * every CG iteration: boundary values of the search direction and the four partial dot products are STORED straight
//     into the consumers' memory over NVLink (CUDA IPC windows, flag-ordered, "peer exchange" below): no NCCL call and
//     no host involvement inside the solve; falls back to grouped send/recv when peer mapping is unavailable;
//   * energies / virial / status words: ncclAllReduce on device scalars, stream-ordered.
// Parity-tested against the single-GPU path (positions, forces, charges, energies, species; incl. migration):
// sw_reaxff_b200/dist.py parity_check, executed by bench.py before anything is timed at N > 1.
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
  DBuf<double> rec_send;                 // [n][kRec] migration records of the local atoms
  DBuf<long long> flag, off;             // stayer flags and their scan
  DBuf<char> temp;
  // boundary exchange plan (rebuilt at every reneighbouring): my ghosts are grouped by source rank (goff), within a rank
  // in the sender's (atom, image) order; sendlist[soff[r] ..) = my atoms rank r holds as ghosts, in that same order
  std::vector<int> need_from, send_to, goff, soff;   // per peer: ghosts I need / atoms I send, and their offsets
  int nsend = 0;
  DBuf<int> greq, sendlist, cnt_d, cnt_all_d;
  DBuf<double> sendbuf, recvbuf;
  DBuf<double> dots_all;                 // [world][4]: every rank's partial CG dot products (dist_forward2_dots)
  DBuf<int> dest, cursor, emit_list;     // exchange/borders scratch: destination rank per atom, emission cursor and list
  DBuf<unsigned long long> keys, keys2;  // sort keys of the leavers / of the emitted (rank, atom, image) triples
  DBuf<double> mig_send, mig_recv, ghost_send, ghost_recv;
  size_t recv_bytes_last = 0;            // payload this rank received at the last reneighbouring (migrants + ghosts)
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
  // ---- host-planned halo (comm_init / comm_set_ghosts): the host owns the decomposition and the ghost ORDER, the plan
  // above (ghosts grouped by source rank) addresses them through a permutation
  bool external = false;
  int plan_nghost = -1;                  // ghost count the current plan was made for
  DBuf<int> gidx, gs_plan;               // plan slot -> ghost index (host order) / -> the ghost's sorted position
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
size_t System::slab() const { return 0; }
size_t System::dist_last_recv_bytes() const { return dist_ ? dist_->recv_bytes_last : 0; }

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

// ---- exchange + borders at reneighbouring: only migrants and boundary-shell atoms travel, and only between the ranks
// concerned (LAMMPS Comm::exchange / Comm::borders semantics; the reference gets both from the LAMMPS core over MPI).
//   exchange : every local atom is wrapped and assigned to the brick that contains it; the few that left travel as
//              144-byte records (x, v, q, s_hist, t_hist, tag, type) straight to their new owner.  New local order =
//              stayers in their old order, then arrivals by source rank: deterministic.
//   borders  : every rank decides itself which of its atoms (and which periodic images of them) lie in the ghost shell
//              of which rank - itself included - and sends 48-byte records (x, q, tag|type, image shift) there; the
//              ghosts of a rank are ordered by (source rank, source atom index, image).  The sender therefore already
//              holds the send lists of the per-step halos (forward x/q, CG direction, reverse f): no request round.
// Two small all-gathers (W ints per rank each) carry the counts; the payload moves in grouped ncclSend/ncclRecv.
namespace {
constexpr int kGRec = 6;  // doubles per ghost record: x y z q (tag,type) shiftcode

__device__ __forceinline__ int brick_of(double l, int g) {
  int b = (int)floor(l * g);
  return b < 0 ? 0 : (b >= g ? g - 1 : b);
}

// wrap into the box, pack the migration record, destination rank; leavers are counted per destination
__global__ void k_mig_classify(int n, BoxD b, int gx, int gy, int gz, int me, double4* __restrict__ xq,
                               const double* __restrict__ vel, const double* __restrict__ s_hist,
                               const double* __restrict__ t_hist, const int* __restrict__ tag, const int* __restrict__ ltype,
                               double* __restrict__ rec, int* __restrict__ dest, long long* __restrict__ stay_flag,
                               int* __restrict__ cnt_to) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { stay_flag[n] = 0; return; }
  double4 p = xq[i];
  double l[3];
  x2lamda(b, p.x, p.y, p.z, l);
  const int s0 = (int)floor(l[0]), s1 = (int)floor(l[1]), s2 = (int)floor(l[2]);
  if (s0 | s1 | s2) {
    double d[3];
    shift_vec(b, s0, s1, s2, d);
    p.x -= d[0]; p.y -= d[1]; p.z -= d[2];
    l[0] -= s0; l[1] -= s1; l[2] -= s2;
  }
  double* r = rec + (size_t)kRec * i;
  r[0] = p.x; r[1] = p.y; r[2] = p.z;
  r[3] = vel[3 * i]; r[4] = vel[3 * i + 1]; r[5] = vel[3 * i + 2];
  r[6] = p.w;
  for (int k = 0; k < 5; k++) { r[7 + k] = s_hist[5 * (size_t)i + k]; r[12 + k] = t_hist[5 * (size_t)i + k]; }
  r[17] = __longlong_as_double(((long long)tag[i] << 32) | (unsigned int)ltype[i]);
  const int d = (brick_of(l[2], gz) * gy + brick_of(l[1], gy)) * gx + brick_of(l[0], gx);
  dest[i] = d;
  stay_flag[i] = d == me ? 1 : 0;
  if (d != me) atomicAdd(&cnt_to[d], 1);
}

// leavers -> sort keys (destination, atom index): sorted, they are grouped by destination in atom order
__global__ void k_mig_keys(int n, int me, const int* __restrict__ dest, const long long* __restrict__ stay_scan,
                           unsigned long long* __restrict__ keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || dest[i] == me) return;
  const long long slot = i - stay_scan[i];                 // leavers before i
  keys[slot] = ((unsigned long long)dest[i] << 32) | (unsigned)i;
}
__global__ void k_mig_pack(int m, const unsigned long long* __restrict__ keys, const double* __restrict__ rec,
                           double* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const int i = (int)(keys[e] & 0xffffffffu);
  for (int k = 0; k < kRec; k++) out[(size_t)kRec * e + k] = rec[(size_t)kRec * i + k];
}
// new local atom a: a < nstay -> the a-th stayer (old index via the scan), else arrival a - nstay
__global__ void k_mig_fill(int n_old, int nstay, int narr, const long long* __restrict__ stay_flag,
                           const long long* __restrict__ stay_scan, const double* __restrict__ rec,
                           const double* __restrict__ arr, const int* __restrict__ map, int maplen, double4* __restrict__ xq,
                           double* __restrict__ vel, double* __restrict__ s_hist, double* __restrict__ t_hist,
                           int* __restrict__ tag, int* __restrict__ ltype, int* __restrict__ type) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  const double* r;
  long long a;
  if (t < n_old) {
    if (!stay_flag[t]) return;
    a = stay_scan[t];
    r = rec + (size_t)kRec * t;
  } else if (t < n_old + narr) {
    a = nstay + (t - n_old);
    r = arr + (size_t)kRec * (t - n_old);
  } else {
    return;
  }
  xq[a] = make_double4(r[0], r[1], r[2], r[6]);
  vel[3 * a] = r[3]; vel[3 * a + 1] = r[4]; vel[3 * a + 2] = r[5];
  for (int k = 0; k < 5; k++) { s_hist[5 * a + k] = r[7 + k]; t_hist[5 * a + k] = r[12 + k]; }
  const long long tt = __double_as_longlong(r[17]);
  const int lt = (int)(tt & 0xffffffffLL);
  tag[a] = (int)(tt >> 32); ltype[a] = lt;
  type[a] = (lt >= 1 && lt < maplen) ? map[lt] : -1;
}

// borders: (atom i, image shift) pairs inside the ghost shell of each rank.  COUNT: per-rank totals; FILL: sort keys
// (rank << 40 | atom << 5 | shift code) at an atomic cursor (sorted afterwards: deterministic order).
struct Shell { double cg[3]; int grid[3]; int m[3]; };
template <bool FILL>
__global__ void k_border_emit(int n, BoxD b, Shell S, int me, const double4* __restrict__ xq, int* __restrict__ cnt_to,
                              unsigned long long* __restrict__ keys, int* __restrict__ cursor) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = xq[i];
  double l[3];
  x2lamda(b, p.x, p.y, p.z, l);
  for (int sz = -S.m[2]; sz <= S.m[2]; sz++) {
    const double l2 = l[2] + sz;
    for (int bz = 0; bz < S.grid[2]; bz++) {
      const double lo2 = (double)bz / S.grid[2], hi2 = bz == S.grid[2] - 1 ? 1.0 : (double)(bz + 1) / S.grid[2];
      if (!(l2 >= lo2 - S.cg[2] && l2 < hi2 + S.cg[2])) continue;
      for (int sy = -S.m[1]; sy <= S.m[1]; sy++) {
        const double l1 = l[1] + sy;
        for (int by = 0; by < S.grid[1]; by++) {
          const double lo1 = (double)by / S.grid[1], hi1 = by == S.grid[1] - 1 ? 1.0 : (double)(by + 1) / S.grid[1];
          if (!(l1 >= lo1 - S.cg[1] && l1 < hi1 + S.cg[1])) continue;
          for (int sx = -S.m[0]; sx <= S.m[0]; sx++) {
            const double l0 = l[0] + sx;
            for (int bx = 0; bx < S.grid[0]; bx++) {
              const double lo0 = (double)bx / S.grid[0], hi0 = bx == S.grid[0] - 1 ? 1.0 : (double)(bx + 1) / S.grid[0];
              if (!(l0 >= lo0 - S.cg[0] && l0 < hi0 + S.cg[0])) continue;
              const int r = (bz * S.grid[1] + by) * S.grid[0] + bx;
              if (r == me && !sx && !sy && !sz) continue;          // the atom itself
              if (FILL) {
                const int code = ((sz + 2) * 5 + (sy + 2)) * 5 + (sx + 2);   // shifts in -2..2
                keys[atomicAdd(cursor, 1)] = ((unsigned long long)r << 40) | ((unsigned long long)(unsigned)i << 7) | (unsigned)code;
              } else {
                atomicAdd(&cnt_to[r], 1);
              }
            }
          }
        }
      }
    }
  }
}
__global__ void k_border_pack(int m, const unsigned long long* __restrict__ keys, const double4* __restrict__ xq,
                              const int* __restrict__ tag, const int* __restrict__ ltype, int* __restrict__ sendlist,
                              double* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const unsigned long long k = keys[e];
  const int i = (int)((k >> 7) & 0xffffffffULL), code = (int)(k & 127);
  sendlist[e] = i;
  const double4 p = xq[i];
  double* r = out + (size_t)kGRec * e;
  r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = p.w;
  r[4] = __longlong_as_double(((long long)tag[i] << 32) | (unsigned int)ltype[i]);
  r[5] = (double)code;
}
__global__ void k_border_unpack(int nghost, int n, BoxD b, const double* __restrict__ rec, const int* __restrict__ map, int maplen,
                                double4* __restrict__ xq, int* __restrict__ tag, int* __restrict__ ltype, int* __restrict__ type,
                                int* __restrict__ shift) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  const double* r = rec + (size_t)kGRec * g;
  const int code = (int)r[5];
  const int sx = code % 5 - 2, sy = (code / 5) % 5 - 2, sz = code / 25 - 2;
  double d[3];
  shift_vec(b, sx, sy, sz, d);
  xq[n + g] = make_double4(r[0] + d[0], r[1] + d[1], r[2] + d[2], r[3]);
  const long long tt = __double_as_longlong(r[4]);
  const int lt = (int)(tt & 0xffffffffLL);
  tag[n + g] = (int)(tt >> 32); ltype[n + g] = lt;
  type[n + g] = (lt >= 1 && lt < maplen) ? map[lt] : -1;
  shift[3 * g] = sx; shift[3 * g + 1] = sy; shift[3 * g + 2] = sz;
}
}  // namespace

// all-gather one row of W ints per rank -> host table[a * W + b] (what rank a reports about rank b)
static std::vector<int> gather_table(Dist& D, const int* row_dev, cudaStream_t st) {
  const int W = D.world;
  D.cnt_all_d.resize((size_t)W * W);
  RXB_NCCL(ncclAllGather(row_dev, D.cnt_all_d.p, W, ncclInt, D.comm, st));
  std::vector<int> all((size_t)W * W);
  RXB_CUDA(cudaMemcpyAsync(all.data(), D.cnt_all_d.p, all.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  RXB_SYNC(st);
  return all;
}

void System::dist_exchange() {
  Dist& D = *dist_;
  const int W = D.world, me = D.rank;
  BoxD b;
  memcpy(b.h, box.h, sizeof(b.h));
  memcpy(b.h_inv, box.h_inv, sizeof(b.h_inv));
  if (map_d.n != ff.map.size()) {
    map_d.resize(ff.map.size());
    RXB_CUDA(cudaMemcpyAsync(map_d.p, ff.map.data(), ff.map.size() * sizeof(int), cudaMemcpyHostToDevice, st_));
  }
  // ---------------- exchange: migrants to their new owners ----------------
  const int n_old = n;
  D.rec_send.resize((size_t)std::max(n_old, 1) * kRec);
  D.dest.resize(std::max(n_old, 1)); D.flag.resize(n_old + 1); D.off.resize(n_old + 1);
  D.cnt_d.resize(W);
  RXB_CUDA(cudaMemsetAsync(D.cnt_d.p, 0, W * sizeof(int), st_));
  k_mig_classify<<<nblk(n_old + 1), 256, 0, st_>>>(n_old, b, D.grid[0], D.grid[1], D.grid[2], me, xq.p, v_d.p, q_s_hist.p,
                                                  q_t_hist.p, tag.p, ltype_d.p, D.rec_send.p, D.dest.p, D.flag.p, D.cnt_d.p);
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, D.flag.p, D.off.p, n_old + 1, st_);
  D.temp.resize(need + 16);
  cub::DeviceScan::ExclusiveSum(D.temp.p, need, D.flag.p, D.off.p, n_old + 1, st_);
  const std::vector<int> mig = gather_table(D, D.cnt_d.p, st_);          // mig[a * W + b]: atoms leaving a for b
  int nleave = 0, narr = 0;
  std::vector<int> lv_off(W + 1, 0), ar_off(W + 1, 0);
  for (int r = 0; r < W; r++) {
    lv_off[r + 1] = lv_off[r] + mig[(size_t)me * W + r];
    ar_off[r + 1] = ar_off[r] + mig[(size_t)r * W + me];
  }
  nleave = lv_off[W]; narr = ar_off[W];
  const int nstay = n_old - nleave;
  D.mig_send.resize((size_t)std::max(nleave, 1) * kRec); D.mig_recv.resize((size_t)std::max(narr, 1) * kRec);
  if (nleave > 0) {
    D.keys.resize(nleave); D.keys2.resize(nleave);
    k_mig_keys<<<nblk(n_old), 256, 0, st_>>>(n_old, me, D.dest.p, D.off.p, D.keys.p);
    size_t ns = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, ns, D.keys.p, D.keys2.p, nleave, 0, 40, st_);
    D.temp.resize(ns + 16);
    cub::DeviceRadixSort::SortKeys(D.temp.p, ns, D.keys.p, D.keys2.p, nleave, 0, 40, st_);
    k_mig_pack<<<nblk(nleave), 256, 0, st_>>>(nleave, D.keys2.p, D.rec_send.p, D.mig_send.p);
  }
  if (nleave > 0 || narr > 0) {
    RXB_NCCL(ncclGroupStart());
    for (int r = 0; r < W; r++) {
      if (r == me) continue;
      const int ns_ = mig[(size_t)me * W + r], nr_ = mig[(size_t)r * W + me];
      if (ns_ > 0) RXB_NCCL(ncclSend(D.mig_send.p + (size_t)kRec * lv_off[r], (size_t)kRec * ns_, ncclDouble, r, D.comm, st_));
      if (nr_ > 0) RXB_NCCL(ncclRecv(D.mig_recv.p + (size_t)kRec * ar_off[r], (size_t)kRec * nr_, ncclDouble, r, D.comm, st_));
    }
    RXB_NCCL(ncclGroupEnd());
  }
  n = nstay + narr;
  N = n;
  ensure_atom_capacity();
  xq.resize_keep((size_t)std::max(n, 1)); xq.n = n;
  tag.resize_keep((size_t)std::max(n, 1)); ltype_d.resize_keep((size_t)std::max(n, 1)); type.resize_keep((size_t)std::max(n, 1));
  v_d.resize_keep((size_t)3 * std::max(n, 1));
  q_s_hist.resize_keep((size_t)5 * std::max(n, 1)); q_t_hist.resize_keep((size_t)5 * std::max(n, 1));
  // (stayers only move DOWN in the arrays and every source record was packed into rec_send first: in-place is safe)
  k_mig_fill<<<nblk(n_old + narr), 256, 0, st_>>>(n_old, nstay, narr, D.flag.p, D.off.p, D.rec_send.p, D.mig_recv.p, map_d.p,
                                                 (int)ff.map.size(), xq.p, v_d.p, q_s_hist.p, q_t_hist.p, tag.p, ltype_d.p, type.p);
  q_s_hist.n = (size_t)5 * n; q_t_hist.n = (size_t)5 * n;

  // ---------------- borders: ghost shells ----------------
  Shell S;
  const double cut = cutneigh();
  S.cg[0] = cut * sqrt(b.h_inv[0] * b.h_inv[0] + b.h_inv[5] * b.h_inv[5] + b.h_inv[4] * b.h_inv[4]);
  S.cg[1] = cut * sqrt(b.h_inv[1] * b.h_inv[1] + b.h_inv[3] * b.h_inv[3]);
  S.cg[2] = cut * b.h_inv[2];
  for (int t = 0; t < 3; t++) {
    S.grid[t] = D.grid[t];
    S.m[t] = (int)ceil(S.cg[t]);
    if (S.m[t] > 2) throw std::runtime_error("rxb dist: the box is thinner than half the ghost cut-off in one direction");
  }
  RXB_CUDA(cudaMemsetAsync(D.cnt_d.p, 0, W * sizeof(int), st_));
  k_border_emit<false><<<nblk(n), 256, 0, st_>>>(n, b, S, me, xq.p, D.cnt_d.p, nullptr, nullptr);
  const std::vector<int> all = gather_table(D, D.cnt_d.p, st_);           // all[a * W + b]: ghosts a sends to b
  D.need_from.assign(W, 0); D.send_to.assign(W, 0); D.goff.assign(W + 1, 0); D.soff.assign(W + 1, 0);
  std::vector<int> emit_off(W + 1, 0);                                    // offsets in the sorted emission list (self included)
  for (int r = 0; r < W; r++) {
    D.need_from[r] = all[(size_t)r * W + me];
    D.send_to[r] = (r == me) ? 0 : all[(size_t)me * W + r];
    D.goff[r + 1] = D.goff[r] + D.need_from[r];
    D.soff[r + 1] = D.soff[r] + D.send_to[r];
    emit_off[r + 1] = emit_off[r] + all[(size_t)me * W + r];
  }
  const int nemit = emit_off[W], nself = all[(size_t)me * W + me], nghost = D.goff[W];
  D.nsend = D.soff[W];
  D.keys.resize(std::max(nemit, 1)); D.keys2.resize(std::max(nemit, 1));
  D.cursor.resize(1);
  RXB_CUDA(cudaMemsetAsync(D.cursor.p, 0, sizeof(int), st_));
  k_border_emit<true><<<nblk(n), 256, 0, st_>>>(n, b, S, me, xq.p, nullptr, D.keys.p, D.cursor.p);
  if (nemit > 0) {
    size_t ns = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, ns, D.keys.p, D.keys2.p, nemit, 0, 48, st_);
    D.temp.resize(ns + 16);
    cub::DeviceRadixSort::SortKeys(D.temp.p, ns, D.keys.p, D.keys2.p, nemit, 0, 48, st_);
  }
  // pack every emitted pair (sorted: by rank, atom, image); emit_list[e] = source atom of emission e
  D.emit_list.resize(std::max(nemit, 1));
  D.ghost_send.resize((size_t)kGRec * std::max(nemit, 1)); D.ghost_recv.resize((size_t)kGRec * std::max(nghost, 1));
  if (nemit > 0) k_border_pack<<<nblk(nemit), 256, 0, st_>>>(nemit, D.keys2.p, xq.p, tag.p, ltype_d.p, D.emit_list.p, D.ghost_send.p);
  RXB_NCCL(ncclGroupStart());
  for (int r = 0; r < W; r++) {
    if (r == me) continue;
    if (D.send_to[r] > 0) RXB_NCCL(ncclSend(D.ghost_send.p + (size_t)kGRec * emit_off[r], (size_t)kGRec * D.send_to[r], ncclDouble, r, D.comm, st_));
    if (D.need_from[r] > 0) RXB_NCCL(ncclRecv(D.ghost_recv.p + (size_t)kGRec * D.goff[r], (size_t)kGRec * D.need_from[r], ncclDouble, r, D.comm, st_));
  }
  RXB_NCCL(ncclGroupEnd());
  if (nself > 0)
    RXB_CUDA(cudaMemcpyAsync(D.ghost_recv.p + (size_t)kGRec * D.goff[me], D.ghost_send.p + (size_t)kGRec * emit_off[me],
                             (size_t)kGRec * nself * sizeof(double), cudaMemcpyDeviceToDevice, st_));
  // send lists of the per-step halos: the emission list without the self segment; greq = sources of my own images
  D.sendlist.resize(std::max(D.nsend, 1));
  D.greq.resize(std::max(nghost, 1));
  for (int r = 0; r < W; r++) {
    const int cnt = all[(size_t)me * W + r];
    if (cnt == 0) continue;
    if (r == me) RXB_CUDA(cudaMemcpyAsync(D.greq.p + D.goff[me], D.emit_list.p + emit_off[r], (size_t)cnt * sizeof(int), cudaMemcpyDeviceToDevice, st_));
    else RXB_CUDA(cudaMemcpyAsync(D.sendlist.p + D.soff[r], D.emit_list.p + emit_off[r], (size_t)cnt * sizeof(int), cudaMemcpyDeviceToDevice, st_));
  }
  N = n + nghost;
  ensure_atom_capacity();
  ghost_shift.resize(std::max<size_t>((size_t)3 * nghost, 3));
  ghost_owner.resize(std::max<size_t>(nghost, 1));
  if (nghost > 0)
    k_border_unpack<<<nblk(nghost), 256, 0, st_>>>(nghost, n, b, D.ghost_recv.p, map_d.p, (int)ff.map.size(), xq.p, tag.p, ltype_d.p,
                                                  type.p, ghost_shift.p);
  D.sendbuf.resize((size_t)4 * std::max(D.nsend, 1));
  D.recvbuf.resize((size_t)3 * std::max(D.nsend, 1));
  // peer exchange: where my values land in each consumer's halo (its ghost offset for source = me), and whether every
  // rank's ghosts fit the halo windows (decided from the all-gathered table: the same answer on every rank)
  {
    std::vector<int> dst_off(W, 0);
    long long max_ghosts = 0;
    for (int p = 0; p < W; p++) {
      long long tot = 0;
      for (int r = 0; r < W; r++) {
        if (r == me) dst_off[p] = (int)tot;
        tot += all[(size_t)r * W + p];
      }
      max_ghosts = std::max(max_ghosts, tot);
    }
    D.peer_plan_ok = D.peer_ok && max_ghosts <= (long long)D.cap_g;
    D.dst_off_d.resize(W); D.soff_d.resize(W + 1);
    RXB_CUDA(cudaMemcpyAsync(D.dst_off_d.p, dst_off.data(), W * sizeof(int), cudaMemcpyHostToDevice, st_));
    RXB_CUDA(cudaMemcpyAsync(D.soff_d.p, D.soff.data(), (W + 1) * sizeof(int), cudaMemcpyHostToDevice, st_));
    RXB_SYNC(st_);          // dst_off is a local that dies with this scope
  }
  D.recv_bytes_last = ((size_t)narr * kRec + (size_t)(nghost - nself) * kGRec) * sizeof(double);
  kernel_launches += 10;
  RXB_CUDA(cudaGetLastError());
}

namespace {
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

// ---- host-planned halo --------------------------------------------------------------------------------------------
// The reference runs one MPI rank per core group and lets LAMMPS' Comm move ghosts (forward_comm_fix inside the CG loop,
// fix_qeq_reax_sunway.cpp:1108-1140).  Here a multi-rank host keeps that decomposition and its own ghost order; what the
// library needs to run the CG halo on the device is, per ghost, the owning rank and the atom's local index there - one
// forward communication of (me, i) after borders() on the LAMMPS side.  From that every rank derives the
// same direct plan the brick decomposition uses (ghosts grouped by source rank, senders' lists from a request exchange),
// so the NVLink peer exchange and the NCCL fallback are shared; only the ghost order goes through a permutation.
void System::comm_init(int rank, int world, const char* id128) {
  RXB_CUDA(cudaSetDevice(device_));
  if (dist_) throw std::runtime_error("rxb_comm_init: this handle already belongs to a communicator");
  if (rank < 0 || rank >= world) throw std::runtime_error("rxb_comm_init: rank outside [0, world)");
  dist_ = new Dist();
  Dist& D = *dist_;
  D.rank = rank; D.world = world; D.external = true;
  D.grid[0] = D.grid[1] = D.grid[2] = 0;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  RXB_NCCL(ncclCommInitRank(&D.comm, world, id, rank));
  dist_peer_setup();
}

void System::comm_set_ghosts(int nghost, const int* owner_rank, const int* owner_index) {
  RXB_CUDA(cudaSetDevice(device_));
  if (!dist_ || !dist_->external) throw std::runtime_error("rxb_comm_set_ghosts: call rxb_comm_init first");
  if (nghost != N - n)
    throw std::runtime_error("rxb_comm_set_ghosts: nghost = " + std::to_string(nghost) + " but rxb_set_atoms was given " +
                             std::to_string(N - n) + " ghosts");
  if (nghost > 0 && (!owner_rank || !owner_index)) throw std::runtime_error("rxb_comm_set_ghosts: null owner arrays");
  if (n == 0) throw std::runtime_error("rxb_comm_set_ghosts: a rank without local atoms cannot take part in the solve (rebalance the decomposition)");
  Dist& D = *dist_;
  const int W = D.world, me = D.rank;
  // ghosts grouped by owning rank, host order kept inside a group
  D.need_from.assign(W, 0); D.send_to.assign(W, 0); D.goff.assign(W + 1, 0); D.soff.assign(W + 1, 0);
  for (int g = 0; g < nghost; g++) {
    const int r = owner_rank[g];
    if (r < 0 || r >= W) throw std::runtime_error("rxb_comm_set_ghosts: owner rank " + std::to_string(r) + " of ghost " + std::to_string(g) + " outside [0, world)");
    if (r == me && (owner_index[g] < 0 || owner_index[g] >= n))
      throw std::runtime_error("rxb_comm_set_ghosts: owner index of ghost " + std::to_string(g) + " is not a local atom");
    D.need_from[r]++;
  }
  for (int r = 0; r < W; r++) D.goff[r + 1] = D.goff[r] + D.need_from[r];
  std::vector<int> gidx(std::max(nghost, 1), 0), req(std::max(nghost, 1), 0), cur(D.goff.begin(), D.goff.end() - 1);
  for (int g = 0; g < nghost; g++) {
    const int p = cur[owner_rank[g]]++;
    gidx[p] = g; req[p] = owner_index[g];
  }
  // who needs how many from whom (the same table on every rank)
  D.cnt_d.resize(W);
  RXB_CUDA(cudaMemcpyAsync(D.cnt_d.p, D.need_from.data(), W * sizeof(int), cudaMemcpyHostToDevice, st_));
  const std::vector<int> need = gather_table(D, D.cnt_d.p, st_);            // need[a * W + b]: ghosts of a owned by b
  for (int r = 0; r < W; r++) {
    D.send_to[r] = (r == me) ? 0 : need[(size_t)r * W + me];
    D.soff[r + 1] = D.soff[r] + D.send_to[r];
  }
  D.nsend = D.soff[W];
  // request exchange: the owner indices of my ghosts go to their owners and come back as my send lists
  D.greq.resize(std::max(nghost, 1)); D.gidx.resize(std::max(nghost, 1)); D.sendlist.resize(std::max(D.nsend, 1));
  RXB_CUDA(cudaMemcpyAsync(D.greq.p, req.data(), (size_t)std::max(nghost, 1) * sizeof(int), cudaMemcpyHostToDevice, st_));
  RXB_CUDA(cudaMemcpyAsync(D.gidx.p, gidx.data(), (size_t)std::max(nghost, 1) * sizeof(int), cudaMemcpyHostToDevice, st_));
  RXB_NCCL(ncclGroupStart());
  for (int r = 0; r < W; r++) {
    if (r == me) continue;
    if (D.need_from[r] > 0) RXB_NCCL(ncclSend(D.greq.p + D.goff[r], (size_t)D.need_from[r], ncclInt, r, D.comm, st_));
    if (D.send_to[r] > 0) RXB_NCCL(ncclRecv(D.sendlist.p + D.soff[r], (size_t)D.send_to[r], ncclInt, r, D.comm, st_));
  }
  RXB_NCCL(ncclGroupEnd());
  std::vector<int> sl(std::max(D.nsend, 1), 0);
  RXB_CUDA(cudaMemcpyAsync(sl.data(), D.sendlist.p, (size_t)D.nsend * sizeof(int), cudaMemcpyDeviceToHost, st_));
  D.sendbuf.resize((size_t)4 * std::max(D.nsend, 1));
  D.recvbuf.resize((size_t)3 * std::max(D.nsend, 1));
  // peer exchange: where my values land in each consumer's halo, and whether every rank's ghosts fit the windows
  std::vector<int> dst_off(W, 0);
  long long max_ghosts = 0;
  for (int p = 0; p < W; p++) {
    long long tot = 0;
    for (int r = 0; r < W; r++) {
      if (r == me) dst_off[p] = (int)tot;
      tot += need[(size_t)p * W + r];
    }
    max_ghosts = std::max(max_ghosts, tot);
  }
  D.peer_plan_ok = D.peer_ok && max_ghosts <= (long long)D.cap_g;
  D.dst_off_d.resize(W); D.soff_d.resize(W + 1);
  RXB_CUDA(cudaMemcpyAsync(D.dst_off_d.p, dst_off.data(), W * sizeof(int), cudaMemcpyHostToDevice, st_));
  RXB_CUDA(cudaMemcpyAsync(D.soff_d.p, D.soff.data(), (W + 1) * sizeof(int), cudaMemcpyHostToDevice, st_));
  RXB_SYNC(st_);
  for (int e = 0; e < D.nsend; e++)
    if (sl[e] < 0 || sl[e] >= n)
      throw std::runtime_error("rxb_comm_set_ghosts: a peer asked for local index " + std::to_string(sl[e]) + " but this rank has " +
                               std::to_string(n) + " local atoms (owner_index must be the index on the OWNING rank)");
  D.plan_nghost = nghost;
  D.recv_bytes_last = (size_t)D.nsend * sizeof(int);
  s2a.n = 0;                             // the sorted maps of the plan are made by the next rxb_neigh_build
}

namespace {
__global__ void k_pack_q(int m, const int* __restrict__ list, const double4* __restrict__ xq, double* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < m) out[e] = xq[list[e]].w;
}
__global__ void k_unpack_q(int nghost, int n, int g0, int g1, const int* __restrict__ gidx, const int* __restrict__ greq,
                           const double* __restrict__ recv, double4* __restrict__ xq) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nghost) return;
  xq[n + gidx[p]].w = (p >= g0 && p < g1) ? xq[greq[p]].w : recv[p];
}
}  // namespace

// host-planned halo: the new charges of the real atoms -> their ghosts on every rank (the reference ends pre_force with
// comm->forward_comm_fix(this) of q, fix_qeq_reax_sunway.cpp:1290-1298); positions are the host's business in this mode
static void ext_forward_q(System& s, Dist& D, cudaStream_t st) {
  const int W = D.world, nghost = s.N - s.n;
  if (D.nsend > 0) k_pack_q<<<nblk(D.nsend), 256, 0, st>>>(D.nsend, D.sendlist.p, s.xq.p, D.sendbuf.p);
  RXB_NCCL(ncclGroupStart());
  for (int r = 0; r < W; r++) {
    if (r == D.rank) continue;
    if (D.send_to[r] > 0) RXB_NCCL(ncclSend(D.sendbuf.p + D.soff[r], (size_t)D.send_to[r], ncclDouble, r, D.comm, st));
    if (D.need_from[r] > 0) RXB_NCCL(ncclRecv(D.recv2.p + D.goff[r], (size_t)D.need_from[r], ncclDouble, r, D.comm, st));
  }
  RXB_NCCL(ncclGroupEnd());
  if (nghost > 0)
    k_unpack_q<<<nblk(nghost), 256, 0, st>>>(nghost, s.n, D.goff[D.rank], D.goff[D.rank + 1], D.gidx.p, D.greq.p, D.recv2.p, s.xq.p);
  s.kernel_launches += 2;
}

void System::dist_forward_xq() {
  Dist& D = *dist_;
  if (D.external) { ext_forward_q(*this, D, st_); return; }
  BoxD b;
  memcpy(b.h, box.h, sizeof(b.h));
  memcpy(b.h_inv, box.h_inv, sizeof(b.h_inv));
  const int nghost = N - n;
  p2p_forward<4>(*this, D, reinterpret_cast<double*>(xq.p), n, st_);
  if (nghost > 0) k_shift_ghosts<<<nblk(nghost), 256, 0, st_>>>(n, nghost, b, ghost_shift.p, xq.p);
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
  pdl_wait(); pdl_release();
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
  // (A/B at N = 2 and N = 8: one fence per CTA after the barrier instead of one per thread, and W lanes polling with
  // volatile loads + a fence instead of one thread with ld.acquire, made the exchange 12 - 18 us per iteration SLOWER)
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
  pdl_wait(); pdl_release();
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

// vec: S-space double2 vector whose ghosts are refreshed (null: pure reduction); dots[ndots]: partial sums -> totals.
// push and pull are separate calls so that work which does not need the ghosts can be enqueued between them.
void peer_push(System& s, Dist& D, double2* vec, const int* gs_pos, double* dots, int ndots, cudaStream_t st) {
  D.seq++;
  const int par = (int)(D.seq & 1);
  const PeerView P = peer_view(D);
  const int g0 = vec ? D.goff[D.rank] : 0, g1 = vec ? D.goff[D.rank + 1] : 0;
  const int nsend = vec ? D.nsend : 0;
  const int work = std::max(nsend, g1 - g0);
  // (plain launches: with programmatic dependent launch the pull CTAs sit resident while the push drains, measured
  // 0.06 ms/step slower at N = 2; the kernels' pdl_wait() is a no-op then)
  k_peer_push<<<std::max(1, std::min(148, (work + 255) / 256)), 256, 0, st>>>(P, par, D.seq, nsend, D.send_s.p, D.soff_d.p,
                                                                           D.dst_off_d.p, vec, g0, g1, gs_pos, D.self_s.p,
                                                                           ndots, dots, D.done_d.p);
  s.kernel_launches++;
}
void peer_pull(System& s, Dist& D, double2* vec, int nghost, const int* gs_pos, double* dots, int ndots, cudaStream_t st) {
  const int par = (int)(D.seq & 1);
  const PeerView P = peer_view(D);
  const int g0 = vec ? D.goff[D.rank] : 0, g1 = vec ? D.goff[D.rank + 1] : 0;
  const int ng = vec ? nghost : 0;
  k_peer_pull<<<std::max(1, std::min(296, (ng + 255) / 256)), 256, 0, st>>>(P, par, D.seq, ng, g0, g1, gs_pos, vec, ndots, dots,
                                                                         D.peer_err_d.p);
  s.kernel_launches++;
}
void peer_exchange(System& s, Dist& D, double2* vec, int nghost, const int* gs_pos, double* dots, int ndots, cudaStream_t st) {
  peer_push(s, D, vec, gs_pos, dots, ndots, st);
  peer_pull(s, D, vec, nghost, gs_pos, dots, ndots, st);
}

// interior rows (every column local or an own periodic image for the whole reneighbouring interval) first, boundary rows
// after: a row is interior when its atom sits at least the ghost cut-off inside every face the brick shares with ANOTHER rank
__global__ void k_row_class(int n, BoxD b, double lo0, double lo1, double lo2, double hi0, double hi1, double hi2, double c0,
                            double c1, double c2, int m0, int m1, int m2, const int* __restrict__ rowpos,
                            const double4* __restrict__ xqs, long long* __restrict__ flag) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  if (r == n) { flag[n] = 0; return; }
  const double4 p = xqs[rowpos[r]];
  double l[3];
  x2lamda(b, p.x, p.y, p.z, l);
  const bool in0 = !m0 || (l[0] - lo0 >= c0 && hi0 - l[0] >= c0);
  const bool in1 = !m1 || (l[1] - lo1 >= c1 && hi1 - l[1] >= c1);
  const bool in2 = !m2 || (l[2] - lo2 >= c2 && hi2 - l[2] >= c2);
  flag[r] = (in0 && in1 && in2) ? 1 : 0;
}
__global__ void k_row_order(int n, long long n_int, const long long* __restrict__ flag, const long long* __restrict__ scan,
                            int* __restrict__ rowlist) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const long long before = scan[r];                       // interior rows before r
  rowlist[flag[r] ? before : n_int + (r - before)] = r;
}
}  // namespace

void System::dist_classify_rows() {
  Dist& D = *dist_;
  n_interior_ = 0;
  if (!(D.peer_ok && D.peer_plan_ok) || n == 0) return;
  BoxD b;
  memcpy(b.h, box.h, sizeof(b.h));
  memcpy(b.h_inv, box.h_inv, sizeof(b.h_inv));
  const double cut = cutneigh();
  const double c0 = cut * sqrt(b.h_inv[0] * b.h_inv[0] + b.h_inv[5] * b.h_inv[5] + b.h_inv[4] * b.h_inv[4]);
  const double c1 = cut * sqrt(b.h_inv[1] * b.h_inv[1] + b.h_inv[3] * b.h_inv[3]);
  const double c2 = cut * b.h_inv[2];
  D.flag.resize(n + 1); D.off.resize(n + 1);
  q_rowlist.resize(n);
  // a dimension with one brick has only this rank's own periodic images beyond its faces: no constraint there
  k_row_class<<<nblk(n + 1), 256, 0, st_>>>(n, b, D.lo[0], D.lo[1], D.lo[2], D.hi[0], D.hi[1], D.hi[2], c0, c1, c2, D.grid[0] > 1,
                                           D.grid[1] > 1, D.grid[2] > 1, rowpos.p, xqs.p, D.flag.p);
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, D.flag.p, D.off.p, n + 1, st_);
  D.temp.resize(need + 16);
  cub::DeviceScan::ExclusiveSum(D.temp.p, need, D.flag.p, D.off.p, n + 1, st_);
  long long n_int = 0;
  RXB_CUDA(cudaMemcpyAsync(&n_int, D.off.p + n, sizeof(long long), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
  k_row_order<<<nblk(n), 256, 0, st_>>>(n, n_int, D.flag.p, D.off.p, q_rowlist.p);
  // every rank takes the split path or none does not matter for correctness (push/pull are the same calls either way);
  // a split with a tiny interior is not worth its extra launch
  n_interior_ = n_int * 20 >= n ? (int)n_int : 0;
  kernel_launches += 3;
}

void System::dist_push2(double2* vecS, double* dots, int ndots) { peer_push(*this, *dist_, vecS, dist_gs(), dots, ndots, st_); }
void System::dist_pull2(double2* vecS, double* dots, int ndots) {
  peer_pull(*this, *dist_, vecS, N - n, dist_gs(), dots, ndots, st_);
}

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
  RXB_SYNC(st_);
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
  RXB_SYNC(st_);
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
  if (D.external) {
    if (D.plan_nghost != nghost)
      throw std::runtime_error("rxb_comm_set_ghosts must follow every rxb_set_atoms (the ghost plan is for " +
                               std::to_string(D.plan_nghost) + " ghosts, the atom set has " + std::to_string(nghost) + ")");
    D.gs_plan.resize(std::max(nghost, 1));
    if (nghost > 0) { k_map_idx<<<nblk(nghost), 256, 0, st_>>>(nghost, D.gidx.p, gs_pos.p, D.gs_plan.p); kernel_launches++; }
  }
}
const int* System::dist_gs() const { return dist_ && dist_->external ? dist_->gs_plan.p : gs_pos.p; }
bool System::dist_external() const { return dist_ && dist_->external; }

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
  s_forward2(*this, *dist_, vec, n, N - n, dist_gs(), nullptr, st_);
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
  s_forward2(*this, D, vec, n, N - n, dist_gs(), dots, st_);
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

// a wait of the peer exchange timed out (a peer process died): the flag travels with the end-of-step status (read_step_status)
const int* System::dist_peer_err_ptr() const { return dist_ && dist_->peer_ok ? dist_->peer_err_d.p : nullptr; }

void System::dist_reverse_f() {
  Dist& D = *dist_;
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
}

// (kept for ABI compatibility: the whole-slab all-gather transport of round 1 is gone, halos are always point to point)
void System::dist_set_p2p(bool) {}

}  // namespace rxb
