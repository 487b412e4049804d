// Cell-sorted full neighbour-list build on the GPU -> device CSR.
//
// Replaces NPairFullBin{Atomonly,Ghost}Sunway::build and its slave kernels
// (/root/reference/npair_full_bin_atomonly_sunway.cpp:39-198, npair_full_bin_ghost_sw5.c:80-228):
// same list semantics (full list, j != i, r^2 <= cut^2 in fp64, rows for ghost atoms when asked),
// different machinery: atoms are radix-sorted by cell (x fastest) so that every (y,z) stencil row is ONE
// contiguous run of the sorted position array; a warp owns a row atom, sweeps the runs with coalesced
// 16-byte loads and ballot-compacts the hits.  Rows sit at a fixed stride (one pass, no count pass).  The Verlet list is
// emitted in "S space": rows = local atoms in cell-sorted order, columns = sorted positions (rxb_dev.cuh).
//
// HBM-bound by design (28(N+G) read + 4 nnz written); the stencil re-reads hit L1/L2.
#include <cub/cub.cuh>

#include <algorithm>

#include "rxb_system.h"

namespace rxb {

namespace {

constexpr double kSlack = 1e-6;  // Angstrom
constexpr int kXFine = 2;        // x bins are bin_size / kXFine wide (x is the contiguous direction of the sorted order)

struct Grid {
  double lo[3], inv[3], size[3];
  int nb[3];
};

__global__ void k_bounds(const double4* __restrict__ xq, int N, double* __restrict__ out6) {
  // out6 = min xyz, max xyz; atomics on order-preserving integer images of the doubles
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    double4 p = xq[i];
    mn[0] = fmin(mn[0], p.x); mn[1] = fmin(mn[1], p.y); mn[2] = fmin(mn[2], p.z);
    mx[0] = fmax(mx[0], p.x); mx[1] = fmax(mx[1], p.y); mx[2] = fmax(mx[2], p.z);
  }
  for (int t = 0; t < 3; t++) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[t] = fmin(mn[t], __shfl_xor_sync(0xffffffffu, mn[t], o));
      mx[t] = fmax(mx[t], __shfl_xor_sync(0xffffffffu, mx[t], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
    auto enc = [](double d) { long long b = __double_as_longlong(d); return b >= 0 ? b : b ^ 0x7fffffffffffffffLL; };
    for (int t = 0; t < 3; t++) {
      atomicMin((long long*)out6 + t, enc(mn[t]));
      atomicMax((long long*)out6 + 3 + t, enc(mx[t]));
    }
  }
}

__device__ __forceinline__ int bin_coord(double x, double lo, double inv, int nb) {
  int b = (int)((x - lo) * inv);
  return min(max(b, 0), nb - 1);
}

__global__ void k_bin_ids(const double4* __restrict__ xq, int N, Grid g, int* __restrict__ key, int* __restrict__ val,
                          int* __restrict__ bin_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double4 p = xq[i];
  int bx = bin_coord(p.x, g.lo[0], g.inv[0], g.nb[0]);
  int by = bin_coord(p.y, g.lo[1], g.inv[1], g.nb[1]);
  int bz = bin_coord(p.z, g.lo[2], g.inv[2], g.nb[2]);
  int id = (bz * g.nb[1] + by) * g.nb[0] + bx;
  key[i] = id;
  val[i] = i;
  atomicAdd(&bin_count[id], 1);
}

// sorted copies: exact fp64 positions (spos) and an fp32 shadow relative to the grid origin with the atom index in .w
// (sposf).  The stencil sweep reads the 16-byte shadow; only pairs whose fp32 distance falls inside the rounding band
// around the cut-off touch the 32-byte exact record, so the list is still decided by the oracle's fp64 arithmetic.
__global__ void k_gather_sorted(const double4* __restrict__ xq, const int* __restrict__ sorted_idx, int N, Grid g,
                                double4* __restrict__ spos, float4* __restrict__ sposf) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  int i = sorted_idx[k];
  double4 p = xq[i];
  sposf[k] = make_float4((float)(p.x - g.lo[0]), (float)(p.y - g.lo[1]), (float)(p.z - g.lo[2]), __int_as_float(i));
  p.w = __longlong_as_double((long long)i);
  spos[k] = p;
}

// One warp per row.  FILL = false: count only.  FILL = true: write row i at i * stride (entries beyond the stride are dropped
// but still counted, the host then re-runs with a larger stride) and store the row length.
// SORTED = false: row i is atom i (position xq[i]) and the columns are atom indices.
// SORTED = true : row i is the i-th LOCAL atom in cell-sorted order, sitting at sorted position rowpos[i]; the columns are
//                 sorted positions too ("S space", rxb_dev.cuh), so a row's neighbours are runs of consecutive integers and
//                 every later gather through this list (shadow positions, charges, CG vectors) touches whole cache lines.
template <bool FILL, bool SORTED>
__global__ void __launch_bounds__(256)
k_build(const double4* __restrict__ xq, const int* __restrict__ rowpos, const double4* __restrict__ spos,
        const float4* __restrict__ sposf, const int* __restrict__ bin_start, Grid g, int nrows, double cut, float band,
        float in2hi, int reach, int* __restrict__ cnt, int* __restrict__ cnt_in, long long* __restrict__ off,
        int* __restrict__ idx, int stride) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= nrows) return;
  const int self = SORTED ? rowpos[i] : i;          // the row's own identity in the column index space
  const double4 pi = SORTED ? spos[self] : xq[i];
  const double c2 = cut * cut;
  const float fx = (float)(pi.x - g.lo[0]), fy = (float)(pi.y - g.lo[1]), fz = (float)(pi.z - g.lo[2]);
  const float c2lo = (float)c2 - band, c2hi = (float)c2 + band;
  const int bx = bin_coord(pi.x, g.lo[0], g.inv[0], g.nb[0]);
  const int by = bin_coord(pi.y, g.lo[1], g.inv[1], g.nb[1]);
  const int bz = bin_coord(pi.z, g.lo[2], g.inv[2], g.nb[2]);
  const long long base_off = (long long)i * stride;
  // inner entries (fp32 r^2 <= in2hi, a superset of r <= cut_in) fill the row from the front, the others from the back;
  // the back block is moved down behind the inner one at the end, so the row stays contiguous
  int n_in = 0, n_out = 0;
  const unsigned lt = (1u << lane) - 1;
  // The (2 reach + 1)^2 bin rows (cz, cy) of the stencil: lane t works out the x run [kbeg, kend) of row t in the sorted
  // order once (fp64 chord, two floors, two bin_start loads), then the warp walks the runs in (cz, cy) order.  Doing this
  // per row inside the loops cost as many instructions as the candidate tests themselves (ncu: issue slots 78 % busy).
  const int span = 2 * reach + 1, nrun = span * span;
  for (int base = 0; base < nrun; base += 32) {
    int kbeg_l = 0, kend_l = 0;
    const int t_l = base + lane;
    if (t_l < nrun) {
      const int cz = bz - reach + t_l / span, cy = by - reach + t_l % span;
      if (cz >= 0 && cz < g.nb[2] && cy >= 0 && cy < g.nb[1]) {
        double dz = 0.0, dy = 0.0;
        if (cz > bz) dz = (g.lo[2] + cz * g.size[2]) - pi.z;
        else if (cz < bz) dz = pi.z - (g.lo[2] + (cz + 1) * g.size[2]);
        dz = fmax(dz - kSlack, 0.0);  // bin edges are rounded; never prune a run that could hold a boundary pair
        if (cy > by) dy = (g.lo[1] + cy * g.size[1]) - pi.y;
        else if (cy < by) dy = pi.y - (g.lo[1] + (cy + 1) * g.size[1]);
        dy = fmax(dy - kSlack, 0.0);
        const double rem = c2 - dz * dz - dy * dy;
        if (dz * dz <= c2 && rem >= 0.0) {
          // trim the x run to the chord of the cutoff sphere.  bin_coord is monotonic in x, so the bins of x_i -+ half
          // (same arithmetic, kSlack = 1e-6 A against the ulp-level rounding of the edges) bracket every atom inside the
          // chord; the x bins are kXFine times finer than the y/z bins, so the run overshoots the chord by half a coarse
          // bin on average
          const double half = sqrt(rem) + kSlack;
          int x0 = (int)floor((pi.x - half - g.lo[0]) * g.inv[0]);
          int x1 = (int)floor((pi.x + half - g.lo[0]) * g.inv[0]);
          x0 = max(max(x0, bx - kXFine * reach), 0);
          x1 = min(min(x1, bx + kXFine * reach), g.nb[0] - 1);
          if (x1 >= x0) {
            const int rowbase = (cz * g.nb[1] + cy) * g.nb[0];
            kbeg_l = bin_start[rowbase + x0]; kend_l = bin_start[rowbase + x1 + 1];
          }
        }
      }
    }
    const int nhere = min(32, nrun - base);
    for (int t = 0; t < nhere; t++) {
      const int kbeg = __shfl_sync(0xffffffffu, kbeg_l, t), kend = __shfl_sync(0xffffffffu, kend_l, t);
      for (int k0 = kbeg; k0 < kend; k0 += 32) {
        const int k = k0 + lane;
        bool hit = false, inner = false;
        int j = -1;
        if (k < kend) {
          const float4 qj = sposf[k];
          j = SORTED ? k : __float_as_int(qj.w);
          const float ex = qj.x - fx, ey = qj.y - fy, ez = qj.z - fz;
          const float r2f = ex * ex + ey * ey + ez * ez;
          inner = r2f <= in2hi;
          if (r2f < c2lo) hit = (j != self);
          else if (r2f <= c2hi) {
            // inside the fp32 rounding band: decide with the exact record.  Explicit rn ops: no FMA contraction, so the
            // r^2 <= cut^2 test is the oracle's arithmetic bit for bit
            const double4 pj = spos[k];
            const double ddx = __dsub_rn(pi.x, pj.x), ddy = __dsub_rn(pi.y, pj.y), ddz = __dsub_rn(pi.z, pj.z);
            const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)), __dmul_rn(ddz, ddz));
            hit = (j != self) && (r2 <= c2);
          }
        }
        const unsigned m_in = __ballot_sync(0xffffffffu, hit && inner), m_out = __ballot_sync(0xffffffffu, hit && !inner);
        if (FILL && hit) {
          if (inner) {
            const int pos = n_in + __popc(m_in & lt);
            if (pos < stride) idx[base_off + pos] = j;
          } else {
            const int pos = n_out + __popc(m_out & lt);
            if (pos < stride) idx[base_off + (stride - 1 - pos)] = j;
          }
        }
        n_in += __popc(m_in); n_out += __popc(m_out);
      }
    }
  }
  const int total = n_in + n_out;
  if (FILL && total <= stride && n_out > 0) {
    // move the back block [stride - n_out, stride) down to [n_in, n_in + n_out): ascending chunks, destination never ahead of
    // the unread source (n_in <= stride - n_out), loads of a chunk complete before its stores
    const int src0 = stride - n_out;
    if (src0 > n_in) {
      __syncwarp();
      for (int t0 = 0; t0 < n_out; t0 += 32) {
        const int t = t0 + lane;
        int val = 0;
        if (t < n_out) val = idx[base_off + src0 + t];
        __syncwarp();
        if (t < n_out) idx[base_off + n_in + t] = val;
        __syncwarp();
      }
    }
  }
  if (lane == 0) {
    cnt[i] = total;
    if (cnt_in) cnt_in[i] = n_in;
    if (FILL) off[i] = base_off;
    if (FILL && i == nrows - 1) off[nrows] = base_off + stride;
  }
}

// max and sum of the row lengths
__global__ void k_cnt_stats(const int* __restrict__ cnt, int n, long long* __restrict__ stats) {
  long long s = 0;
  int m = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { s += cnt[i]; m = max(m, cnt[i]); }
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); m = max(m, __shfl_xor_sync(0xffffffffu, m, o)); }
  if ((threadIdx.x & 31) == 0) {
    atomicMax((unsigned long long*)&stats[0], (unsigned long long)m);
    atomicAdd((unsigned long long*)&stats[1], (unsigned long long)s);
  }
}

__global__ void k_cnt_to_ll(const int* __restrict__ cnt, int n, long long* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = cnt[i];
  if (i == n) out[n] = 0;
}

}  // namespace

// Half-width (in r^2) of the band in which an fp32 distance cannot decide r^2 <= cut^2: each shadow coordinate carries
// a rounding error <= extent * 2^-24, a difference twice that, r^2 therefore ~ 2 r sqrt(3) * that, plus the fp32 arithmetic
// of the sum (relative 2^-22).  A factor 2 of slack is added on top.
float CellList::fp32_band(double cut) const {
  const double coord = (extent + 8.0) * 5.97e-8;
  return (float)(2.0 * (2.0 * cut * 1.7320508 * 2.0 * coord + cut * cut * 2.4e-7));
}

void CellList::bin(const double4* xq, int N, double bin_size, int reach_, cudaStream_t st) {
  reach = reach_;
  // bounding box (one small D2H per rebuild)
  bounds.resize(6);
  double init[6];
  {
    auto enc = [](double d) { long long b; memcpy(&b, &d, 8); b = b >= 0 ? b : b ^ 0x7fffffffffffffffLL; double o; memcpy(&o, &b, 8); return o; };
    for (int t = 0; t < 3; t++) { init[t] = enc(1e300); init[3 + t] = enc(-1e300); }
  }
  RXB_CUDA(cudaMemcpyAsync(bounds.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
  k_bounds<<<296, 256, 0, st>>>(xq, N, bounds.p);
  double got[6];
  RXB_CUDA(cudaMemcpyAsync(got, bounds.p, sizeof(got), cudaMemcpyDeviceToHost, st));
  RXB_SYNC(st);
  for (int t = 0; t < 6; t++) { long long b; memcpy(&b, &got[t], 8); b = b >= 0 ? b : b ^ 0x7fffffffffffffffLL; memcpy(&got[t], &b, 8); }
  Grid g;
  long long nbins = 1;
  for (int t = 0; t < 3; t++) {
    double ext = got[3 + t] - got[t];
    if (!(ext > 1e-9)) ext = 1e-9;
    int nb = (int)(ext / (t == 0 ? bin_size / kXFine : bin_size));
    if (nb < 1) nb = 1;
    g.nb[t] = nb; g.lo[t] = got[t]; g.size[t] = ext / nb; g.inv[t] = nb / ext;
    nbins *= nb;
  }
  memcpy(grid_blob, &g, sizeof(g));
  static_assert(sizeof(Grid) <= sizeof(grid_blob), "grid blob too small");
  num_bins = (int)nbins;
  key.resize(N); val.resize(N); key2.resize(N); sorted_idx.resize(N); spos.resize(N); sposf.resize(N);
  bin_count.resize(num_bins + 1); bin_start.resize(num_bins + 1);
  RXB_CUDA(cudaMemsetAsync(bin_count.p, 0, (num_bins + 1) * sizeof(int), st));
  k_bin_ids<<<(N + 255) / 256, 256, 0, st>>>(xq, N, g, key.p, val.p, bin_count.p);
  size_t need = 0, need2 = 0;
  int bits = 1;
  while ((1LL << bits) < nbins) bits++;
  cub::DeviceRadixSort::SortPairs(nullptr, need, key.p, key2.p, val.p, sorted_idx.p, N, 0, bits, st);
  cub::DeviceScan::ExclusiveSum(nullptr, need2, bin_count.p, bin_start.p, num_bins + 1, st);
  if (need2 > need) need = need2;
  temp.resize(need + 16);
  cub::DeviceRadixSort::SortPairs(temp.p, need, key.p, key2.p, val.p, sorted_idx.p, N, 0, bits, st);
  cub::DeviceScan::ExclusiveSum(temp.p, need, bin_count.p, bin_start.p, num_bins + 1, st);
  k_gather_sorted<<<(N + 255) / 256, 256, 0, st>>>(xq, sorted_idx.p, N, g, spos.p, sposf.p);
  extent = 0.0;
  for (int t = 0; t < 3; t++) { origin[t] = g.lo[t]; extent = std::max(extent, got[3 + t] - got[t]); }
  RXB_CUDA(cudaGetLastError());
}

void CellList::build(const double4* xq, int nrows, double cut, double cut_in, Csr& out, cudaStream_t st, const int* rowpos) {
  Grid g;
  memcpy(&g, grid_blob, sizeof(g));
  const float band = fp32_band(cut);
  // inner class: fp32 r^2 <= cut_in^2 + band is a superset of r <= cut_in (the band bounds the fp32 error of r^2)
  const bool part = cut_in > 0.0 && cut_in < cut;
  const float in2hi = part ? __builtin_nextafterf((float)(cut_in * cut_in), INFINITY) + band : 3.0e38f;
  out.cut_in = part ? cut_in : 0.0;
  out.cnt_in.resize(nrows + 1);
  const int warps_per_block = 8;
  const int blocks = (nrows + warps_per_block - 1) / warps_per_block;
  out.cnt.resize(nrows + 1);
  out.off.resize(nrows + 1);
  out.stats.resize(2);
  out.nrows = nrows;
  auto stats = [&](long long* host2) {
    RXB_CUDA(cudaMemsetAsync(out.stats.p, 0, 2 * sizeof(long long), st));
    if (nrows > 0) k_cnt_stats<<<std::min(1024, (nrows + 255) / 256), 256, 0, st>>>(out.cnt.p, nrows, out.stats.p);
    RXB_CUDA(cudaMemcpyAsync(host2, out.stats.p, 2 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    RXB_SYNC(st);
  };
  auto stride_for = [](long long longest) { return (int)(((longest + longest / 16 + 32) + 31) / 32 * 32); };
  long long got[2] = {0, 0};
  if (out.stride == 0 && nrows > 0) {   // first build: one counting pass sizes the stride
    if (rowpos)
      k_build<false, true><<<blocks, 256, 0, st>>>(xq, rowpos, spos.p, sposf.p, bin_start.p, g, nrows, cut, band, in2hi, reach,
                                                   out.cnt.p, nullptr, nullptr, nullptr, 0);
    else
      k_build<false, false><<<blocks, 256, 0, st>>>(xq, nullptr, spos.p, sposf.p, bin_start.p, g, nrows, cut, band, in2hi, reach,
                                                    out.cnt.p, nullptr, nullptr, nullptr, 0);
    stats(got);
    out.stride = stride_for(got[0]);
  }
  if (out.stride == 0) out.stride = 32;
  for (int attempt = 0; attempt < 4; attempt++) {
    out.slots = (long long)nrows * out.stride;
    out.idx.resize((size_t)std::max<long long>(out.slots, 1));
    if (nrows > 0 && rowpos)
      k_build<true, true><<<blocks, 256, 0, st>>>(xq, rowpos, spos.p, sposf.p, bin_start.p, g, nrows, cut, band, in2hi, reach,
                                                  out.cnt.p, out.cnt_in.p, out.off.p, out.idx.p, out.stride);
    else if (nrows > 0)
      k_build<true, false><<<blocks, 256, 0, st>>>(xq, nullptr, spos.p, sposf.p, bin_start.p, g, nrows, cut, band, in2hi, reach,
                                                   out.cnt.p, out.cnt_in.p, out.off.p, out.idx.p, out.stride);
    stats(got);
    if (got[0] <= out.stride) break;
    out.stride = stride_for(got[0]);     // a row outgrew the stride of the previous build: run the pass again
  }
  out.nnz = got[1];
  RXB_CUDA(cudaGetLastError());
}

}  // namespace rxb
