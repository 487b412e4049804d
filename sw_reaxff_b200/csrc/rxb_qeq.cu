// Charge equilibration (fix qeq/reax) on the GPU: both solves (H s = -chi, H t = -1) run as ONE dual-right-hand-side
// pipelined Jacobi-PCG; s and t are interleaved as double2 so every gathered x[j] serves both systems and H is read
// once per iteration instead of twice.
//
// Restated from /root/reference/fix_qeq_reax_sunway.cpp:
//   init_matvec :618-723 (Hdia_inv, b_s, b_t, cubic / quadratic extrapolation), sparse_matvec :1601-1622,
//   CG_v2 :983-1167 (Ghysels-Vanroose pipelined PCG, one reduction per iteration, imax = 200, test on sqrt(u.r)/|b|),
//   calculate_Q :1697-1755 (q = s - (sum s / sum t) t, 5-deep history).
// Each right-hand side keeps its own alpha/beta/eta and its own convergence flag ON THE DEVICE, so the fused solve
// performs exactly the iterations the reference's two sequential solves would (same matvec counts).  The per-iteration
// vector work of the reference (two sweeps + a serial MPE sweep overlapped with the SpMV) is one fused sweep here.
//
// Index spaces (rxb_dev.cuh): the solve runs in S space.  Row vectors (r, u, w, p, ...) are indexed by row = local atom in
// cell-sorted order; the two gathered vectors (the initial guess and the search direction d) are indexed by sorted
// position over all atoms, so the SpMV's x[col] gathers walk runs of consecutive 16-byte elements.
// H entries are 8 bytes (22-bit column + 42-bit fixed-point value in one word, rxb_nonbonded.cu) or 12 bytes (exact).
// Host involvement: none inside the solve.  The host launches as many iterations as the previous step needed (+ margin;
// converged iterations are gated off on the device and cost ~2 us each), and the convergence state is read with the
// end-of-step status; only an under-prediction (rare) continues the solve and replays the force phase.
// Roofline: the SpMV is HBM-bound: bytes per stored H entry + 16 B per gathered x (L1/L2-resident).
#include <algorithm>

#include "rxb_system.h"

namespace rxb {

namespace {

constexpr int kWarps = 8;
constexpr int kVecBlocks = 148 * 4;
constexpr int kVecThreads = 256;

__global__ void k_qeq_init(int n, const int* __restrict__ rowpos, const int* __restrict__ row_atom,
                           const int* __restrict__ type_s, const AtomPar* __restrict__ atom,
                           const double* __restrict__ s_hist, const double* __restrict__ t_hist, double2* __restrict__ x_row,
                           double2* __restrict__ xS, double2* __restrict__ b, double* __restrict__ Hdia_inv,
                           double* __restrict__ eta_row, QeqDev* __restrict__ Q, const int* __restrict__ ltype_s,
                           const double* __restrict__ chi_lt, const double* __restrict__ eta_lt) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int k = rowpos[j], i = row_atom[j];
    const int ti = type_s[k];
    double eta = 1.0, chi = 0.0;
    if (ti >= 0) {
      if (chi_lt) { const int lt = ltype_s[k]; eta = eta_lt[lt]; chi = chi_lt[lt]; }   // fix qeq/reax <param file>
      else { eta = atom[ti].eta; chi = atom[ti].chi; }
    }
    Hdia_inv[j] = 1. / eta;
    eta_row[j] = ti >= 0 ? eta : 0.0;
    b[j] = make_double2(-chi, -1.0);
    const double* sh = s_hist + 5 * (size_t)i;
    const double* th = t_hist + 5 * (size_t)i;
    const double2 x0 = make_double2(4 * (sh[0] + sh[2]) - (6 * sh[1] + sh[3]), th[2] + 3 * (th[0] - th[1]));
    x_row[j] = x0;
    xS[k] = x0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int k = 0; k < 3; k++) for (int c = 0; c < 4; c++) Q->dots[k][c] = 0.0;
    for (int c = 0; c < 6; c++) Q->pro[c] = 0.0;
    Q->sums[0] = Q->sums[1] = 0.0;
  }
}

// periodic-image ghosts of an S-space vector <- their owners (single-rank forward_comm_fix)
__global__ void k_forward2S(int nghost, const int* __restrict__ gs_pos, const int* __restrict__ gs_own, double2* __restrict__ vec) {
  pdl_wait(); pdl_release();
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < nghost; g += gridDim.x * blockDim.x) {
    const int o = gs_own[g];
    if (o >= 0) vec[gs_pos[g]] = vec[o];
  }
}

// y[row] = eta x[rowpos[row]] + sum_k H_k x[col_k] for the rows rowlist[r0 .. r1) (rowlist == null: rows r0 .. r1);
// gate != null: skip when neither system is active
template <bool PACKED>
__global__ void __launch_bounds__(kWarps * 32)
k_spmv2(int r0, int r1, const int* __restrict__ rowlist, int stride, const int* __restrict__ num,
        const unsigned long long* __restrict__ hpk, const int* __restrict__ col, const double* __restrict__ val, double inv_quant,
        const int* __restrict__ rowpos, const double* __restrict__ eta_row, const double2* __restrict__ x,
        double2* __restrict__ y, const QeqDev* __restrict__ Q, int parity, unsigned long long* __restrict__ active_launches) {
  if (Q != nullptr && !(Q->st[parity].active[0] | Q->st[parity].active[1])) return;
  if (active_launches && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(active_launches, 1ULL);   // launches that really multiply
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
  for (int t = r0 + wg; t < r1; t += nwg) {
    const int i = rowlist ? rowlist[t] : t;
    const long long beg = (long long)i * stride;
    const int m = num[i];
    double ax = 0, ay = 0;
    if (PACKED) {
      // H is streamed exactly once per SpMV (evict-first), 8 bytes per entry: the column is the top 22 bits, the value the
      // low 42 as an integer multiple of 1 / quant.  The integer becomes a double without a conversion instruction:
      // OR it into the mantissa of 2^52 and subtract 2^52 (exact); the row sum is scaled by 1 / quant once at the end.
      const unsigned long long* __restrict__ hp = hpk + beg;
#pragma unroll 4
      for (int k = lane; k < m; k += 32) {
        const unsigned long long w = __ldcs(hp + k);
        const double h = __longlong_as_double((long long)((w & kHValMask) | 0x4330000000000000ULL)) - 4503599627370496.0;
        const double2 xj = __ldg(x + (int)(w >> kHColShift));
        ax += h * xj.x; ay += h * xj.y;
      }
      ax *= inv_quant; ay *= inv_quant;
    } else {
#pragma unroll 4
      for (int k = lane; k < m; k += 32) {
        const double h = __ldcs(val + beg + k);
        const double2 xj = __ldg(x + __ldcs(col + beg + k));
        ax += h * xj.x; ay += h * xj.y;
      }
    }
    ax = warp_sum(ax); ay = warp_sum(ay);
    if (lane == 0) {
      const double eta = eta_row[i];
      const double2 xi = x[rowpos[i]];
      y[i] = make_double2(eta * xi.x + ax, eta * xi.y + ay);
    }
  }
}

// Deep-pipeline form of the packed SpMV (experiment, RXB_SPMV_DEEP=U): every lane issues U word loads, then U gathers, then
// the FMAs, so a warp keeps 32 U entries in flight and the kernel can saturate HBM from fewer resident warps - which would
// leave registers and warp slots to the bonded chain running beside it.
template <int U>
__global__ void __launch_bounds__(kWarps * 32)
k_spmv2_deep(int r0, int r1, int stride, const int* __restrict__ num, const unsigned long long* __restrict__ hpk,
             double inv_quant, const int* __restrict__ rowpos, const double* __restrict__ eta_row, const double2* __restrict__ x,
             double2* __restrict__ y, const QeqDev* __restrict__ Q, int parity, unsigned long long* __restrict__ active_launches) {
  pdl_wait(); pdl_release();
  if (Q != nullptr && !(Q->st[parity].active[0] | Q->st[parity].active[1])) return;
  if (active_launches && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(active_launches, 1ULL);   // launches that really multiply
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
  for (int i = r0 + wg; i < r1; i += nwg) {
    const unsigned long long* __restrict__ hp = hpk + (long long)i * stride;
    const int m = num[i];
    double ax = 0, ay = 0;
    for (int k0 = 0; k0 < m; k0 += 32 * U) {
      unsigned long long w[U];
      double2 xj[U];
#pragma unroll
      for (int u = 0; u < U; u++) { const int k = k0 + 32 * u + lane; w[u] = k < m ? __ldcs(hp + k) : 0ULL; }
#pragma unroll
      for (int u = 0; u < U; u++) xj[u] = __ldg(x + (int)(w[u] >> kHColShift));     // padding words gather x[0] times 0
#pragma unroll
      for (int u = 0; u < U; u++) {
        const double h = __longlong_as_double((long long)((w[u] & kHValMask) | 0x4330000000000000ULL)) - 4503599627370496.0;
        ax += h * xj[u].x; ay += h * xj[u].y;
      }
    }
    ax = warp_sum(ax * inv_quant); ay = warp_sum(ay * inv_quant);
    if (lane == 0) {
      const double eta = eta_row[i];
      const double2 xi = x[rowpos[i]];
      y[i] = make_double2(eta * xi.x + ax, eta * xi.y + ay);
    }
  }
}

// Bulk-async form of the packed SpMV (experiment, RXB_SPMV_BULK=1): the H words of a row are contiguous, so one elected lane
// per warp fetches them with cp.async.bulk (the TMA engine, SASS UBLKCP) into a warp-private double buffer in shared memory
// and completion is signalled on an mbarrier; the copy of stage s+1 - the next 256 entries of the row, or the head of the
// warp's NEXT row - is in flight while stage s is multiplied.  The word stream then costs no registers, no global LSU
// wavefronts and no scoreboard slots; the 16-byte gathers of the CG vector stay as they are.  Persistent grid, one warp
// walks rows wg, wg + nwarps, ...  Measured against k_spmv2_deep<8>: profiles/r02_spmv_ab.txt.
constexpr int kBulkChunk = 256;   // entries per stage (2 KB): 8 per lane, the depth of the deep kernel
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_issue(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  if (bytes == 0) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); return; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the lanes' earlier generic reads of this buffer are done
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  for (int spin = 0; !ok; spin++) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (spin > (1 << 24)) __trap();    // a copy that never completes must not hang the GPU
  }
}
__global__ void __launch_bounds__(kWarps * 32)
k_spmv2_bulk(int r0, int r1, int stride, const int* __restrict__ num, const unsigned long long* __restrict__ hpk,
             double inv_quant, const int* __restrict__ rowpos, const double* __restrict__ eta_row, const double2* __restrict__ x,
             double2* __restrict__ y, const QeqDev* __restrict__ Q, int parity, unsigned long long* __restrict__ active_launches) {
  pdl_wait(); pdl_release();
  if (Q != nullptr && !(Q->st[parity].active[0] | Q->st[parity].active[1])) return;
  if (active_launches && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(active_launches, 1ULL);
  __shared__ __align__(16) unsigned long long s_buf[kWarps][2][kBulkChunk];
  __shared__ __align__(8) unsigned long long s_bar[kWarps][2];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
  const unsigned bar[2] = {smem_u32(&s_bar[wib][0]), smem_u32(&s_bar[wib][1])};
  const unsigned dst[2] = {smem_u32(&s_buf[wib][0][0]), smem_u32(&s_buf[wib][1][0])};
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar[0]) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar[1]) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  int i = r0 + wg;
  if (i >= r1) return;
  int m = num[i], c = 0;
  auto stage_bytes = [](int mm, int cc) { return (unsigned)(((min(kBulkChunk, mm - cc * kBulkChunk) * 8) + 15) & ~15); };
  if (lane == 0) bulk_issue(dst[0], hpk + (long long)i * stride, m > 0 ? stage_bytes(m, 0) : 0u, bar[0]);
  double ax = 0, ay = 0;
  for (unsigned s = 0;; s++) {
    // the stage after this one: the next chunk of the row, or the head of the warp's next row
    int ni = i, nc = c + 1, nm = m;
    if (nc * kBulkChunk >= m) { ni = i + nwg; nc = 0; nm = ni < r1 ? num[ni] : 0; }
    const bool has_next = ni < r1;
    if (has_next && lane == 0)
      bulk_issue(dst[(s + 1) & 1], hpk + (long long)ni * stride + (long long)nc * kBulkChunk, nm > 0 ? stage_bytes(nm, nc) : 0u,
                 bar[(s + 1) & 1]);
    bulk_wait(bar[s & 1], (s >> 1) & 1);
    const unsigned long long* wbuf = &s_buf[wib][s & 1][0];
    const int cnt = min(kBulkChunk, m - c * kBulkChunk);
    unsigned long long w[8];
    double2 xj[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { const int k = 32 * u + lane; w[u] = k < cnt ? wbuf[k] : 0ULL; }
#pragma unroll
    for (int u = 0; u < 8; u++) xj[u] = __ldg(x + (int)(w[u] >> kHColShift));
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const double h = __longlong_as_double((long long)((w[u] & kHValMask) | 0x4330000000000000ULL)) - 4503599627370496.0;
      ax += h * xj[u].x; ay += h * xj[u].y;
    }
    if ((c + 1) * kBulkChunk >= m) {   // last stage of the row
      ax = warp_sum(ax * inv_quant); ay = warp_sum(ay * inv_quant);
      if (lane == 0) {
        const double eta = eta_row[i];
        const double2 xi = x[rowpos[i]];
        y[i] = make_double2(eta * xi.x + ax, eta * xi.y + ay);
      }
      ax = 0; ay = 0;
    }
    __syncwarp();                      // every lane has read this buffer before stage s + 2 refills it
    if (!has_next) break;
    i = ni; c = nc; m = nm;
  }
}

// prologue steps (fix_qeq_reax_sunway.cpp:1024-1105)
__global__ void k_pro1(int n, const int* __restrict__ rowpos, const double2* __restrict__ b, const double2* __restrict__ q,
                       const double* __restrict__ Hd, double2* __restrict__ r, double2* __restrict__ u, double2* __restrict__ dS) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double2 bj = b[j], qj = q[j];
    const double2 rj = make_double2(bj.x - 1 * qj.x, bj.y - 1 * qj.y);
    const double2 uj = make_double2(rj.x * Hd[j], rj.y * Hd[j]);
    r[j] = rj; u[j] = uj; dS[rowpos[j]] = uj;
  }
}
__global__ void k_pro2(int n, const int* __restrict__ rowpos, const double2* __restrict__ q, const double* __restrict__ Hd,
                       double2* __restrict__ w, double2* __restrict__ m, double2* __restrict__ dS) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double2 qj = q[j];
    const double2 mj = make_double2(qj.x * Hd[j], qj.y * Hd[j]);
    w[j] = qj; m[j] = mj; dS[rowpos[j]] = mj;
  }
}

template <int K>
__device__ __forceinline__ void block_reduce_add(double (&v)[K], double* dst) {
  __shared__ double sh[K][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; k++) { const double s = warp_sum(v[k]); if (lane == 0) sh[k][w] = s; }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      double s = lane < nw ? sh[k][lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) atomicAdd(&dst[k], s);
    }
  }
}

__global__ void __launch_bounds__(kVecThreads)
k_pro3(int n, const double2* __restrict__ b, const double2* __restrict__ r, const double2* __restrict__ u,
       const double2* __restrict__ w, const double2* __restrict__ m, const double2* __restrict__ q, double2* __restrict__ p,
       double2* __restrict__ ss, double2* __restrict__ v, double2* __restrict__ z, QeqDev* __restrict__ Q) {
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double2 uj = u[j], wj = w[j], bj = b[j], rj = r[j];
    p[j] = uj; ss[j] = wj; v[j] = m[j]; z[j] = q[j];
    acc[0] += bj.x * bj.x; acc[1] += uj.x * rj.x; acc[2] += uj.x * wj.x;
    acc[3] += bj.y * bj.y; acc[4] += uj.y * rj.y; acc[5] += uj.y * wj.y;
  }
  block_reduce_add<6>(acc, Q->pro);
}

__global__ void k_scal_init(QeqDev* Q, double tol, int imax) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  QeqState& S = Q->st[1];  // iteration 1 reads parity 1
  for (int c = 0; c < 2; c++) {
    const double bb = Q->pro[3 * c], ur = Q->pro[3 * c + 1], uw = Q->pro[3 * c + 2];
    S.b_norm[c] = sqrt(bb);
    S.sig_old[c] = ur;
    S.heta[c] = uw;
    S.alpha[c] = ur / uw;
    S.dot0[c] = ur;
    S.iters[c] = 1;
    S.active[c] = (1 < imax) && (sqrt(ur) / S.b_norm[c] > tol);
  }
}

// Fused sweep for loop index `it` (>= 1):  [B_{it-1}: x,p,ss,v,z update with the SpMV result]  then
// [A_it: r,u,w update, the two dot products, d = M^-1 w].  `first` skips the B part (it == 1).
// d lives in S space (dS[rowpos[j]]); everything else is a row vector.
__global__ void __launch_bounds__(kVecThreads)
k_cg_sweep(int n, int it, int first, double tol, int imax, const int* __restrict__ rowpos, const double* __restrict__ Hd,
           const double2* __restrict__ q, double2* __restrict__ x, double2* __restrict__ r, double2* __restrict__ u,
           double2* __restrict__ w, double2* __restrict__ p, double2* __restrict__ ss, double2* __restrict__ v,
           double2* __restrict__ z, double2* __restrict__ dS, QeqDev* __restrict__ Q, const int* __restrict__ img_off,
           const int* __restrict__ img_pos) {
  pdl_wait(); pdl_release();
  const int par = it & 1;
  const QeqState S = Q->st[par];
  QeqState T = S;  // state after the B part; identical in every thread
  double beta[2] = {0, 0}, alpha_old[2] = {S.alpha[0], S.alpha[1]};
  int doB[2] = {0, 0};
  if (!first) {
    const double* D = Q->dots[it % 3];  // accumulated by the previous sweep's A part
    for (int c = 0; c < 2; c++) {
      if (S.active[c]) {
        doB[c] = 1;
        const double d0 = D[c], d1 = D[2 + c];
        beta[c] = d0 / S.sig_old[c];
        T.heta[c] = d1 - beta[c] * beta[c] * S.heta[c];
        T.alpha[c] = d0 / T.heta[c];
        T.sig_old[c] = d0;
        T.dot0[c] = d0;
        T.iters[c] = S.iters[c] + 1;
        T.active[c] = (T.iters[c] < imax) && (sqrt(d0) / S.b_norm[c] > tol);
      }
    }
  }
  const int actA0 = T.active[0], actA1 = T.active[1];
  double acc[4] = {0, 0, 0, 0};
  if (doB[0] | doB[1] | actA0 | actA1)
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
      const int kj = rowpos[j];
      double2 xj = x[j], rj = r[j], uj = u[j], wj = w[j], pj = p[j], sj = ss[j], vj = v[j], zj = z[j], dj = dS[kj];
      if (doB[0] | doB[1]) {
        const double2 qj = q[j];
        if (doB[0]) {
          xj.x += alpha_old[0] * pj.x;
          pj.x = uj.x + pj.x * beta[0]; sj.x = wj.x + sj.x * beta[0]; vj.x = dj.x + vj.x * beta[0];
          zj.x = qj.x + zj.x * beta[0];
        }
        if (doB[1]) {
          xj.y += alpha_old[1] * pj.y;
          pj.y = uj.y + pj.y * beta[1]; sj.y = wj.y + sj.y * beta[1]; vj.y = dj.y + vj.y * beta[1];
          zj.y = qj.y + zj.y * beta[1];
        }
        x[j] = xj; p[j] = pj; ss[j] = sj; v[j] = vj; z[j] = zj;
      }
      if (actA0 | actA1) {
        const double hd = Hd[j];
        if (actA0) {
          rj.x -= T.alpha[0] * sj.x; uj.x -= T.alpha[0] * vj.x; wj.x -= T.alpha[0] * zj.x;
          acc[0] += uj.x * rj.x; acc[2] += uj.x * wj.x;
          dj.x = wj.x * hd;
        }
        if (actA1) {
          rj.y -= T.alpha[1] * sj.y; uj.y -= T.alpha[1] * vj.y; wj.y -= T.alpha[1] * zj.y;
          acc[1] += uj.y * rj.y; acc[3] += uj.y * wj.y;
          dj.y = wj.y * hd;
        }
        r[j] = rj; u[j] = uj; w[j] = wj; dS[kj] = dj;
        if (img_off)   // this row's periodic images (single-rank forward_comm_fix, fused)
          for (int k = img_off[j], k1 = img_off[j + 1]; k < k1; k++) dS[img_pos[k]] = dj;
      }
    }
  if (actA0 | actA1) block_reduce_add<4>(acc, Q->dots[(it + 1) % 3]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    Q->st[par ^ 1] = T;
    double* Z = Q->dots[(it + 2) % 3];
    Z[0] = Z[1] = Z[2] = Z[3] = 0.0;
  }
}

__global__ void __launch_bounds__(kVecThreads)
k_q_sums(int n, const double2* __restrict__ x, QeqDev* __restrict__ Q) {
  double acc[2] = {0, 0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double2 xj = x[j];
    acc[0] += xj.x; acc[1] += xj.y;
  }
  block_reduce_add<2>(acc, Q->sums);
}

// phase 0: q of the local atoms + history (shift = 0: the solve was continued after the history had already been shifted
// for this step, only slot 0 is overwritten); phase 1: ghost charges <- owners (single rank)
__global__ void k_q_final(int n, int N, const int* __restrict__ row_atom, const int* __restrict__ owner,
                          const double2* __restrict__ x, const QeqDev* __restrict__ Q, double* __restrict__ s_hist,
                          double* __restrict__ t_hist, double4* __restrict__ xq, int phase, int shift) {
  if (phase == 0) {
    const double uu = Q->sums[0] / Q->sums[1];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
      const int i = row_atom[j];
      const double2 xi = x[j];
      xq[i].w = xi.x - uu * xi.y;
      double* sh = s_hist + 5 * (size_t)i;
      double* th = t_hist + 5 * (size_t)i;
      if (shift) for (int k = 4; k > 0; --k) { sh[k] = sh[k - 1]; th[k] = th[k - 1]; }
      sh[0] = xi.x; th[0] = xi.y;
    }
  } else {
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < N - n; g += gridDim.x * blockDim.x) {
      const int o = owner[g];
      if (o >= 0) xq[n + g].w = xq[o].w;
    }
  }
}

__global__ void k_q_to_S(int N, const int* __restrict__ s2a, const double4* __restrict__ xq, double4* __restrict__ xqs) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) xqs[k].w = xq[s2a[k]].w;
}

__global__ void k_zero2(double* p) { if (threadIdx.x == 0) { p[0] = 0.0; p[1] = 0.0; } }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
void System::qeq_reset_history() {
  q_s_hist.resize((size_t)5 * n); q_t_hist.resize((size_t)5 * n);
  RXB_CUDA(cudaMemsetAsync(q_s_hist.p, 0, (size_t)5 * n * sizeof(double), st_));
  RXB_CUDA(cudaMemsetAsync(q_t_hist.p, 0, (size_t)5 * n * sizeof(double), st_));
}
void System::qeq_set_history(const double* s_hist, const double* t_hist) {
  q_s_hist.resize((size_t)5 * n); q_t_hist.resize((size_t)5 * n);
  RXB_CUDA(cudaMemcpyAsync(q_s_hist.p, s_hist, (size_t)5 * n * sizeof(double), cudaMemcpyHostToDevice, st_));
  RXB_CUDA(cudaMemcpyAsync(q_t_hist.p, t_hist, (size_t)5 * n * sizeof(double), cudaMemcpyHostToDevice, st_));
  RXB_SYNC(st_);
}
void System::qeq_get_history(double* s_hist, double* t_hist) {
  RXB_CUDA(cudaMemcpyAsync(s_hist, q_s_hist.p, (size_t)5 * n * sizeof(double), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaMemcpyAsync(t_hist, q_t_hist.p, (size_t)5 * n * sizeof(double), cudaMemcpyDeviceToHost, st_));
  RXB_SYNC(st_);
}

// one CG iteration of loop index `it`: fused sweep, halo of d (+ the dot products in multi-GPU runs), gated SpMV
void System::qeq_iteration(int it) {
  QeqDev* Q = reinterpret_cast<QeqDev*>(q_scal.p);
  const bool fused = !dist_ && img_valid_;   // the sweep stores each row's value into its periodic images itself
  // two CTAs per SM (A/B on one box, RXB_SWEEP_BLOCKS: 148 -> 10.67, 296 -> 10.53, 444 -> 10.63, 592 -> 10.60, 1184 -> 10.72 ms/step)
  static const int sweep_blocks = getenv("RXB_SWEEP_BLOCKS") ? atoi(getenv("RXB_SWEEP_BLOCKS")) : 148 * 2;
  launch_pdl(k_cg_sweep, sweep_blocks, kVecThreads, 0, st_, n, it, (int)(it == 1), qeq_tol, qeq_imax, rowpos.p, q_Hdia_inv.p, q_q.p,
             q_x.p, q_r.p, q_u.p, q_w.p, q_p.p, q_ss.p, q_v.p, q_z.p, q_d.p, Q, fused ? img_off.p : nullptr,
             fused ? img_pos.p : nullptr);
  kernel_launches++;
  const int par_next = (it & 1) ^ 1;  // state written by this sweep (from the dot products of the sweep before it)
  // MPI_Allreduce(dot_local, 2) of each solve (:1132) and the boundary values of d travel in one exchange.  Multi-GPU with
  // peer windows: the values are pushed, the INTERIOR rows (no column owned by another rank) are multiplied while they
  // fly, then the ghosts are pulled and the boundary rows follow - the overlap the reference gets from
  // sparse_matvec_C_spawn ... sparse_matvec_C_join around its MPI calls (fix_qeq_reax_sunway.cpp:1130-1143).
  // Measured at N = 2 (A/B on one box, profiles/r02_split_ab.txt): the split costs more than the wait it hides - two launches
  // per SpMV add two ramp-up/drain phases (+10 % SpMV time), the row list another 2 % - so it is OFF unless RXB_SPLIT=1.
  if (qeq_split_rows() && dist_ && dist_peer_active() && n_interior_ > 0) {
    double* dots = Q->dots[(it + 1) % 3];
    dist_push2(q_d.p, dots, 4);
    qeq_spmv(q_d.p, q_q.p, true, par_next, 0, n_interior_);
    dist_pull2(q_d.p, dots, 4);
    qeq_spmv(q_d.p, q_q.p, true, par_next, n_interior_, n);
    return;
  }
  if (dist_) dist_forward2_dots(q_d.p, Q->dots[(it + 1) % 3]);
  else if (!fused) qeq_forward_S(q_d.p);
  qeq_spmv(q_d.p, q_q.p, true, par_next);
}

void System::qeq_forward_S(double2* vecS) {
  if (dist_) { dist_forward2(vecS); return; }
  const int nghost = N - n;
  if (nghost > 0) {
    launch_pdl(k_forward2S, std::min(148 * 8, (nghost + 255) / 256), 256, 0, st_, nghost, gs_pos.p, gs_own.p, vecS);
    kernel_launches++;
  }
}

void System::qeq_spmv(const double2* xS, double2* y_row, bool gated, int parity, int r0, int r1) {
  const QeqDev* Q = reinterpret_cast<const QeqDev*>(q_scal.p);
  if (r1 < 0) r1 = n;
  if (r1 <= r0) return;
  const int* rowlist = (qeq_split_rows() && dist_ && n_interior_ > 0) ? q_rowlist.p : nullptr;   // interior rows first, then boundary rows
  const int ts = tick(r0 > 0 ? StepTimers::SPMV_B : StepTimers::SPMV);
  // one row per warp, blocks retire continuously: the bond-chain stream (low priority) picks up the slots they free
  const int grid = std::max(1, (r1 - r0 + kWarps - 1) / kWarps);
  // RXB_SPMV_SMEM=<bytes> (development knob): an unused dynamic shared-memory request per CTA caps the CTAs resident per SM,
  // i.e. how many warp slots the SpMV leaves to the bonded chain running beside it on the second stream
  static const int smem = [] {
    const char* e = getenv("RXB_SPMV_SMEM");
    const int b = e ? atoi(e) : 0;
    if (b > 48 * 1024) {
      cudaFuncSetAttribute(k_spmv2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
      cudaFuncSetAttribute(k_spmv2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    }
    return b;
  }();
  // default: the deep-pipeline form with 8 words + 8 gathers in flight per lane (A/B on one box, profiles/r02_spmv_ab.txt:
  // SpMV 5.72 -> 5.31 ms/step, step 11.64 -> 11.22; 16 per lane is no better; any occupancy cap is far worse)
  static const int deep = getenv("RXB_SPMV_DEEP") ? atoi(getenv("RXB_SPMV_DEEP")) : 8;
  static const int deep_grid = getenv("RXB_SPMV_GRID") ? atoi(getenv("RXB_SPMV_GRID")) : 0;   // persistent grid (CTAs), 0 = one row per warp
  static const int bulk = getenv("RXB_SPMV_BULK") ? atoi(getenv("RXB_SPMV_BULK")) : 0;     // CTAs per SM of the bulk-async form, 0 = off
  if (h_packed_ && bulk > 0 && !rowlist) {
    launch_pdl(k_spmv2_bulk, std::min(148 * bulk, grid), kWarps * 32, 0, st_, r0, r1, vl.stride, far_num.p, hpk.p, 1.0 / h_quant_,
               rowpos.p, q_eta.p, xS, y_row, gated ? Q : nullptr, parity, r0 == 0 ? spmv_active_d.p : nullptr);
  } else if (h_packed_ && deep && !rowlist) {
    // four waves of resident CTAs, each warp walking rows wg, wg + nwarps, ... (A/B on one box, TATB 8x8x8, step time: one row
    // per warp = 24576 CTAs 10.53 ms; 1 wave 10.87; 2 waves 10.53; 4 waves 10.13; 6 waves 10.22; 8 waves 10.22; grids that are
    // not a whole number of waves are worse: 4.67 waves 10.20, 5.33 waves 10.35) - SpMV 149 -> 136 us per launch
    static int occ_deep = 0;
    const int g = deep_grid > 0 ? deep_grid : (deep_grid < 0 ? grid : std::min(grid, wave_grid(k_spmv2_deep<8>, kWarps * 32, 4, occ_deep)));
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
      cudaFuncSetAttribute(k_spmv2_deep<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaFuncSetAttribute(k_spmv2_deep<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      attr_set = true;
    }
    if (deep >= 16)
      k_spmv2_deep<16><<<g, kWarps * 32, smem, st_>>>(r0, r1, vl.stride, far_num.p, hpk.p, 1.0 / h_quant_, rowpos.p, q_eta.p, xS, y_row,
                                                      gated ? Q : nullptr, parity, r0 == 0 ? spmv_active_d.p : nullptr);
    else
      launch_pdl(k_spmv2_deep<8>, g, kWarps * 32, smem, st_, r0, r1, vl.stride, far_num.p, hpk.p, 1.0 / h_quant_, rowpos.p, q_eta.p,
                 xS, y_row, gated ? Q : nullptr, parity, r0 == 0 ? spmv_active_d.p : nullptr);
  } else if (h_packed_)
    k_spmv2<true><<<grid, kWarps * 32, smem, st_>>>(r0, r1, rowlist, vl.stride, far_num.p, hpk.p, nullptr, nullptr, 1.0 / h_quant_,
                                                    rowpos.p, q_eta.p, xS, y_row, gated ? Q : nullptr, parity, r0 == 0 ? spmv_active_d.p : nullptr);
  else
    k_spmv2<false><<<grid, kWarps * 32, smem, st_>>>(r0, r1, rowlist, vl.stride, far_num.p, nullptr, far_idx.p, H_val.p, 1.0,
                                                     rowpos.p, q_eta.p, xS, y_row, gated ? Q : nullptr, parity, r0 == 0 ? spmv_active_d.p : nullptr);
  tock(ts);
  kernel_launches++;
}

// final charges from the current x (calculate_Q); shift_hist = false when this step's history slot was already opened
void System::qeq_finish(bool shift_hist) {
  QeqDev* Q = reinterpret_cast<QeqDev*>(q_scal.p);
  const int nghost = N - n;
  k_zero2<<<1, 32, 0, st_>>>(Q->sums);
  k_q_sums<<<kVecBlocks, kVecThreads, 0, st_>>>(n, q_x.p, Q);
  if (dist_) dist_sum_small(Q->sums, 2);
  k_q_final<<<kVecBlocks, kVecThreads, 0, st_>>>(n, N, row_atom.p, ghost_owner.p, q_x.p, Q, q_s_hist.p, q_t_hist.p, xq.p, 0,
                                                shift_hist ? 1 : 0);
  if (dist_) dist_forward_xq();
  else if (nghost > 0)
    k_q_final<<<std::min(148 * 8, (nghost + 255) / 256), 256, 0, st_>>>(n, N, row_atom.p, ghost_owner.p, q_x.p, Q, q_s_hist.p,
                                                                     q_t_hist.p, xq.p, 1, 0);
  k_q_to_S<<<kVecBlocks, kVecThreads, 0, st_>>>(N, s2a.p, xq.p, xqs.p);
  kernel_launches += 5;
}

// Reads the convergence state written by the last launched sweep.  Returns true when both solves have stopped.
bool System::qeq_poll() {
  QeqDev* Q = reinterpret_cast<QeqDev*>(q_scal.p);
  const int par_next = (qeq_it_ & 1) ^ 1;
  int host[4];
  RXB_CUDA(cudaMemcpyAsync(host, Q->st[par_next].active, 4 * sizeof(int), cudaMemcpyDeviceToHost, st_));   // active[2], iters[2]
  RXB_SYNC(st_);
  matvecs_s = host[2]; matvecs_t = host[3];
  return !(host[0] | host[1]);
}

void System::qeq_pre_force(bool wait_for_convergence) {
  if (n == 0) return;
  if (q_s_hist.n != (size_t)5 * n && (!dist_ || dist_external())) qeq_reset_history();
  last_swb_ = qeq_swb;
  choose_h_format();                    // (taper start / exact request may have changed since the build)
  DevView v = view();
  update_shadow(st_);
  // taper and shielding of the fix (init_taper :458-484, init_shielding :440-454), host side, tiny
  double Tap[8];
  {
    const double swa = qeq_swa, swb = qeq_swb, d7 = pow(swb - swa, 7);
    const double a2 = swa * swa, a3 = a2 * swa, b2 = swb * swb, b3 = b2 * swb;
    Tap[7] = 20.0 / d7;
    Tap[6] = -70.0 * (swa + swb) / d7;
    Tap[5] = 84.0 * (a2 + 3.0 * swa * swb + b2) / d7;
    Tap[4] = -35.0 * (a3 + 9.0 * a2 * swb + 9.0 * swa * b2 + b3) / d7;
    Tap[3] = 140.0 * (a3 * swb + 3.0 * a2 * b2 + swa * b3) / d7;
    Tap[2] = -210.0 * (a3 * b2 + a2 * b3) / d7;
    Tap[1] = 140.0 * a3 * b3 / d7;
    Tap[0] = (-35.0 * a3 * b2 * b2 + 21.0 * a2 * b3 * b2 + 7.0 * swa * b3 * b3 + b3 * b3 * swb) / d7;
  }
  const int t_QEQ_H = tick(StepTimers::QEQ_H);
  launch_far_and_H(*this, v, dp_, Tap, shld_d.p, qeq_swb, st_);
  memcpy(last_tap_, Tap, sizeof(last_tap_)); last_swb_ = qeq_swb;
  tock(t_QEQ_H);
  after_far_hook();

  const int t_QEQ_CG = tick(StepTimers::QEQ_CG);
  const size_t nn = n, NN = std::max((size_t)N, slab());
  q_xS.resize(NN); q_d.resize(NN);
  q_x.resize(nn); q_r.resize(nn); q_u.resize(nn); q_w.resize(nn); q_p.resize(nn); q_ss.resize(nn); q_v.resize(nn); q_z.resize(nn);
  q_q.resize(nn); q_b.resize(nn); q_m.resize(nn); q_Hdia_inv.resize(nn); q_eta.resize(nn);
  q_scal.resize(sizeof(QeqDev) / sizeof(double) + 8);
  QeqDev* Q = reinterpret_cast<QeqDev*>(q_scal.p);
  k_qeq_init<<<kVecBlocks, kVecThreads, 0, st_>>>(n, rowpos.p, row_atom.p, type_s.p, dp_.atom, q_s_hist.p, q_t_hist.p, q_x.p,
                                                 q_xS.p, q_b.p, q_Hdia_inv.p, q_eta.p, Q, v.ltype_s, v.chi_lt, v.eta_lt);
  qeq_forward_S(q_xS.p);
  qeq_spmv(q_xS.p, q_q.p, false, 0);
  k_pro1<<<kVecBlocks, kVecThreads, 0, st_>>>(n, rowpos.p, q_b.p, q_q.p, q_Hdia_inv.p, q_r.p, q_u.p, q_d.p);
  qeq_forward_S(q_d.p);
  qeq_spmv(q_d.p, q_q.p, false, 0);
  k_pro2<<<kVecBlocks, kVecThreads, 0, st_>>>(n, rowpos.p, q_q.p, q_Hdia_inv.p, q_w.p, q_m.p, q_d.p);
  qeq_forward_S(q_d.p);
  qeq_spmv(q_d.p, q_q.p, false, 0);
  k_pro3<<<kVecBlocks, kVecThreads, 0, st_>>>(n, q_b.p, q_r.p, q_u.p, q_w.p, q_m.p, q_q.p, q_p.p, q_ss.p, q_v.p, q_z.p, Q);
  if (dist_) dist_sum_small(Q->pro, 6);
  k_scal_init<<<1, 32, 0, st_>>>(Q, qeq_tol, qeq_imax);
  kernel_launches += 5;

  // main loop: sweep(it) ; halo ; SpMV.  The sweep after the last active iteration applies the final x update; sweeps
  // and SpMVs launched beyond convergence are gated off on the device.
  //  * resident run: as many iterations as the previous solve needed + a margin are enqueued WITHOUT a host round trip;
  //    the convergence state is read with the end-of-step status (System::qeq_settle), which continues the solve and
  //    replays the force phase in the rare case the prediction fell short;
  //  * plugin call (the caller wants matvecs back) / first solve: the host polls after the predicted count, then every
  //    few iterations.
  const int cap = qeq_imax + 1;
  // prediction = the largest count of the last reneighbouring cycle + a margin (counts wander by a few iterations from
  // step to step; a gated iteration costs ~7 us, an under-prediction a replay of the whole force phase)
  const int target = std::min(cap, qeq_predict_ > 0 ? qeq_predict_ + 3 + qeq_predict_ / 8 : 8);
  qeq_it_ = 0;
  for (int it = 1; it <= target; it++) { qeq_iteration(it); qeq_it_ = it; }
  qeq_unsettled_ = false;
  if (wait_for_convergence || qeq_predict_ <= 0) {
    while (!qeq_poll() && qeq_it_ < cap) {
      const int more = std::min(cap - qeq_it_, qeq_check_every);
      for (int k = 0; k < more; k++) { qeq_iteration(qeq_it_ + 1); qeq_it_++; }
    }
    qeq_record_iterations();
  } else {
    qeq_unsettled_ = true;          // settled by qeq_settle() at the end-of-step synchronisation
  }
  qeq_finish(true);
  qeq_ran_this_step_ = true;
  tock(t_QEQ_CG);
}

void System::qeq_record_iterations() {
  const int it = std::max(matvecs_s, matvecs_t);
  qeq_iters_total += it;
  qeq_recent_[qeq_recent_at_++ % 5] = it;
  qeq_predict_ = 0;
  for (int k = 0; k < 5; k++) qeq_predict_ = std::max(qeq_predict_, qeq_recent_[k]);
}

// End-of-step: has the solve that was enqueued without polling converged?  (Called right after the end-of-step
// synchronisation, so the poll below costs one small copy.)  Returns true when it had to be continued - the charges
// changed and the caller must replay the force phase.
bool System::qeq_settle() {
  if (!qeq_unsettled_) return false;
  qeq_unsettled_ = false;
  const int cap = qeq_imax + 1;
  bool continued = false;
  while (!qeq_poll() && qeq_it_ < cap) {
    const int more = std::min(cap - qeq_it_, qeq_check_every);
    for (int k = 0; k < more; k++) { qeq_iteration(qeq_it_ + 1); qeq_it_++; }
    continued = true;
  }
  qeq_record_iterations();
  if (continued) { qeq_finish(false); qeq_replays++; }
  return continued;
}

}  // namespace rxb
