// Charge equilibration (fix qeq/reax) on the GPU: both solves (H s = -chi, H t = -1) run as ONE dual-right-hand-side
// pipelined Jacobi-PCG; s and t are interleaved as double2 so every gathered x[j] serves both systems and H is read
// once per iteration instead of twice.
//
// Restated from /root/reference/fix_qeq_reax_sunway.cpp:
//   init_matvec :618-723 (Hdia_inv, b_s, b_t, cubic / quadratic extrapolation), sparse_matvec :1601-1622,
//   CG_v2 :983-1167 (Ghysels-Vanroose pipelined PCG, one reduction per iteration, imax = 200, test on sqrt(u.r)/|b|),
//   calculate_Q :1697-1755 (q = s - (sum s / sum t) t, 5-deep history).
// Each right-hand side keeps its own alpha/beta/eta and its own convergence flag ON THE DEVICE, so the fused solve
// performs exactly the iterations the reference's two sequential solves would (same matvec counts), and the host
// only polls a flag every few iterations.  The per-iteration vector work of the reference (two sweeps + a serial
// MPE sweep overlapped with the SpMV) is one fused sweep here.
// Roofline: SpMV is HBM-bound: 12 B per stored H entry + 16 B per gathered x (L2-resident).
#include <algorithm>

#include "rxb_system.h"

namespace rxb {

struct QeqState {       // device-resident CG scalars, double-buffered by iteration parity
  double alpha[2], heta[2], sig_old[2], b_norm[2], dot0[2];
  int active[2], iters[2];
};
struct QeqDev {
  QeqState st[2];
  double dots[3][4];    // rotating accumulators: (u.r)_s, (u.r)_t, (u.w)_s, (u.w)_t
  double pro[6];        // prologue: b.b, u.r, u.w for s and t
  double sums[2];       // sum s, sum t
};

namespace {

constexpr int kWarps = 8;
constexpr int kVecBlocks = 148 * 4;
constexpr int kVecThreads = 256;

__device__ __forceinline__ double2 operator*(double a, double2 b) { return make_double2(a * b.x, a * b.y); }

__global__ void k_qeq_init(int n, const int* __restrict__ type, const AtomPar* __restrict__ atom,
                           const double* __restrict__ s_hist, const double* __restrict__ t_hist, double2* __restrict__ x,
                           double2* __restrict__ b, double* __restrict__ Hdia_inv, QeqDev* __restrict__ Q) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int ti = type[i];
    double eta = 1.0, chi = 0.0;
    if (ti >= 0) { eta = atom[ti].eta; chi = atom[ti].chi; }
    Hdia_inv[i] = 1. / eta;
    b[i] = make_double2(-chi, -1.0);
    const double* sh = s_hist + 5 * (size_t)i;
    const double* th = t_hist + 5 * (size_t)i;
    x[i] = make_double2(4 * (sh[0] + sh[2]) - (6 * sh[1] + sh[3]), th[2] + 3 * (th[0] - th[1]));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int k = 0; k < 3; k++) for (int c = 0; c < 4; c++) Q->dots[k][c] = 0.0;
    for (int c = 0; c < 6; c++) Q->pro[c] = 0.0;
    Q->sums[0] = Q->sums[1] = 0.0;
  }
}

__global__ void k_forward2(int n, int N, const int* __restrict__ owner, double2* __restrict__ vec) {
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < N - n; g += gridDim.x * blockDim.x) {
    const int o = owner[g];
    if (o >= 0) vec[n + g] = vec[o];
  }
}

// y[i] = eta_i x_i + sum_j H_ij x_j for local rows; gate != null: skip when neither system is active
__global__ void __launch_bounds__(kWarps * 32)
k_spmv2(int n, const long long* __restrict__ off, const int* __restrict__ num, const int* __restrict__ col,
        const double* __restrict__ val, const int* __restrict__ type, const AtomPar* __restrict__ atom,
        const double2* __restrict__ x, double2* __restrict__ y, const QeqDev* __restrict__ Q, int parity) {
  if (Q != nullptr && !(Q->st[parity].active[0] | Q->st[parity].active[1])) return;
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
  for (int i = wg; i < n; i += nwg) {
    const long long beg = off[i];
    const int m = num[i];
    double ax = 0, ay = 0;
    for (int k = lane; k < m; k += 32) {
      // H is streamed exactly once per SpMV: evict-first loads keep the gathered x vector (16 B/atom) resident in L2/L1
      const double h = __ldcs(val + beg + k);
      const double2 xj = __ldg(x + __ldcs(col + beg + k));
      ax += h * xj.x; ay += h * xj.y;
    }
    ax = warp_sum(ax); ay = warp_sum(ay);
    if (lane == 0) {
      const int ti = type[i];
      const double eta = ti >= 0 ? atom[ti].eta : 0.0;
      const double2 xi = x[i];
      y[i] = make_double2(eta * xi.x + ax, eta * xi.y + ay);
    }
  }
}

// prologue steps (fix_qeq_reax_sunway.cpp:1024-1105)
__global__ void k_pro1(int n, const double2* __restrict__ b, const double2* __restrict__ q, const double* __restrict__ Hd,
                       double2* __restrict__ r, double2* __restrict__ u, double2* __restrict__ d) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double2 bj = b[j], qj = q[j];
    const double2 rj = make_double2(bj.x - 1 * qj.x, bj.y - 1 * qj.y);
    const double2 uj = make_double2(rj.x * Hd[j], rj.y * Hd[j]);
    r[j] = rj; u[j] = uj; d[j] = uj;
  }
}
__global__ void k_pro2(int n, const double2* __restrict__ q, const double* __restrict__ Hd, double2* __restrict__ w,
                       double2* __restrict__ m, double2* __restrict__ d) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double2 qj = q[j];
    const double2 mj = make_double2(qj.x * Hd[j], qj.y * Hd[j]);
    w[j] = qj; m[j] = mj; d[j] = mj;
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
__global__ void __launch_bounds__(kVecThreads)
k_cg_sweep(int n, int it, int first, double tol, int imax, const double* __restrict__ Hd, const double2* __restrict__ q,
           double2* __restrict__ x, double2* __restrict__ r, double2* __restrict__ u, double2* __restrict__ w,
           double2* __restrict__ p, double2* __restrict__ ss, double2* __restrict__ v, double2* __restrict__ z,
           double2* __restrict__ d, QeqDev* __restrict__ Q) {
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
      double2 xj = x[j], rj = r[j], uj = u[j], wj = w[j], pj = p[j], sj = ss[j], vj = v[j], zj = z[j], dj = d[j];
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
        r[j] = rj; u[j] = uj; w[j] = wj; d[j] = dj;
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

__global__ void k_q_final(int n, int N, const int* __restrict__ owner, const double2* __restrict__ x,
                          const QeqDev* __restrict__ Q, double* __restrict__ s_hist, double* __restrict__ t_hist,
                          double4* __restrict__ xq, int phase) {
  if (phase == 0) {
    const double uu = Q->sums[0] / Q->sums[1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const double2 xi = x[i];
      xq[i].w = xi.x - uu * xi.y;
      double* sh = s_hist + 5 * (size_t)i;
      double* th = t_hist + 5 * (size_t)i;
      for (int k = 4; k > 0; --k) { sh[k] = sh[k - 1]; th[k] = th[k - 1]; }
      sh[0] = xi.x; th[0] = xi.y;
    }
  } else {
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < N - n; g += gridDim.x * blockDim.x) {
      const int o = owner[g];
      if (o >= 0) xq[n + g].w = xq[o].w;
    }
  }
}

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
  RXB_CUDA(cudaStreamSynchronize(st_));
}
void System::qeq_get_history(double* s_hist, double* t_hist) {
  RXB_CUDA(cudaMemcpyAsync(s_hist, q_s_hist.p, (size_t)5 * n * sizeof(double), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaMemcpyAsync(t_hist, q_t_hist.p, (size_t)5 * n * sizeof(double), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaStreamSynchronize(st_));
}

void System::qeq_pre_force() {
  if (n == 0) return;
  if (q_s_hist.n != (size_t)5 * n && !dist_) qeq_reset_history();
  last_swb_ = qeq_swb;
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
  const size_t nn = n, NN = std::max((size_t)N, slab());  // all-gathered arrays must hold a whole slab
  q_x.resize(NN); q_d.resize(NN);
  q_r.resize(nn); q_u.resize(nn); q_w.resize(nn); q_p.resize(nn); q_ss.resize(nn); q_v.resize(nn); q_z.resize(nn);
  q_q.resize(nn); q_b.resize(nn); q_m.resize(nn); q_Hdia_inv.resize(nn);
  q_scal.resize(sizeof(QeqDev) / sizeof(double) + 8);
  QeqDev* Q = reinterpret_cast<QeqDev*>(q_scal.p);
  const int nghost = N - n;
  const int fb = nghost > 0 ? (nghost + 255) / 256 : 0;
  auto forward = [&](double2* vec) {
    if (dist_) dist_forward2(vec);
    else if (fb) { k_forward2<<<fb, 256, 0, st_>>>(n, N, ghost_owner.p, vec); kernel_launches++; }
  };
  auto spmv = [&](const double2* x, double2* y, const QeqDev* gate, int parity) {
    const int ts = tick(StepTimers::SPMV);
    // one row per warp, blocks retire continuously: lets the high-priority bond-chain stream interleave on every SM
    k_spmv2<<<std::max(1, (n + kWarps - 1) / kWarps), kWarps * 32, 0, st_>>>(n, vl.off.p, far_num.p, far_idx.p, H_val.p, type.p, dp_.atom, x, y, gate, parity);
    tock(ts);
    kernel_launches++;
  };
  k_qeq_init<<<kVecBlocks, kVecThreads, 0, st_>>>(n, type.p, dp_.atom, q_s_hist.p, q_t_hist.p, q_x.p, q_b.p, q_Hdia_inv.p, Q);
  forward(q_x.p);
  spmv(q_x.p, q_q.p, nullptr, 0);
  k_pro1<<<kVecBlocks, kVecThreads, 0, st_>>>(n, q_b.p, q_q.p, q_Hdia_inv.p, q_r.p, q_u.p, q_d.p);
  forward(q_d.p);
  spmv(q_d.p, q_q.p, nullptr, 0);
  k_pro2<<<kVecBlocks, kVecThreads, 0, st_>>>(n, q_q.p, q_Hdia_inv.p, q_w.p, q_m.p, q_d.p);
  forward(q_d.p);
  spmv(q_d.p, q_q.p, nullptr, 0);
  k_pro3<<<kVecBlocks, kVecThreads, 0, st_>>>(n, q_b.p, q_r.p, q_u.p, q_w.p, q_m.p, q_q.p, q_p.p, q_ss.p, q_v.p, q_z.p, Q);
  if (dist_) dist_allreduce(Q->pro, 6);
  k_scal_init<<<1, 32, 0, st_>>>(Q, qeq_tol, qeq_imax);
  kernel_launches += 5;

  // main loop: sweep(it) ; halo ; SpMV.  The sweep after the last active iteration applies the final x update.
  int active_host[2] = {1, 1};
  int it = 1;
  for (; it <= qeq_imax + 1; it++) {
    k_cg_sweep<<<kVecBlocks, kVecThreads, 0, st_>>>(n, it, it == 1, qeq_tol, qeq_imax, q_Hdia_inv.p, q_q.p, q_x.p, q_r.p,
                                                   q_u.p, q_w.p, q_p.p, q_ss.p, q_v.p, q_z.p, q_d.p, Q);
    kernel_launches++;
    const int par_next = (it & 1) ^ 1;  // state written by this sweep (from the dot products of the sweep before it)
    if (it % qeq_check_every == 0 || it > qeq_imax) {
      RXB_CUDA(cudaMemcpyAsync(active_host, Q->st[par_next].active, 2 * sizeof(int), cudaMemcpyDeviceToHost, st_));
      RXB_CUDA(cudaStreamSynchronize(st_));
      if (!(active_host[0] | active_host[1])) break;
    }
    // MPI_Allreduce(dot_local, 2) of each solve (:1132) and the boundary values of d travel in one exchange
    if (dist_) dist_forward2_dots(q_d.p, Q->dots[(it + 1) % 3]);
    else forward(q_d.p);
    spmv(q_d.p, q_q.p, Q, par_next);
  }
  // final charges
  k_q_sums<<<kVecBlocks, kVecThreads, 0, st_>>>(n, q_x.p, Q);
  if (dist_) dist_allreduce(Q->sums, 2);
  k_q_final<<<kVecBlocks, kVecThreads, 0, st_>>>(n, N, ghost_owner.p, q_x.p, Q, q_s_hist.p, q_t_hist.p, xq.p, 0);
  if (dist_) dist_forward_xq();
  else if (fb) k_q_final<<<fb, 256, 0, st_>>>(n, N, ghost_owner.p, q_x.p, Q, q_s_hist.p, q_t_hist.p, xq.p, 1);
  kernel_launches += 3;
  int iters_host[2];
  const int par_final = (it & 1) ^ 1;
  RXB_CUDA(cudaMemcpyAsync(iters_host, Q->st[par_final].iters, 2 * sizeof(int), cudaMemcpyDeviceToHost, st_));
  RXB_CUDA(cudaStreamSynchronize(st_));
  matvecs_s = iters_host[0];
  matvecs_t = iters_host[1];
  qeq_iters_total += (matvecs_s > matvecs_t ? matvecs_s : matvecs_t);
  qeq_ran_this_step_ = true;
  tock(t_QEQ_CG);
}

}  // namespace rxb
