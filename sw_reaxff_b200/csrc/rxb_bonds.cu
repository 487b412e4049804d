// Bond list (uncorrected BO') and bond-order correction kernels: one warp per atom, lanes over neighbours/bonds.
//
// Semantics restated from the reference (the *arithmetic* follows its serial forms so results agree to fp64
// round-off; the *structure* is new):
//   K-bond  : Init_Forces_noQEq_Full + BOp_single   /root/reference/reaxc_forces_sunway.cpp:677-825,852-935
//             row order = ascending neighbour index  /root/reference/reaxc_forces_sw64.c:31-75,1289
//             fused with BO section 1 (Deltap)       /root/reference/reaxc_bond_orders_sw64.c:25-66
//   K-bo    : BO sections 2+3 fused                  /root/reference/reaxc_bond_orders_sw64.c:67-484
// Differences by design: rows are carved from one atomic cursor instead of a prefix sum over last step's
// 2x-padded estimates (reaxc_reset_tools_sunway.cpp:122-190); candidates come from a (bond_cut+skin) Verlet
// list instead of the 64-byte far-neighbour records; dBOp / dln_BOp_pi / dln_BOp_pi2 are stored as one scalar
// each (they are all scalar * dvec).  Bounded by fp64 pow/exp throughput, not HBM.
#include "rxb_system.h"

namespace rxb {
namespace {

constexpr int kWarps = 4;      // warps per block
// Bonds per atom held in shared staging: v.row_cap entries per warp (dynamic shared memory; 64 to start with - the
// reference allows 35 while building and 20 later).  A row that outgrows it is not fatal: the kernel raises overflow bit 1,
// the host doubles row_cap and replays the force phase (System::compute), like every other list of this path.
constexpr int kStageDoubles = 11;   // d,dx,dy,dz, BO,BO_s,BO_pi,BO_pi2, cBOp,cPi,cPi2
inline size_t bond_stage_bytes(int row_cap) {
  return (size_t)kWarps * ((size_t)row_cap * (kStageDoubles * sizeof(double) + sizeof(int)) + 64 * sizeof(int));
}

__global__ void __launch_bounds__(kWarps * 32, 7)   // 6 / 7 / 8 CTAs per SM measured: 0.692 / 0.676 / 0.689 ms
k_bond_list(DevView v, DevParams P) {
  extern __shared__ __align__(16) double bl_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int row_cap = v.row_cap;
  double* S_val = bl_smem + (size_t)wib * row_cap * kStageDoubles;                       // [row_cap][11]
  int* S_nbr = reinterpret_cast<int*>(bl_smem + (size_t)kWarps * row_cap * kStageDoubles) + (size_t)wib * row_cap;
  int* s_queue = reinterpret_cast<int*>(bl_smem + (size_t)kWarps * row_cap * kStageDoubles) + (size_t)kWarps * row_cap;
  const int i = blockIdx.x * kWarps + wib;
  if (i >= v.N) return;
  const int ti = v.type[i];
  int cnt = 0;
  if (ti >= 0) {
    const double4 pi = v.xq[i];
    const AtomPar ai = P.atom[ti];
    const double bo_cut = P.ctl.bo_cut, bond_cut = P.ctl.bond_cut, nonb_cut2 = P.ctl.nonb_cut * P.ctl.nonb_cut;
    const long long beg = v.bc_off[i], end = beg + v.bc_cnt[i];
    int* queue = s_queue + wib * 64;
    int qn = 0;
    const float4 fi = v.xf[i];
    // per partner element: (reach of a bond of this pair)^2 + fp32 rounding band; lane t holds the threshold of element t
    const int nt = P.nt;
    float thr_lane = 0.0f;
    if (lane < nt) {
      const double dm = P.pair[ti * nt + lane].d_bond_max;
      thr_lane = __double2float_ru(dm * dm) + v.bond_band;
    }
    // Phase 1 (cheap, all lanes): distance filter, survivors are queued.  Phase 2 (6 transcendentals per pair) runs on
    // FULL warps drained from the queue: only ~30 % of the (bond_cut + skin) candidates are inside bond_cut, so doing the
    // math in place would leave two thirds of the lanes idle.
    // the candidate index of chunk c+2 and the shadow-position gather of chunk c+1 are issued while chunk c is tested
    int j_cur = beg + lane < end ? v.bc_idx[beg + lane] : -1;
    int j_nxt = beg + 32 + lane < end ? v.bc_idx[beg + 32 + lane] : -1;
    float4 f_cur = j_cur >= 0 ? v.xf[j_cur] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long k0 = beg; k0 < end || qn > 0; k0 += 32) {
      if (k0 < end) {
        const int j_nn = k0 + 64 + lane < end ? v.bc_idx[k0 + 64 + lane] : -1;
        const float4 f_nxt = j_nxt >= 0 ? v.xf[j_nxt] : make_float4(0.f, 0.f, 0.f, 0.f);
        bool near = false;
        const int j = j_cur;
        int tjf = -1;
        float r2f = 0.0f;
        if (j >= 0) {
          // fp32 shadow (position + type in 16 bytes): a conservative superset goes to the queue, the exact fp64 tests
          // (d <= bond_cut, BO' >= bo_cut) are applied when the candidate is drained
          const float4 fj = f_cur;
          const float ex = fj.x - fi.x, ey = fj.y - fi.y, ez = fj.z - fi.z;
          r2f = ex * ex + ey * ey + ez * ez;
          tjf = __float_as_int(fj.w);
        }
        j_cur = j_nxt; f_cur = f_nxt; j_nxt = j_nn;
        float thr = __shfl_sync(0xffffffffu, thr_lane, tjf & 31);
        if (tjf >= 32) { const double dm = P.pair[ti * nt + tjf].d_bond_max; thr = __double2float_ru(dm * dm) + v.bond_band; }
        near = tjf >= 0 && r2f <= thr;
        const unsigned m = __ballot_sync(0xffffffffu, near);
        if (near) queue[qn + __popc(m & ((1u << lane) - 1))] = j;
        qn += __popc(m);
        __syncwarp();
        if (qn < 32 && k0 + 32 < end) continue;   // keep filling until a full warp of work (or the row ends)
      }
      // ---- drain up to 32 queued candidates ----
      const int take = min(qn, 32);
      bool hit = false;
      int j = -1;
      double d = 0, dx = 0, dy = 0, dz = 0, BO = 0, BO_s = 0, BO_pi = 0, BO_pi2 = 0, cBOp = 0, cPi = 0, cPi2 = 0;
      if (lane < take) {
        j = queue[lane];
        const double4 pj = v.xq[j];
        dx = pj.x - pi.x; dy = pj.y - pi.y; dz = pj.z - pi.z;
        const double r2 = dx * dx + dy * dy + dz * dz;
        d = sqrt(r2);
        const int tj = v.type[j];
        const bool inside = r2 <= nonb_cut2 && d <= bond_cut;   // exact test (the queue holds an fp32 superset)
        const AtomPar& aj = P.atom[tj];
        const PairPar& tw = P.pair[ti * P.nt + tj];
        // (d/r)^p as exp(p (log d - log r)): one log shared by the three terms instead of three pow() calls
        // (agrees with pow to ~1e-15 relative; the reference's own CPE kernels use polynomial exp/pow)
        double C12 = 0, C34 = 0, C56 = 0;
        const double ld = log(d);
        if (ai.r_s > 0.0 && aj.r_s > 0.0) { C12 = tw.p_bo1 * exp(tw.p_bo2 * (ld - tw.log_r_s)); BO_s = (1.0 + bo_cut) * exp(C12); }
        if (ai.r_pi > 0.0 && aj.r_pi > 0.0) { C34 = tw.p_bo3 * exp(tw.p_bo4 * (ld - tw.log_r_p)); BO_pi = exp(C34); }
        if (ai.r_pi_pi > 0.0 && aj.r_pi_pi > 0.0) { C56 = tw.p_bo5 * exp(tw.p_bo6 * (ld - tw.log_r_pp)); BO_pi2 = exp(C56); }
        BO = BO_s + BO_pi + BO_pi2;
        if (inside && BO >= bo_cut) {
          hit = true;
          const double rr2 = d * d;
          const double Cln_s = tw.p_bo2 * C12 / rr2, Cln_pi = tw.p_bo4 * C34 / rr2, Cln_pi2 = tw.p_bo6 * C56 / rr2;
          cBOp = -(BO_s * Cln_s + BO_pi * Cln_pi + BO_pi2 * Cln_pi2);
          cPi = -BO_pi * Cln_pi;
          cPi2 = -BO_pi2 * Cln_pi2;
          BO_s -= bo_cut;
          BO -= bo_cut;
        }
      }
      // shift the remainder of the queue down
      const int rest = qn - take;
      int moved = -1;
      if (lane < rest) moved = queue[take + lane];
      __syncwarp();
      if (lane < rest) queue[lane] = moved;
      qn = rest;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int slot = cnt + __popc(m & ((1u << lane) - 1));
        if (slot < row_cap) {
          S_nbr[slot] = j;
          double* o = S_val + (size_t)slot * kStageDoubles;
          o[0] = d; o[1] = dx; o[2] = dy; o[3] = dz; o[4] = BO; o[5] = BO_s; o[6] = BO_pi; o[7] = BO_pi2;
          o[8] = cBOp; o[9] = cPi; o[10] = cPi2;
        }
      }
      cnt += __popc(m);
      __syncwarp();
    }
  }
  if (cnt > row_cap) { if (lane == 0) { atomicOr(v.overflow, 1); atomicMax(v.need_row, cnt); } cnt = row_cap; }
  __syncwarp();
  // carve the row
  int start = 0;
  if (lane == 0) {
    start = atomicAdd(v.b_cursor, cnt);
    // a row that does not fit the bond arrays is published EMPTY: the kernels that still run before the host grows the
    // arrays and replays the step must not walk rows that were never written (the cursor keeps the true total)
    const bool fits_row = start + cnt <= v.cap_bonds;
    if (!fits_row) atomicOr(v.overflow, 2);
    v.b_start[i] = fits_row ? start : 0;
    v.b_cnt[i] = fits_row ? cnt : 0;
  }
  start = __shfl_sync(0xffffffffu, start, 0);
  const bool fits = start + cnt <= v.cap_bonds;
  // rank sort by neighbour index, write out, and reduce in rank order (deterministic)
  double sBO = 0, sx = 0, sy = 0, sz = 0;
  for (int e0 = 0; e0 < cnt; e0 += 32) {
    const int e = e0 + lane;
    if (e < cnt) {
      const int mine = S_nbr[e];
      int rank = 0;
      for (int t = 0; t < cnt; t++) rank += (S_nbr[t] < mine);
      const double* o = S_val + (size_t)e * kStageDoubles;
      if (fits) {
        const int p = start + rank;
        v.b_nbr[p] = mine;
        v.b_owner[p] = i;
        v.b_sym[p] = -1;
        v.b_geo[p] = make_double4(o[0], o[1], o[2], o[3]);
        v.b_bo[p] = make_double4(o[4], o[5], o[6], o[7]);
        v.b_der[p] = make_double4(o[8], o[9], o[10], 0.0);
        v.b_Cdbo[p] = 0.0; v.b_Cdbopi[p] = 0.0; v.b_Cdbopi2[p] = 0.0;
      }
      sBO += o[4];
      sx += o[8] * o[1]; sy += o[8] * o[2]; sz += o[8] * o[3];
    }
  }
  sBO = warp_sum(sBO); sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
  if (lane == 0) {
    v.total_bop[i] = sBO;
    v.dDeltap_self[3 * i] = sx; v.dDeltap_self[3 * i + 1] = sy; v.dDeltap_self[3 * i + 2] = sz;
    double val = 0, val_val = 0;
    if (ti >= 0) { val = P.atom[ti].valency; val_val = P.atom[ti].valency_val; }
    v.Deltap[i] = make_double2(sBO - val, sBO - val_val);
  }
}

// One thread per DIRECTED bond (dense: ~7.5 bonds/atom would leave 3/4 of a warp-per-atom idle in this 8-exp kernel)
__global__ void __launch_bounds__(256)
k_bond_orders(DevView v, DevParams P) {
  const int nb = min(*v.b_cursor, v.cap_bonds);
  const double p_boc1 = P.gp[0], p_boc2 = P.gp[1];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nb; p += gridDim.x * blockDim.x) {
    {
      const int i = v.b_owner[p];
      const int ti = v.type[i];
      const double val_i = P.atom[ti].valency;
      const double2 Dpi = v.Deltap[i];
      const int j = v.b_nbr[p];
      const int tj = v.type[j];
      if (tj < 0) continue;
      // sym_index: position of i in j's (sorted) row
      {
        const int sj = v.b_start[j], cj = v.b_cnt[j];
        int lo = 0, hi = cj - 1, found = -1;
        while (lo <= hi) {
          const int mid = (lo + hi) >> 1;
          const int cand = v.b_nbr[sj + mid];
          if (cand == i) { found = sj + mid; break; }
          if (cand < i) lo = mid + 1; else hi = mid - 1;
        }
        v.b_sym[p] = found;
      }
      const PairPar& tw = P.pair[ti * P.nt + tj];
      double4 bo = v.b_bo[p];  // BO, BO_s, BO_pi, BO_pi2 (uncorrected)
      double C1dbo, C2dbo, C3dbo, C1dbopi, C2dbopi, C3dbopi, C4dbopi, C1dbopi2, C2dbopi2, C3dbopi2, C4dbopi2;
      if (tw.ovc < 0.001 && tw.v13cor < 0.001) {
        C1dbo = 1.0; C2dbo = 0.0; C3dbo = 0.0;
        C1dbopi = bo.z; C2dbopi = 0.0; C3dbopi = 0.0; C4dbopi = 0.0;
        C1dbopi2 = bo.w; C2dbopi2 = 0.0; C3dbopi2 = 0.0; C4dbopi2 = 0.0;
      } else {
        const double val_j = P.atom[tj].valency;
        const double2 Dpj = v.Deltap[j];
        double f1, f4, f5, f4f5, Cf1_ij, Cf1_ji, Cf45_ij, Cf45_ji;
        if (tw.ovc >= 0.001) {
          const double exp_p1i = exp(-p_boc1 * Dpi.x), exp_p2i = exp(-p_boc2 * Dpi.x);
          const double exp_p1j = exp(-p_boc1 * Dpj.x), exp_p2j = exp(-p_boc2 * Dpj.x);
          const double f2 = exp_p1i + exp_p1j;
          const double f3 = -1.0 / p_boc2 * log(0.5 * (exp_p2i + exp_p2j));
          f1 = 0.5 * ((val_i + f2) / (val_i + f2 + f3) + (val_j + f2) / (val_j + f2 + f3));
          const double temp = f2 + f3;
          const double u1_ij = val_i + temp, u1_ji = val_j + temp;
          const double Cf1A_ij = 0.5 * f3 * (1.0 / (u1_ij * u1_ij) + 1.0 / (u1_ji * u1_ji));
          const double Cf1B_ij = -0.5 * ((u1_ij - f3) / (u1_ij * u1_ij) + (u1_ji - f3) / (u1_ji * u1_ji));
          const double e2 = exp_p2i / (exp_p2i + exp_p2j);
          Cf1_ij = 0.50 * (-p_boc1 * exp_p1i / u1_ij - ((val_i + f2) / (u1_ij * u1_ij)) * (-p_boc1 * exp_p1i + e2) +
                           -p_boc1 * exp_p1i / u1_ji - ((val_j + f2) / (u1_ji * u1_ji)) * (-p_boc1 * exp_p1i + e2));
          Cf1_ji = -Cf1A_ij * p_boc1 * exp_p1j + Cf1B_ij * exp_p2j / (exp_p2i + exp_p2j);
        } else {
          f1 = 1.0; Cf1_ij = Cf1_ji = 0.0;
        }
        if (tw.v13cor >= 0.001) {
          const double exp_f4 = exp(-(tw.p_boc4 * (bo.x * bo.x) - Dpi.y) * tw.p_boc3 + tw.p_boc5);
          const double exp_f5 = exp(-(tw.p_boc4 * (bo.x * bo.x) - Dpj.y) * tw.p_boc3 + tw.p_boc5);
          f4 = 1. / (1. + exp_f4);
          f5 = 1. / (1. + exp_f5);
          f4f5 = f4 * f5;
          Cf45_ij = -f4 * exp_f4;
          Cf45_ji = -f5 * exp_f5;
        } else {
          f4 = f5 = f4f5 = 1.0; Cf45_ij = Cf45_ji = 0.0;
        }
        const double A0_ij = f1 * f4f5;
        const double A1_ij = -2 * tw.p_boc3 * tw.p_boc4 * bo.x * (Cf45_ij + Cf45_ji);
        const double A2_ij = Cf1_ij / f1 + tw.p_boc3 * Cf45_ij;
        const double A2_ji = Cf1_ji / f1 + tw.p_boc3 * Cf45_ji;
        const double A3_ij = A2_ij + Cf1_ij / f1;
        const double A3_ji = A2_ji + Cf1_ji / f1;
        bo.x = bo.x * A0_ij;
        bo.z = bo.z * A0_ij * f1;
        bo.w = bo.w * A0_ij * f1;
        bo.y = bo.x - (bo.z + bo.w);
        C1dbo = A0_ij + bo.x * A1_ij; C2dbo = bo.x * A2_ij; C3dbo = bo.x * A2_ji;
        C1dbopi = f1 * f1 * f4 * f5; C2dbopi = bo.z * A1_ij; C3dbopi = bo.z * A3_ij; C4dbopi = bo.z * A3_ji;
        C1dbopi2 = f1 * f1 * f4 * f5; C2dbopi2 = bo.w * A1_ij; C3dbopi2 = bo.w * A3_ij; C4dbopi2 = bo.w * A3_ji;
      }
      if (bo.x < 1e-10) bo.x = 0.0;
      if (bo.y < 1e-10) bo.y = 0.0;
      if (bo.z < 1e-10) bo.z = 0.0;
      if (bo.w < 1e-10) bo.w = 0.0;
      v.b_bo[p] = bo;
      v.b_c1[p] = make_double4(C1dbo, C2dbo, C3dbo, C1dbopi);
      v.b_c2[p] = make_double4(C2dbopi, C3dbopi, C4dbopi, C1dbopi2);
      v.b_c3[p] = make_double4(C2dbopi2, C3dbopi2, C4dbopi2, 0.0);
    }
  }
}

// BO section 3 (reaxc_bond_orders_sw64.c:449-482): one thread per atom, row sum in slot order (deterministic)
__global__ void __launch_bounds__(256)
k_bond_order_atoms(DevView v, DevParams P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.N) return;
  const int ti = v.type[i];
  double tot = 0.0;
  {
    const int start = v.b_start[i], cnt = v.b_cnt[i];
    for (int e = 0; e < cnt; e++) tot += v.b_bo[start + e].x;
  }
  {
    v.total_bo[i] = tot;
    if (ti >= 0) {
      const AtomPar& a = P.atom[ti];
      const double p_lp1 = P.gp[15];
      const double Delta_e = tot - a.valency_e;
      const double vlpex = Delta_e - 2.0 * (int)(Delta_e / 2.0);
      const double explp1 = exp(-p_lp1 * (2.0 + vlpex) * (2.0 + vlpex));
      const double nlp = explp1 - (int)(Delta_e / 2.0);
      const double Clp = 2.0 * p_lp1 * explp1 * (2.0 + vlpex);
      v.Delta[i] = tot - a.valency;
      v.Delta_boc[i] = tot - a.valency_boc;
      v.Delta_val[i] = tot - a.valency_val;
      v.vlpex[i] = vlpex;
      v.nlp[i] = nlp;
      v.Delta_lp[i] = a.nlp_opt - nlp;
      v.dDelta_lp[i] = Clp;
      const double nlp_temp = (a.mass > 21.0) ? 0.5 * (a.valency_e - a.valency) : nlp;
      v.Delta_lp_temp[i] = a.nlp_opt - nlp_temp;
    } else {
      v.Delta[i] = 0; v.Delta_boc[i] = 0; v.Delta_val[i] = 0; v.vlpex[i] = 0; v.nlp[i] = 0; v.Delta_lp[i] = 0;
      v.dDelta_lp[i] = 0; v.Delta_lp_temp[i] = 0;
    }
  }
}

}  // namespace

void launch_bond_list(System& s, DevView& v, const DevParams& P, cudaStream_t st) {
  RXB_CUDA(cudaMemsetAsync(v.b_cursor, 0, sizeof(int), st));
  if (v.N == 0) return;
  const size_t smem = bond_stage_bytes(v.row_cap);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    RXB_CUDA(cudaFuncSetAttribute(k_bond_list, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  k_bond_list<<<(v.N + kWarps - 1) / kWarps, kWarps * 32, smem, st>>>(v, P);
  s.kernel_launches++;
}

void launch_bond_orders(System& s, DevView& v, const DevParams& P, cudaStream_t st) {
  if (v.N == 0) return;
  static int occ = 0;
  k_bond_orders<<<wave_grid(k_bond_orders, 256, 2 * chain_waves(), occ), 256, 0, st>>>(v, P);
  k_bond_order_atoms<<<(v.N + 255) / 256, 256, 0, st>>>(v, P);
  s.kernel_launches += 2;
}

}  // namespace rxb
