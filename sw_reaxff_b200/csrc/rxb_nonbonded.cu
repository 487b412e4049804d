// Per-step far-neighbour filter fused with the QEq H-matrix build, and the tapered vdW/Coulomb kernel.
//
//   K-farH : write_reax_lists_c (r <= nonb_cut filter)   /root/reference/pair_reaxc_sw64.c:49-95,193-341
//            + compute_H_Full_C / calculate_H            /root/reference/fix_qeq_reax_sw64.c:147-189, fix_qeq_reax_sunway.cpp:965-981
//            The reference materialises 64-byte far_neighbor_data_full records (29 KB/atom/step) AND a separate
//            600-wide H row; here both are ONE compacted list in the slots of the Verlet row - a 64-bit word per entry
//            (22-bit column + 42-bit fixed-point H value; or int32 column + fp64 value in the exact format) - because
//            "r <= nonb_cut" and "r <= swb" select the same pairs (both 10 A in every shipped input; if they differ the
//            list takes the larger cut-off and each consumer re-tests its own).  One pass (k_far_H1): exact gather, fp64
//            decision, H value, warp compaction; the hydrogen-bond candidates ride on the same sweep.
//   K-nb   : vdW_Coulomb_Energy_Full_C                   /root/reference/reaxc_nonbonded_sw64.c:40-258 (serial twin)
//            full list, local i only, force on i only (no scatter), 1/2 energy per directed pair,
//            pair virial + (-x_i (x) f_i) correction as reaxc_nonbonded_cpe.h:531-536 / reaxc_nonbonded_sw64.c:247-252.
// Both kernels work in S space (rxb_dev.cuh): rows are the local atoms in cell-sorted order, columns are sorted positions,
// so the per-pair gathers (16-byte shadow per candidate, 32-byte record per survivor) fall into runs of consecutive
// addresses instead of one sector per lane (round 1: L1 data pipe 78 % busy, DRAM 23 %).
// Roofline: K-farH streams 4 B/Verlet entry in and 8 B (packed) or 12 B (exact) per far entry out (HBM target); K-nb is
// fp64-issue bound (143 DP instructions per pair: 2 log + 3 exp + cube root + rsqrt + 1 division, all from rxb_math.cuh).
// Numbers: DESIGN.md 3.
#include <cstdlib>
#include <type_traits>

#include "rxb_math.cuh"
#include "rxb_system.h"

namespace rxb {
namespace {

constexpr double kCele = 332.06371;  // C_ele
constexpr double kEvToKcal = 14.4;   // EV_TO_KCAL_PER_MOL
constexpr double kKcalToEv = 23.02;  // KCALpMOL_to_EV
constexpr int kWarps = 8;

__device__ __forceinline__ double dist2_rn(double dx, double dy, double dz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

struct QeqConst { double Tap[8]; double swb2; double far2; double inner_lim2; double h_quant; };

// Packed H entry: column (22 bits) << 42 | round(H * 2^shift) in 42 bits.  H = Tap(r) * 14.4 / cbrt(r^3 + shld) lies in
// [0, 14.4 / cbrt(min shld)] for a taper that starts at 0 (Tap in [0,1]); the host picks the shift so that this bound fits, so
// the absolute quantisation error is 2^-(shift+1) (1.8e-12 for the TATB force field) - relative to the diagonal eta ~ 7
// that is 3e-13, the size of the cube-root identity already in use, and five orders below the CG tolerance.
__device__ __forceinline__ unsigned long long h_pack(int col, double val, double quant) {
  long long m = __double2ll_rn(val * quant);
  m = m < 0 ? 0 : (m > (long long)kHValMask ? (long long)kHValMask : m);
  return ((unsigned long long)(unsigned)col << kHColShift) | (unsigned long long)m;
}

template <bool PACKED>
__global__ void __launch_bounds__(kWarps * 32)
k_far_H(DevView v, int nt, QeqConst qc, const double* __restrict__ shld, const AtomPar* __restrict__ atom, double hbond_cut,
        double hbond_r2max, BondedWork W) {
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
  const bool inner_ok = *v.disp2 <= qc.inner_lim2;
  const int stride = v.vl_stride;
  for (int r = wg; r < v.n; r += nwg) {
    const int kself = v.rowpos[r];
    const double4 pi = v.xqs[kself];
    const float4 fi = v.xs[kself];
    const int ti = __float_as_int(fi.w);
    const long long beg = (long long)r * stride;
    // pass 1: distance filter + ballot compaction of the column indices (all lanes busy, no transcendental work).
    // The chain index load -> position gather is pure latency, so four chunks of 32 candidates are kept in flight:
    // all index loads are issued first, then all gathers, then the tests.  Columns are sorted positions: the candidates of
    // a row are runs of consecutive integers, so the 32 shadow gathers of a chunk fall into a handful of 128-byte lines.
    const int* __restrict__ vl = v.vl_idx + beg;
    // compacted columns: exact format -> far_idx row; packed format -> the upper half of the row's own 64-bit slots (read
    // back by pass 2 before the packed words, written from the bottom, reach them: see pass 2)
    int* far = PACKED ? reinterpret_cast<int*>(v.hpk + beg) + stride : v.far_idx + beg;   // (aliases hpk: no __restrict__)
    // adaptive inner skin: while no atom has moved more than half the margin since the build, every pair inside the far
    // cut-off is in the inner block of its row
    const int cnt_i = inner_ok ? v.vl_cnt_in[r] : v.vl_cnt[r];
    int w = 0;
    const float lo2 = (float)qc.far2 - v.far_band, hi2 = (float)qc.far2 + v.far_band;
    const unsigned lt_mask = (1u << lane) - 1;
    constexpr int kU = 4;   // chunks in flight per lane (A/B: 2 -> 0.838 ms, 4 -> 0.815 ms, 8 -> 1.26 ms at 102 registers)
    for (int k0 = 0; k0 < cnt_i; k0 += 32 * kU) {
      int jj[kU];
      float4 fj[kU];
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int k = k0 + 32 * u + lane;
        jj[u] = k < cnt_i ? __ldcs(vl + k) : -1;
      }
#pragma unroll
      for (int u = 0; u < kU; u++) fj[u] = jj[u] >= 0 ? v.xs[jj[u]] : make_float4(1e18f, 1e18f, 1e18f, 0.f);
#pragma unroll
      for (int u = 0; u < kU; u++) {
        // 16-byte fp32 shadow first; only distances inside the rounding band load the exact 32-byte record, so the
        // r^2 <= cut^2 decision is still the fp64 one.  Idle lanes carry a far-away dummy position.
        const float ex = fj[u].x - fi.x, ey = fj[u].y - fi.y, ez = fj[u].z - fi.z;
        const float r2f = ex * ex + ey * ey + ez * ez;
        bool hit = r2f < lo2;
        if (!hit && r2f <= hi2) {
          const double4 pj = v.xqs[jj[u]];
          hit = dist2_rn(pj.x - pi.x, pj.y - pi.y, pj.z - pi.z) <= qc.far2;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) far[w + __popc(m & lt_mask)] = jj[u];
        w += __popc(m);
      }
    }
    const int num = w;
    if (lane == 0) v.far_num[r] = num;
    __syncwarp();
    // pass 2: H values on the compacted row (every lane does the taper + cube root; xqs[j] is an L1/L2 hit).
    // Rows of hydrogen atoms also emit their hydrogen-bond partner candidates here (acceptor-type j within hbond_cut:
    // Init_Forces_noQEq_HB_Full_C, reaxc_forces_sw64.c:787-863) while x_j, type_j and r are in registers.
    // Packed format: word k of the row overwrites the int slots 2k, 2k+1, i.e. compacted columns 2k - stride and
    // 2k + 1 - stride <= k: columns of this or an earlier chunk, all of which are in registers by then (the __syncwarp
    // orders the chunk's loads before its stores).
    const bool is_H = ti >= 0 && atom[ti].p_hbond == 1 && hbond_cut > 0.0;
    const int lti = v.shld_lt ? v.ltype_s[kself] : 0;
    const int i_atom = is_H ? v.row_atom[r] : 0;
    constexpr int kV = 2;
    for (int k0 = 0; k0 < num; k0 += 32 * kV) {
      int jj[kV];
      double4 pjv[kV];
      int tjv[kV];
#pragma unroll
      for (int u = 0; u < kV; u++) {
        const int k = k0 + 32 * u + lane;
        jj[u] = k < num ? far[k] : -1;
      }
      if (PACKED) __syncwarp();
#pragma unroll
      for (int u = 0; u < kV; u++) {
        pjv[u] = jj[u] >= 0 ? v.xqs[jj[u]] : make_double4(0, 0, 0, 0);
        tjv[u] = jj[u] >= 0 ? v.type_s[jj[u]] : -1;
      }
#pragma unroll
      for (int u = 0; u < kV; u++) {
        const int k = k0 + 32 * u + lane;
        bool cand = false;
        const int j = jj[u];
        if (j >= 0) {
          const double4 pj = pjv[u];
          const double r2 = dist2_rn(pj.x - pi.x, pj.y - pi.y, pj.z - pi.z);
          const int tj = tjv[u];
          double val = 0.0;
          if (ti >= 0 && tj >= 0) {
            // r to 1 ulp from rsqrt (coincident atoms: r = 0 as sqrt gives, not NaN), the 7-op cube root of rxb_math.cuh,
            // and the hydrogen-bond reach tested on r^2 against the largest r^2 whose correctly rounded root is
            // <= hbond_cut (the same decision as sqrt(r2) <= cut)
            const double r = r2 > 0.0 ? r2 * rsqrt(r2) : 0.0;
            if (r2 <= qc.swb2) {
              double T = qc.Tap[7] * r + qc.Tap[6];
              T = T * r + qc.Tap[5]; T = T * r + qc.Tap[4]; T = T * r + qc.Tap[3];
              T = T * r + qc.Tap[2]; T = T * r + qc.Tap[1]; T = T * r + qc.Tap[0];
              const double x3 = r2 * r + (v.shld_lt ? v.shld_lt[lti * v.nlt + v.ltype_s[j]] : shld[ti * nt + tj]);
              // reference: Taper * 14.4 / pow(r^3 + shld, 0.3333333333333); the cube root differs by < 3e-13 relative
              val = T * kEvToKcal * fm::rcbrt_b(x3);
            }
            cand = is_H && atom[tj].p_hbond == 2 && r2 <= hbond_r2max;
          }
          if (PACKED) v.hpk[beg + k] = h_pack(j, val, qc.h_quant);
          else v.H_val[beg + k] = val;
        }
        if (is_H) {
          const unsigned m = __ballot_sync(0xffffffffu, cand);
          if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(W.n_hb, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (cand) {
              const int o = base + __popc(m & ((1u << lane) - 1));
              if (o < W.cap_hb) W.hb[o] = make_int4(i_atom, v.s2a[j], 0, 0);
            }
          }
        }
      }
    }
  }
}

// Single-pass form (default): with the adaptive inner skin 88 % of the candidates a row reads ARE far-list entries, so the
// fp32 prefilter + compaction + second gather of the two-pass kernel above cost more than they save.  Here every candidate
// gathers the exact 32-byte record once, the fp64 r^2 decides (the same decision, bit for bit), the H value is computed for
// the hits and the packed (or exact) entry goes straight to its compacted slot: no column scratch, no second gather, about
// half the instructions per row (ncu r02c: 3178 warp instructions per row, 65 % issue-slot utilisation, for the two-pass form).
// Per-row constants ride in lanes: lane t holds the shielding of (row element, element t) and is read by shuffle; the
// acceptor elements of hydrogen bonds are one bit mask.
template <bool PACKED>
__global__ void __launch_bounds__(kWarps * 32)
k_far_H1(DevView v, int nt, QeqConst qc, const double* __restrict__ shld, const AtomPar* __restrict__ atom, double hbond_cut,
         double hbond_r2max, BondedWork W) {
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
  const bool inner_ok = *v.disp2 <= qc.inner_lim2;
  const int stride = v.vl_stride;
  unsigned acc_mask = 0;                                   // elements that accept hydrogen bonds (p_hbond == 2)
  for (int t = 0; t < nt && t < 32; t++) acc_mask |= (atom[t].p_hbond == 2 ? 1u : 0u) << t;
  const unsigned lt_mask = (1u << lane) - 1;
  for (int r = wg; r < v.n; r += nwg) {
    const int kself = v.rowpos[r];
    const double4 pi = v.xqs[kself];
    const int ti = v.type_s[kself];
    const long long beg = (long long)r * stride;
    const int* __restrict__ vl = v.vl_idx + beg;
    const int cnt_i = inner_ok ? v.vl_cnt_in[r] : v.vl_cnt[r];
    const bool is_H = ti >= 0 && atom[ti].p_hbond == 1 && hbond_cut > 0.0;
    const int i_atom = is_H ? v.row_atom[r] : 0;
    const int lti = v.shld_lt ? v.ltype_s[kself] : 0;
    const double shld_lane = (ti >= 0 && lane < nt) ? shld[ti * nt + lane] : 0.0;
    int w = 0;
    constexpr int kU = 4;   // chunks in flight per lane (A/B: 2 -> 0.838 ms, 4 -> 0.815 ms, 8 -> 1.26 ms at 102 registers)
    for (int k0 = 0; k0 < cnt_i; k0 += 32 * kU) {
      int jj[kU], tjv[kU];
      double4 pjv[kU];
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int k = k0 + 32 * u + lane;
        jj[u] = k < cnt_i ? __ldcs(vl + k) : -1;
      }
#pragma unroll
      for (int u = 0; u < kU; u++) {
        pjv[u] = jj[u] >= 0 ? v.xqs[jj[u]] : make_double4(1e30, 1e30, 1e30, 0.0);   // idle lanes: far away
        tjv[u] = jj[u] >= 0 ? v.type_s[jj[u]] : -1;
      }
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int j = jj[u], tj = tjv[u];
        const double4 pj = pjv[u];
        const double r2 = dist2_rn(pj.x - pi.x, pj.y - pi.y, pj.z - pi.z);
        const bool hit = r2 <= qc.far2;                    // idle lanes carry r2 = 3e60
        // shielding of the pair: a shuffle from the lane that holds element tj (warp-uniform call: every lane takes part)
        double sh = __shfl_sync(0xffffffffu, shld_lane, tj & 31);
        double val = 0.0;
        bool cand = false;
        if (hit && ti >= 0 && tj >= 0) {
          if (v.shld_lt) sh = v.shld_lt[lti * v.nlt + v.ltype_s[j]];        // fix qeq/reax <param file>
          else if (tj >= 32) sh = shld[ti * nt + tj];
          // r to 1 ulp from rsqrt (coincident atoms: r = 0 as sqrt gives, not NaN), the 7-op cube root of rxb_math.cuh
          const double rr = r2 > 0.0 ? r2 * rsqrt(r2) : 0.0;
          if (r2 <= qc.swb2) {
            double T = qc.Tap[7] * rr + qc.Tap[6];
            T = T * rr + qc.Tap[5]; T = T * rr + qc.Tap[4]; T = T * rr + qc.Tap[3];
            T = T * rr + qc.Tap[2]; T = T * rr + qc.Tap[1]; T = T * rr + qc.Tap[0];
            // reference: Taper * 14.4 / pow(r^3 + shld, 0.3333333333333); the cube root differs by < 3e-13 relative
            val = T * kEvToKcal * fm::rcbrt_b(r2 * rr + sh);
          }
          // hydrogen-bond partner: acceptor element within hbond_cut, tested on r^2 against the largest r^2 whose correctly
          // rounded root is <= hbond_cut (the same decision as sqrt(r2) <= cut)
          cand = is_H && ((acc_mask >> (tj & 31)) & 1u) && (tj < 32 || atom[tj].p_hbond == 2) && r2 <= hbond_r2max;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) {
          const int pos = w + __popc(m & lt_mask);
          if (PACKED) {
            // val lies in [0, bound] by construction of the shift (launch_far_and_H); the mask only guards the column bits
            const unsigned long long q = (unsigned long long)__double2ll_rn(fmax(val, 0.0) * qc.h_quant) & kHValMask;
            v.hpk[beg + pos] = ((unsigned long long)(unsigned)j << kHColShift) | q;
          } else {
            v.far_idx[beg + pos] = j;
            v.H_val[beg + pos] = val;
          }
        }
        w += __popc(m);
        if (is_H) {
          const unsigned mc = __ballot_sync(0xffffffffu, cand);
          if (mc) {
            int base = 0;
            if (lane == 0) base = atomicAdd(W.n_hb, __popc(mc));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (cand) {
              const int o = base + __popc(mc & lt_mask);
              if (o < W.cap_hb) W.hb[o] = make_int4(i_atom, v.s2a[j], 0, 0);
            }
          }
        }
      }
    }
    if (lane == 0) v.far_num[r] = w;
  }
}

// K-nb.  One warp per row, one pair per lane, owner computes.  The kernel is bound by fp64 instruction issue (ncu: fp64
// pipe 50 % busy with library exp/log at ~450 instructions per pair), so the work per pair is cut instead:
//   * the straight-line table exp / log / cube root of rxb_math.cuh (11 / 10 / 7 DP ops instead of ~22 / ~35 / ~28),
//     tables staged in shared memory and small enough that a warp's 32 lookups never conflict,
//   * r^p, (r^p + g^-p)^(1/p) and the derivative powers from 2 log + 2 exp (the serial form calls pow 5x), exp1 = exp2^2,
//     1/(r^3 + g) = (cube root)^-3: each identity holds to ~1e-15 relative, far inside the 1e-8 parity tolerance,
//   * the per-type-pair constants sit in shared memory (12 doubles per pair),
//   * the column index of chunk t+2 and the (position, type) gather of chunk t+1 are in flight while chunk t computes.
// 1.97 -> 1.44 ms at 89.5 M pairs.  (Carrying two pairs per lane for more ILP was measured slower: 168 registers, 12 warps/SM.)
constexpr int kNbPar = 12;  // D, alpha, inv_r_vdW, powgi | alpha_over_r_vdW, gamma, r_vdW, rcore | ecore, acore, lgcij, lgre
// launch shape: 256 threads x 2 CTAs per SM (128 registers, 16 warps per SM) by default; RXB_NB_CFG=1 / 2 select 128 threads x
// 5 / 6 CTAs (102 / 85 registers, 20 / 24 warps per SM) for A/B

// column of far-list entry k of a row (packed: the top 22 bits of the 64-bit word)
template <bool PACKED>
__device__ __forceinline__ int far_col(const DevView& v, long long beg, int k) {
  // streamed once per sweep: evict-first, so the column stream does not push the gathered positions out of L1
  if (PACKED) return (int)(__ldcs(v.hpk + beg + k) >> kHColShift);
  return __ldcs(v.far_idx + beg + k);
}

template <bool PACKED> using FarRaw = typename std::conditional<PACKED, unsigned long long, int>::type;
template <bool PACKED>
__device__ __forceinline__ FarRaw<PACKED> far_raw(const DevView& v, long long beg, int k) {
  if constexpr (PACKED) return __ldcs(v.hpk + beg + k);
  else return __ldcs(v.far_idx + beg + k);
}
template <bool PACKED>
__device__ __forceinline__ int raw_col(FarRaw<PACKED> r) {
  if constexpr (PACKED) return (int)(r >> kHColShift);
  else return r;
}

template <bool EV, bool PACKED, int kNbThreads = 256, int kNbCtas = 2>
__global__ void __launch_bounds__(kNbThreads, kNbCtas)
k_nonbonded(DevView v, DevParams P) {
  constexpr int kNbWarps = kNbThreads / 32;
  extern __shared__ __align__(128) double nb_smem[];
  double* tab = nb_smem;                                   // fm::kTabDoubles
  double* red = nb_smem + fm::kTabDoubles;                 // 9 * kNbWarps
  const double2* pt = reinterpret_cast<const double2*>(red + 9 * kNbWarps);   // nt * nt * kNbPar doubles
  const int nt = P.nt;
  for (int t = threadIdx.x; t < fm::kTabDoubles; t += kNbThreads) tab[t] = fm::d_fm_tab[t];
  for (int t = threadIdx.x; t < nt * nt; t += kNbThreads) {
    const PairPar& w = P.pair[t];
    double2* o = reinterpret_cast<double2*>(red + 9 * kNbWarps) + t * (kNbPar / 2);
    o[0] = make_double2(w.D, w.alpha); o[1] = make_double2(w.inv_r_vdW, w.powgi_vdW1);
    o[2] = make_double2(w.alpha_over_r_vdW, w.gamma); o[3] = make_double2(w.r_vdW, w.rcore);
    o[4] = make_double2(w.ecore, w.acore); o[5] = make_double2(w.lgcij, w.lgre);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
  const double p_vdW1 = P.gp[28], p_vdW1i = 1.0 / p_vdW1;
  const double nonb_cut2 = P.ctl.nonb_cut * P.ctl.nonb_cut;
  const int vdw_type = P.ctl.vdw_type, lgflag = P.ctl.lgflag;
  double Tap[8];
#pragma unroll
  for (int t = 0; t < 8; t++) Tap[t] = P.ctl.Tap[t];
  double e_vdw = 0, e_ele = 0, e_pol = 0, vir[6] = {0, 0, 0, 0, 0, 0};
  const int stride = v.vl_stride;
  for (int i = wg; i < v.n; i += nwg) {      // i = row (the i-th local atom in sorted order)
    const int kself = v.rowpos[i];
    const int ti = v.type_s[kself];
    if (ti < 0) continue;
    const double4 pi = v.xqs[kself];
    const long long beg = (long long)i * stride;
    const int num = v.far_num[i];
    double fx = 0, fy = 0, fz = 0;
    // software pipeline: the raw list word of chunk t+2 and the (position, type) gather of chunk t+1 are in flight while
    // chunk t computes.  The word stays RAW (packed: column still in its top bits) until the gather that needs it, so no
    // instruction touches the register of the in-flight load before then.
    double4 p_cur = make_double4(0, 0, 0, 0);
    int t_cur = -1;
    if (lane < num) { const int j0 = far_col<PACKED>(v, beg, lane); p_cur = v.xqs[j0]; t_cur = v.type_s[j0]; }
    FarRaw<PACKED> r_nxt = 32 + lane < num ? far_raw<PACKED>(v, beg, 32 + lane) : FarRaw<PACKED>(0);
    for (int k0 = 0; k0 < num; k0 += 32) {
      const FarRaw<PACKED> r_nn = k0 + 64 + lane < num ? far_raw<PACKED>(v, beg, k0 + 64 + lane) : FarRaw<PACKED>(0);
      double4 p_nxt = make_double4(0, 0, 0, 0);
      int t_nxt = -1;
      if (k0 + 32 + lane < num) { const int jn = raw_col<PACKED>(r_nxt); p_nxt = v.xqs[jn]; t_nxt = v.type_s[jn]; }
      const int tj = t_cur;
      const double4 pj = p_cur;
      p_cur = p_nxt; t_cur = t_nxt; r_nxt = r_nn;
      if (tj < 0) continue;
      const double dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
      const double r2 = dist2_rn(dx, dy, dz);
      if (!(r2 <= nonb_cut2)) continue;
      const double rinv = r2 > 0.0 ? rsqrt(r2) : 0.0;   // coincident atoms: r = 0 like sqrt, not NaN
      const double r_ij = r2 * rinv;
      const double2* w = pt + (ti * nt + tj) * (kNbPar / 2);
      double T = Tap[7] * r_ij + Tap[6];
      T = T * r_ij + Tap[5]; T = T * r_ij + Tap[4]; T = T * r_ij + Tap[3];
      T = T * r_ij + Tap[2]; T = T * r_ij + Tap[1]; T = T * r_ij + Tap[0];
      double dT = 7 * Tap[7] * r_ij + 6 * Tap[6];
      dT = dT * r_ij + 5 * Tap[5]; dT = dT * r_ij + 4 * Tap[4]; dT = dT * r_ij + 3 * Tap[3];
      dT = dT * r_ij + 2 * Tap[2];
      dT += Tap[1] * rinv;
      const double2 w0 = w[0];   // D, alpha
      double e_vdW, CEvd, e_core = 0, e_lg = 0;
      if (vdw_type == 1 || vdw_type == 3) {
        const double2 w1 = w[1];   // inv_r_vdW, powgi_vdW1
        // arguments: p log r <= p log(nonb_cut); (1/p) log(r^p + g^-p) is bounded on both sides; the exp2 argument is
        // <= alpha/2.  Only the two that can run away for r -> 0 / huge alpha are clamped (from below).
        const double powr = fm::exp_b(fmax(p_vdW1 * (0.5 * fm::log_b(r2, tab)), -700.0), tab);
        const double ssum = powr + w1.y;
        const double fn13 = fm::exp_b(p_vdW1i * fm::log_b(ssum, tab), tab);
        const double exp2 = fm::exp_b(fmax(0.5 * w0.y * (1.0 - fn13 * w1.x), -700.0), tab);
        const double exp1 = exp2 * exp2;
        e_vdW = w0.x * (exp1 - 2.0 * exp2);
        const double dfn13 = (fn13 / ssum) * (powr * rinv * rinv);
        CEvd = dT * e_vdW - T * w0.x * w[2].x * (exp1 - exp2) * dfn13;
      } else {
        const double r_vdW = w[3].x;
        const double exp2 = fm::exp_c(0.5 * w0.y * (1.0 - r_ij / r_vdW), tab);
        const double exp1 = exp2 * exp2;
        e_vdW = w0.x * (exp1 - 2.0 * exp2);
        CEvd = dT * e_vdW - T * w0.x * (w0.y / r_vdW) * (exp1 - exp2) / r_ij;
      }
      if (vdw_type == 2 || vdw_type == 3) {
        const double rcore = w[3].y;
        const double2 w4 = w[4];   // ecore, acore
        e_core = w4.x * fm::exp_c(w4.y * (1.0 - (r_ij / rcore)), tab);
        const double de_core = -(w4.y / rcore) * e_core;
        CEvd += dT * e_core + T * de_core / r_ij;
        if (lgflag) {
          const double2 w5 = w[5];   // lgcij, lgre
          const double r5 = pow(r_ij, 5.0), r6 = pow(r_ij, 6.0), re6 = pow(w5.y, 6.0);
          e_lg = -(w5.x / (r6 + re6));
          const double de_lg = -6.0 * e_lg * r5 / (r6 + re6);
          CEvd += dT * e_lg + T * de_lg / r_ij;
        }
      }
      const double dr3gamij_1 = r2 * r_ij + w[2].y;
      // reference: 1/pow(x, 0.33333333333333); the cube root differs from it by < 3e-14 relative.  1/x = inv3^3.
      const double inv3 = fm::rcbrt_b(dr3gamij_1);
      const double qq = kCele * pi.w * pj.w;
      const double CEclmb = qq * (dT - T * r_ij * (inv3 * inv3 * inv3)) * inv3;
      const double ftot = CEvd + CEclmb;  // f_i = +ftot * dvec  (reference: fCdDelta[i] += -ftot*dvec, f = -fCdDelta)
      fx += ftot * dx; fy += ftot * dy; fz += ftot * dz;
      if (EV) {
        e_vdw += 0.5 * T * (e_vdW + e_core + e_lg);
        e_ele += 0.5 * qq * (T * inv3);
        const double fpair = -ftot;
        vir[0] += 0.5 * dx * dx * fpair; vir[1] += 0.5 * dy * dy * fpair; vir[2] += 0.5 * dz * dz * fpair;
        vir[3] += 0.5 * dx * dy * fpair; vir[4] += 0.5 * dx * dz * fpair; vir[5] += 0.5 * dy * dz * fpair;
      }
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) {
      // only writer of f[atom] so far in the step for a local atom would be a plain store, but bonded kernels may run
      // concurrently on another stream: keep it an atomic
      const int ia = v.row_atom[i];
      atomicAdd(&v.f[3 * ia], fx); atomicAdd(&v.f[3 * ia + 1], fy); atomicAdd(&v.f[3 * ia + 2], fz);
      if (EV) {
        // polarisation energy (reaxc_multi_body_sw64.c:100-103, plain sum) lives here because it needs this step's q
        e_pol += kKcalToEv * (P.atom[ti].chi * pi.w + (P.atom[ti].eta / 2.) * pi.w * pi.w);
        // -x_i (x) f_i^nb : cancels this kernel's share of the later f.x sum over all atoms
        vir[0] -= pi.x * fx; vir[1] -= pi.y * fy; vir[2] -= pi.z * fz;
        vir[3] -= pi.x * fy; vir[4] -= pi.x * fz; vir[5] -= pi.y * fz;
      }
    }
  }
  if (EV) {
    double vals[9] = {e_vdw, e_ele, vir[0], vir[1], vir[2], vir[3], vir[4], vir[5], e_pol};
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const double s = warp_sum(vals[k]);
      if (lane == 0) red[k * kNbWarps + wib] = s;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
      double s = 0;
      for (int q = 0; q < kNbWarps; q++) s += red[threadIdx.x * kNbWarps + q];
      if (s != 0.0) {
        if (threadIdx.x == 0) atomicAdd(&v.en[E_VDW], s);
        else if (threadIdx.x == 1) atomicAdd(&v.en[E_ELE], s);
        else if (threadIdx.x == 8) atomicAdd(&v.en[E_POL], s);
        else atomicAdd(&v.virial[threadIdx.x - 2], s);
      }
    }
  }
}

// Tabulated long-range mode (control: tabulate_long_range N): the same owner-computes full-list sweep, but the pair terms
// come from the cubic-spline records built by ForceField::lookup_tables (evaluation form of Tabulated_vdW_Coulomb_Energy,
// reaxc_nonbonded_sunway.cpp:498-519).  One or two 64-byte reads per pair from a table that stays L2-resident
// (nt^2 x (N+2) x 128 B = 20 MB for N = 10000) replace 2 log + 3 exp + cbrt; the kernel becomes L2-gather bound.
__device__ __forceinline__ double4 ldg4(const double4* p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

template <bool EV, bool PACKED>
__global__ void __launch_bounds__(kWarps * 32, 4)
k_nonbonded_tab(DevView v, DevParams P) {
  __shared__ double sh[9][kWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
  const double nonb_cut2 = P.ctl.nonb_cut * P.ctl.nonb_cut;
  const int nt = P.nt, ln = P.lut_n;
  const double dx = P.lut_dx, inv_dx = P.lut_inv_dx;
  double e_vdw = 0, e_ele = 0, e_pol = 0, vir[6] = {0, 0, 0, 0, 0, 0};
  const int stride = v.vl_stride;
  for (int i = wg; i < v.n; i += nwg) {      // i = row
    const int kself = v.rowpos[i];
    const int ti = v.type_s[kself];
    if (ti < 0) continue;
    const double4 pi = v.xqs[kself];
    const long long beg = (long long)i * stride;
    const int num = v.far_num[i];
    double fx = 0, fy = 0, fz = 0;
    for (int k0 = 0; k0 < num; k0 += 32) {
      const int k = k0 + lane;
      if (k >= num) continue;
      const int j = far_col<PACKED>(v, beg, k);
      const int tj = v.type_s[j];
      if (tj < 0) continue;
      const double4 pj = v.xqs[j];
      const double dx_ = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
      const double r2 = dist2_rn(dx_, dy, dz);
      if (!(r2 <= nonb_cut2)) continue;
      const double r_ij = sqrt(r2);
      int r = (int)(r_ij * inv_dx);
      if (r == 0) ++r;
      const double dif = r_ij - (double)(r + 1) * dx;
      const double4* rec = P.lut + ((size_t)(ti * nt + tj) * ln + r) * 4;
      const double4 cv = ldg4(rec), cc = ldg4(rec + 1);
      const double qq = pi.w * pj.w;
      const double CEvd = ((cv.w * dif + cv.z) * dif + cv.y) * dif + cv.x;
      const double CEclmb = (((cc.w * dif + cc.z) * dif + cc.y) * dif + cc.x) * qq;
      const double ftot = CEvd + CEclmb;
      fx += ftot * dx_; fy += ftot * dy; fz += ftot * dz;
      if (EV) {
        const double4 ce = ldg4(rec + 2), cl = ldg4(rec + 3);
        e_vdw += 0.5 * (((ce.w * dif + ce.z) * dif + ce.y) * dif + ce.x);
        e_ele += 0.5 * qq * (((cl.w * dif + cl.z) * dif + cl.y) * dif + cl.x);
        const double fpair = -ftot;
        vir[0] += 0.5 * dx_ * dx_ * fpair; vir[1] += 0.5 * dy * dy * fpair; vir[2] += 0.5 * dz * dz * fpair;
        vir[3] += 0.5 * dx_ * dy * fpair; vir[4] += 0.5 * dx_ * dz * fpair; vir[5] += 0.5 * dy * dz * fpair;
      }
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) {
      const int ia = v.row_atom[i];
      atomicAdd(&v.f[3 * ia], fx); atomicAdd(&v.f[3 * ia + 1], fy); atomicAdd(&v.f[3 * ia + 2], fz);
      if (EV) {
        e_pol += kKcalToEv * (P.atom[ti].chi * pi.w + (P.atom[ti].eta / 2.) * pi.w * pi.w);
        vir[0] -= pi.x * fx; vir[1] -= pi.y * fy; vir[2] -= pi.z * fz;
        vir[3] -= pi.x * fy; vir[4] -= pi.x * fz; vir[5] -= pi.y * fz;
      }
    }
  }
  if (EV) {
    double vals[9] = {e_vdw, e_ele, vir[0], vir[1], vir[2], vir[3], vir[4], vir[5], e_pol};
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const double s = warp_sum(vals[k]);
      if (lane == 0) sh[k][wib] = s;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
      double s = 0;
      for (int w = 0; w < kWarps; w++) s += sh[threadIdx.x][w];
      if (s != 0.0) {
        if (threadIdx.x == 0) atomicAdd(&v.en[E_VDW], s);
        else if (threadIdx.x == 1) atomicAdd(&v.en[E_ELE], s);
        else if (threadIdx.x == 8) atomicAdd(&v.en[E_POL], s);
        else atomicAdd(&v.virial[threadIdx.x - 2], s);
      }
    }
  }
}

}  // namespace

void launch_far_and_H(System& s, DevView& v, const DevParams& P, const double* qeq_tap, const double* shld, double swb,
                      cudaStream_t st) {
  if (v.n == 0) return;
  QeqConst qc;
  for (int t = 0; t < 8; t++) qc.Tap[t] = qeq_tap[t];
  qc.swb2 = swb * swb;
  const double far = swb > P.ctl.nonb_cut ? swb : P.ctl.nonb_cut;
  qc.far2 = far * far;
  // the inner block holds every pair that was within vl_cut_in at the build; a pair inside `far` now was within
  // far + 2 max|dx| then: usable while max|dx| <= (vl_cut_in - far) / 2 (0.1 % slack for the roundings)
  const double margin = v.vl_cut_in - far;
  qc.inner_lim2 = (v.vl_cut_in > 0.0 && margin > 0.0) ? 0.25 * 0.998 * margin * margin : -1.0;
  qc.h_quant = v.h_quant;
  BondedWork W = s.bonded_work();
  RXB_CUDA(cudaMemsetAsync(W.n_hb, 0, sizeof(int), st));
  // largest r^2 with sqrt(r^2) <= hbond_cut in round-to-nearest
  const double hc = P.ctl.hbond_cut;
  double hb2 = hc * hc;
  while (std::sqrt(std::nextafter(hb2, INFINITY)) <= hc) hb2 = std::nextafter(hb2, INFINITY);
  while (hb2 > 0.0 && std::sqrt(hb2) > hc) hb2 = std::nextafter(hb2, 0.0);
  // rows are dealt round-robin to the warps of a grid that is a whole number of waves of resident CTAs (measured:
  // 1 / 2 / 4 waves 0.997 / 0.985 / 0.977 ms; the former fixed 1184 CTAs were 1.6 waves at this register count: 1.12 ms)
  static int occ_p = 0, occ_e = 0, occ_p1 = 0, occ_e1 = 0;
  static const bool two_pass = getenv("RXB_FARH_TWOPASS") && atoi(getenv("RXB_FARH_TWOPASS")) != 0;   // the round-2a kernel, for A/B
  if (two_pass) {
    if (v.hpk) k_far_H<true><<<wave_grid(k_far_H<true>, kWarps * 32, 4, occ_p), kWarps * 32, 0, st>>>(v, P.nt, qc, shld, P.atom, hc, hb2, W);
    else k_far_H<false><<<wave_grid(k_far_H<false>, kWarps * 32, 4, occ_e), kWarps * 32, 0, st>>>(v, P.nt, qc, shld, P.atom, hc, hb2, W);
  } else {
    if (v.hpk) k_far_H1<true><<<wave_grid(k_far_H1<true>, kWarps * 32, 4, occ_p1), kWarps * 32, 0, st>>>(v, P.nt, qc, shld, P.atom, hc, hb2, W);
    else k_far_H1<false><<<wave_grid(k_far_H1<false>, kWarps * 32, 4, occ_e1), kWarps * 32, 0, st>>>(v, P.nt, qc, shld, P.atom, hc, hb2, W);
  }
  s.kernel_launches++;
}

// Throughput form of the table mode for force-only steps (4 of 5 at thermo 5): when the CEvd / CEclmb coefficient sets of the
// nt (nt + 1) / 2 unordered type pairs fit in shared memory (64 B per interval: N <= 318 knots for the 4 elements of TATB)
// every CTA stages them once and a pair costs one square root, one index computation and two cubic Horner evaluations -
// ~40 DP instructions against the 143 of the analytic kernel - with the coefficients coming from shared memory instead of
// L2.  Same records, same operations and the same summation order as k_nonbonded_tab<false>, so the forces are
// bit-identical to that kernel; what a coarse table costs in accuracy against the analytic form is measured by
// tests/test_gpu_parity.py::test_tabulated_shared_memory_mode.  (reference: the table mode of reaxc_nonbonded_sunway.cpp:
// 430-560 with LR_lookup_table from reaxc_lookup_sunway.cpp, whose builder is commented out there.)
constexpr int kTabSmemThreads = 1024;
__host__ __device__ inline int tab_pair_index(int a, int b, int nt) {   // unordered pair -> 0 .. nt (nt + 1) / 2 - 1
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return lo * nt - lo * (lo - 1) / 2 + (hi - lo);
}
template <bool PACKED>
__global__ void __launch_bounds__(kTabSmemThreads, 1)
k_nonbonded_tab_smem(DevView v, DevParams P) {
  extern __shared__ double4 s_lut[];   // [pair][interval][CEvd, CEclmb]
  const int nt = P.nt, ln = P.lut_n;
  const int npair = nt * (nt + 1) / 2;
  for (int e = threadIdx.x; e < npair * ln * 2; e += blockDim.x) {
    const int c = e & 1, r = (e >> 1) % ln, u = (e >> 1) / ln;
    int ti = 0, rem = u;
    while (rem >= nt - ti) { rem -= nt - ti; ti++; }
    s_lut[e] = P.lut[((size_t)(ti * nt + ti + rem) * ln + r) * 4 + c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
  const double nonb_cut2 = P.ctl.nonb_cut * P.ctl.nonb_cut;
  const double dx = P.lut_dx, inv_dx = P.lut_inv_dx;
  const int stride = v.vl_stride;
  for (int i = wg; i < v.n; i += nwg) {
    const int kself = v.rowpos[i];
    const int ti = v.type_s[kself];
    if (ti < 0) continue;
    const double4 pi = v.xqs[kself];
    const long long beg = (long long)i * stride;
    const int num = v.far_num[i];
    double fx = 0, fy = 0, fz = 0;
    // same software pipeline as k_nonbonded: list word of chunk t+2 and the gather of chunk t+1 in flight
    double4 p_cur = make_double4(0, 0, 0, 0);
    int t_cur = -1;
    if (lane < num) { const int j0 = far_col<PACKED>(v, beg, lane); p_cur = v.xqs[j0]; t_cur = v.type_s[j0]; }
    FarRaw<PACKED> r_nxt = 32 + lane < num ? far_raw<PACKED>(v, beg, 32 + lane) : FarRaw<PACKED>(0);
    for (int k0 = 0; k0 < num; k0 += 32) {
      const FarRaw<PACKED> r_nn = k0 + 64 + lane < num ? far_raw<PACKED>(v, beg, k0 + 64 + lane) : FarRaw<PACKED>(0);
      double4 p_nxt = make_double4(0, 0, 0, 0);
      int t_nxt = -1;
      if (k0 + 32 + lane < num) { const int jn = raw_col<PACKED>(r_nxt); p_nxt = v.xqs[jn]; t_nxt = v.type_s[jn]; }
      const int tj = t_cur;
      const double4 pj = p_cur;
      p_cur = p_nxt; t_cur = t_nxt; r_nxt = r_nn;
      if (tj < 0) continue;
      const double dx_ = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
      const double r2 = dist2_rn(dx_, dy, dz);
      if (!(r2 <= nonb_cut2)) continue;
      const double r_ij = sqrt(r2);
      int r = (int)(r_ij * inv_dx);
      if (r == 0) ++r;
      const double dif = r_ij - (double)(r + 1) * dx;
      const double4* rec = s_lut + ((size_t)tab_pair_index(ti, tj, nt) * ln + r) * 2;
      const double4 cv = rec[0], cc = rec[1];
      const double qq = pi.w * pj.w;
      const double CEvd = ((cv.w * dif + cv.z) * dif + cv.y) * dif + cv.x;
      const double CEclmb = (((cc.w * dif + cc.z) * dif + cc.y) * dif + cc.x) * qq;
      const double ftot = CEvd + CEclmb;
      fx += ftot * dx_; fy += ftot * dy; fz += ftot * dz;
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) {
      const int ia = v.row_atom[i];
      atomicAdd(&v.f[3 * ia], fx); atomicAdd(&v.f[3 * ia + 1], fy); atomicAdd(&v.f[3 * ia + 2], fz);
    }
  }
}

void launch_nonbonded(System& s, DevView& v, const DevParams& P, bool evflag, cudaStream_t st) {
  if (v.n == 0) return;
  if (P.lut) {   // Compute_NonBonded_Forces: tabulate == 0 ? analytic : tables (reaxc_forces_sunway.cpp:148-160)
    // force-only step and the force coefficient sets fit in shared memory: the throughput form (RXB_TAB_SMEM=0 keeps the
    // L2 form; read at every launch so that a test can compare the two in one process)
    const size_t tab_bytes = (size_t)(P.nt * (P.nt + 1) / 2) * P.lut_n * 2 * sizeof(double4);
    const char* ts = getenv("RXB_TAB_SMEM");
    if (!evflag && tab_bytes <= 200 * 1024 && !(ts && atoi(ts) == 0)) {
      auto ks = v.hpk ? k_nonbonded_tab_smem<true> : k_nonbonded_tab_smem<false>;
      RXB_CUDA(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_bytes));
      ks<<<148, kTabSmemThreads, tab_bytes, st>>>(v, P);
      s.kernel_launches++;
      return;
    }
    static int occ[4] = {0, 0, 0, 0};
    auto kt = v.hpk ? (evflag ? k_nonbonded_tab<true, true> : k_nonbonded_tab<false, true>)
                    : (evflag ? k_nonbonded_tab<true, false> : k_nonbonded_tab<false, false>);
    kt<<<wave_grid(kt, kWarps * 32, 2, occ[(v.hpk ? 2 : 0) + (evflag ? 1 : 0)]), kWarps * 32, 0, st>>>(v, P);
  } else {
    // math tables + block reduction scratch + 12 constants per type pair (96 B x nt^2: 4 types 1.5 KB, 40 types 154 KB)
    // A/B on one box (TATB 8x8x8): 256 x 2 (128 registers) 1.425 ms, 128 x 5 (96 registers, 8 B of spills) 1.396 ms,
    // 128 x 6 (80 registers, 56 B of spills) 1.500 ms
    static const int cfg_env = getenv("RXB_NB_CFG") ? atoi(getenv("RXB_NB_CFG")) : 1;
    const int cfg = v.hpk ? cfg_env : 0;                 // the A/B shapes exist for the packed list only
    const int threads = cfg == 0 ? 256 : 128, ctas = cfg == 0 ? 2 : (cfg == 1 ? 5 : 6);
    const size_t smem = sizeof(double) * (fm::kTabDoubles + 9 * (threads / 32) + (size_t)P.nt * P.nt * kNbPar);
    using Kern = void (*)(DevView, DevParams);
    Kern kern;
    if (cfg == 0 || !v.hpk)
      kern = v.hpk ? (evflag ? (Kern)k_nonbonded<true, true> : (Kern)k_nonbonded<false, true>)
                   : (evflag ? (Kern)k_nonbonded<true, false> : (Kern)k_nonbonded<false, false>);
    else if (cfg == 1) kern = evflag ? (Kern)k_nonbonded<true, true, 128, 5> : (Kern)k_nonbonded<false, true, 128, 5>;
    else kern = evflag ? (Kern)k_nonbonded<true, true, 128, 6> : (Kern)k_nonbonded<false, true, 128, 6>;
    if (smem > 48 * 1024) RXB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<148 * ((cfg == 0 || !v.hpk) ? 2 : ctas) * 4, (cfg == 0 || !v.hpk) ? 256 : threads, smem, st>>>(v, P);
  }
  s.kernel_launches++;
}

}  // namespace rxb
