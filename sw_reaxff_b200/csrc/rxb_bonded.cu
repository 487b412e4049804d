// Bonded energy terms and the dE/dBO chain rule, all from the device bond CSR.  A light integer kernel (K-enum, eight
// lanes per centre atom) emits dense work lists of angles / torsions / hydrogen-bond candidates; the heavy fp64 kernels run
// one thread per list item with every lane busy; K-multi runs one thread per atom (a 13-exp scalar chain), K-dbond 8 lanes
// per atom.  One wave of resident CTAs per item kernel, fp64 atomics for the scatters.
//
//   K-multi : lone pair / over- / under-coordination + bond energy + e_pol
//             /root/reference/reaxc_multi_body_sw64.c:21-333 (runs SERIALLY on the MPE in the reference)
//   K-hb    : hydrogen bonds            /root/reference/reaxc_hydrogen_bonds_sunway.cpp:310-436 (== reaxc_hydrogen_bonds_cpe.h)
//   K-vt    : valence angle + torsion + conjugation, no stored three-body list
//             /root/reference/reaxc_torsion_angles_sunway.cpp:652-1312 (== reaxc_torsion_angles_cpe.h:1-768)
//             Calculate_Theta / dCos_Theta  reaxc_valence_angles_sunway.cpp:50-84, Calculate_Omega reaxc_torsion_angles_sunway.cpp:44-125
//   K-dbond : Add_All_dBond_to_Forces_C, no-branch directed form /root/reference/reaxc_forces_sw64.c:500-587
// The reference serialises scatter conflicts with locked software write caches (SWCACHE_UPDATE); here the per-bond
// coefficient sums that every angle of a centre adds to ALL of its bonds (CEval5 / CEval6) are accumulated per centre and
// applied inside K-dbond, everything else goes through atomicAdd(double).
// Roofline: fp64 / latency bound (exp and pow per angle and torsion; no trigonometric call is left: the angles are acos()
// of clamped cosines, so their sines, cos(n omega) and sin^4(theta/2) are algebra), SURVEY.md §8d, DESIGN.md §3.
#include "rxb_system.h"

namespace rxb {
namespace {

constexpr double kConstPI = 3.14159265;  // reaxc_defs_sunway.h:55 (8 digits on purpose)
constexpr double kKcalToEv = 23.02;      // KCALpMOL_to_EV
constexpr double kHbThreshold = 1e-2;    // HB_THRESHOLD
constexpr double kMinSine = 1e-10;       // MIN_SINE
constexpr int kWarps = 8;

__device__ __forceinline__ double sqr(double a) { return a * a; }
__device__ __forceinline__ double deg2rad(double a) { return a * kConstPI / 180.0; }

__device__ __forceinline__ void fadd3(double* f, int i, double c, double x, double y, double z) {
  atomicAdd(&f[3 * i], c * x);
  atomicAdd(&f[3 * i + 1], c * y);
  atomicAdd(&f[3 * i + 2], c * z);
}

// tag-ordered half selection with z,y,x tie-break (reaxc_multi_body_sw64.c:256-268)
__device__ __forceinline__ bool half_select(int tag_i, int tag_j, const double4& xi, const double4& xj) {
  if (tag_i > tag_j) return false;
  if (tag_i == tag_j) {
    if (xj.z < xi.z) return false;
    if (xj.z == xi.z && xj.y < xi.y) return false;
    if (xj.z == xi.z && xj.y == xi.y && xj.x < xi.x) return false;
  }
  return true;
}

// commit per-thread partial energies: warp shuffle, then one atomic per block per slot
template <int K>
__device__ __forceinline__ void block_commit(double* en, const int (&slots)[K], double (&vals)[K]) {
  __shared__ double sh[K][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; k++) {
    double s = warp_sum(vals[k]);
    if (lane == 0) sh[k][w] = s;
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      double s = lane < nw ? sh[k][lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0 && s != 0.0) atomicAdd(&en[slots[k]], s);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// one thread per local atom: the per-atom scalar chain (13 exp) dominates, a warp per atom would redo it 32x
__global__ void __launch_bounds__(kWarps * 32)
k_multi(DevView v, DevParams P) {
  const int wg = blockIdx.x * blockDim.x + threadIdx.x, nwg = gridDim.x * blockDim.x;
  const double p_lp3 = P.gp[5], p_ovun3 = P.gp[32], p_ovun4 = P.gp[31], p_ovun6 = P.gp[6], p_ovun7 = P.gp[8], p_ovun8 = P.gp[9];
  const double gp3 = P.gp[3], gp4 = P.gp[4], gp7 = P.gp[7], gp10 = P.gp[10];
  const int gp37 = (int)P.gp[37];
  double e_lp = 0, e_ov = 0, e_un = 0, e_bond = 0;  // e_pol (needs the new charges) is tallied by k_nonbonded
  for (int i = wg; i < v.n; i += nwg) {
    const int ti = v.type[i];
    if (ti < 0) continue;
    const AtomPar& ai = P.atom[ti];
    const double4 xi = v.xq[i];
    const int tag_i = v.tag[i];
    const int start = v.b_start[i], cnt = v.b_cnt[i];
    const double dfvl = (ai.mass > 21.0) ? 0.0 : 1.0;
    const double Delta_i = v.Delta[i], Delta_lp_i = v.Delta_lp[i], dDelta_lp_i = v.dDelta_lp[i], Delta_lp_temp_i = v.Delta_lp_temp[i];
    const double total_bo_i = v.total_bo[i];

    double sum_ovun1 = 0, sum_ovun2 = 0, cdd_i = 0;
    for (int e = 0; e < cnt; e++) {
      const int p = start + e;
      const int j = v.b_nbr[p], tj = v.type[j];
      if (tj < 0) continue;
      const PairPar& tw = P.pair[ti * P.nt + tj];
      const double4 bo = v.b_bo[p];
      sum_ovun1 += tw.p_ovun1 * tw.De_s * bo.x;
      sum_ovun2 += (v.Delta[j] - dfvl * v.Delta_lp_temp[j]) * (bo.z + bo.w);
    }

    const double p_lp2 = ai.p_lp2, p_ovun2 = ai.p_ovun2, p_ovun5 = ai.p_ovun5;
    const double expvd2 = exp(-75 * Delta_lp_i);
    const double inv_expvd2 = 1. / (1. + expvd2);
    const double dElp = p_lp2 * inv_expvd2 + 75 * p_lp2 * Delta_lp_i * expvd2 * sqr(inv_expvd2);
    const double CElp = dElp * dDelta_lp_i;

    const double exp_ovun1 = p_ovun3 * exp(p_ovun4 * sum_ovun2);
    const double inv_exp_ovun1 = 1.0 / (1 + exp_ovun1);
    const double Delta_lpcorr = Delta_i - (dfvl * Delta_lp_temp_i) * inv_exp_ovun1;
    const double exp_ovun2 = exp(p_ovun2 * Delta_lpcorr);
    const double inv_exp_ovun2 = 1.0 / (1.0 + exp_ovun2);
    const double DlpVi = 1.0 / (Delta_lpcorr + ai.valency + 1e-8);
    const double CEover1 = Delta_lpcorr * DlpVi * inv_exp_ovun2;
    const double CEover2 = sum_ovun1 * DlpVi * inv_exp_ovun2 * (1.0 - Delta_lpcorr * (DlpVi + p_ovun2 * exp_ovun2 * inv_exp_ovun2));
    const double CEover3 = CEover2 * (1.0 - dfvl * dDelta_lp_i * inv_exp_ovun1);
    const double CEover4 = CEover2 * (dfvl * Delta_lp_temp_i) * p_ovun4 * exp_ovun1 * sqr(inv_exp_ovun1);
    const double exp_ovun2n = 1.0 / exp_ovun2;
    const double exp_ovun6 = exp(p_ovun6 * Delta_lpcorr);
    const double exp_ovun8 = p_ovun7 * exp(p_ovun8 * sum_ovun2);
    const double inv_exp_ovun2n = 1.0 / (1.0 + exp_ovun2n);
    const double inv_exp_ovun8 = 1.0 / (1.0 + exp_ovun8);
    double eun = 0.0;
    const bool active = (cnt > 0) || P.ctl.enobondsflag;
    if (active) eun = -p_ovun5 * (1.0 - exp_ovun6) * inv_exp_ovun2n * inv_exp_ovun8;
    const double CEunder1 = inv_exp_ovun2n * (p_ovun5 * p_ovun6 * exp_ovun6 * inv_exp_ovun8 + p_ovun2 * eun * exp_ovun2n);
    const double CEunder2 = -eun * p_ovun8 * exp_ovun8 * inv_exp_ovun8;
    const double CEunder3 = CEunder1 * (1.0 - dfvl * dDelta_lp_i * inv_exp_ovun1);
    const double CEunder4 = CEunder1 * (dfvl * Delta_lp_temp_i) * p_ovun4 * exp_ovun1 * sqr(inv_exp_ovun1) + CEunder2;
    {
      e_ov += sum_ovun1 * CEover1;
      cdd_i += CEover3;
      if (active) {
        e_un += eun;
        cdd_i += CEunder3;
        cdd_i += CElp;
        e_lp += p_lp2 * Delta_lp_i * inv_expvd2;
      }
    }

    const bool c2corr = p_lp3 > 0.001 && ai.is_carbon;
    for (int e = 0; e < cnt; e++) {
      const int p = start + e;
      const int j = v.b_nbr[p], tj = v.type[j];
      if (tj < 0) continue;
      const AtomPar& aj = P.atom[tj];
      const PairPar& tw = P.pair[ti * P.nt + tj];
      const double4 bo = v.b_bo[p];
      double cdbo = 0, cdbopi = 0, cdbopi2 = 0;
      if (c2corr && aj.is_carbon) {
        const double vov3 = bo.x - Delta_i - 0.040 * sqr(sqr(Delta_i));
        if (vov3 > 3.) {
          e_lp += p_lp3 * sqr(vov3 - 3.0);
          cdbo += 2. * p_lp3 * (vov3 - 3.);
          cdd_i += 2. * p_lp3 * (vov3 - 3.) * (-1. - 0.16 * (Delta_i * Delta_i * Delta_i));
        }
      }
      const double Delta_j = v.Delta[j], Dlt_j = v.Delta_lp_temp[j];
      cdbo += CEover1 * tw.p_ovun1 * tw.De_s;
      const double ftmp2 = (1.0 - dfvl * v.dDelta_lp[j]) * (bo.z + bo.w);
      const double dj = Delta_j - dfvl * Dlt_j;
      double cdd_j = CEover4 * ftmp2;
      cdbopi += CEover4 * dj; cdbopi2 += CEover4 * dj;
      cdd_j += CEunder4 * ftmp2;
      cdbopi += CEunder4 * dj; cdbopi2 += CEunder4 * dj;
      if (half_select(tag_i, v.tag[j], xi, v.xq[j])) {
        const double pow_BOs_be2 = (bo.y == 0.0) ? 0.0 : pow(bo.y, tw.p_be2);
        const double exp_be12 = exp(tw.p_be1 * (1.0 - pow_BOs_be2));
        const double CEbo = -tw.De_s * exp_be12 * (1.0 - tw.p_be1 * tw.p_be2 * pow_BOs_be2);
        e_bond += -tw.De_s * bo.y * exp_be12 - tw.De_p * bo.z - tw.De_pp * bo.w;
        cdbo += CEbo;
        cdbopi -= (CEbo + tw.De_p);
        cdbopi2 -= (CEbo + tw.De_pp);
        if (bo.x >= 1.00) {
          if (gp37 == 2 || (ai.mass == 12.0000 && aj.mass == 15.9990) || (aj.mass == 12.0000 && ai.mass == 15.9990)) {
            const double exphu = exp(-gp7 * sqr(bo.x - 2.50));
            const double exphua1 = exp(-gp3 * (total_bo_i - bo.x));
            const double exphub1 = exp(-gp3 * (v.total_bo[j] - bo.x));
            const double exphuov = exp(gp4 * (Delta_i + Delta_j));
            const double hulpov = 1.0 / (1.0 + 25.0 * exphuov);
            e_bond += gp10 * exphu * hulpov * (exphua1 + exphub1);
            cdbo += gp10 * exphu * hulpov * (exphua1 + exphub1) * (gp3 - 2.0 * gp7 * (bo.x - 2.50));
            cdd_i += -gp10 * exphu * hulpov * (gp3 * exphua1 + 25.0 * gp4 * exphuov * hulpov * (exphua1 + exphub1));
            cdd_j += -gp10 * exphu * hulpov * (gp3 * exphub1 + 25.0 * gp4 * exphuov * hulpov * (exphua1 + exphub1));
          }
        }
      }
      // own row: this lane is the only writer of slot p in this kernel
      v.b_Cdbo[p] += cdbo;
      v.b_Cdbopi[p] += cdbopi;
      v.b_Cdbopi2[p] += cdbopi2;
      if (cdd_j != 0.0) atomicAdd(&v.CdDelta[j], cdd_j);
    }
    if (cdd_i != 0.0) atomicAdd(&v.CdDelta[i], cdd_i);
  }
  const int slots[4] = {E_LP, E_OV, E_UN, E_BOND};
  double vals[4] = {e_lp, e_ov, e_un, e_bond};
  block_commit<4>(v.en, slots, vals);
}

// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void calc_theta(const double4& a, const double4& b, double& theta, double& cos_theta) {
  cos_theta = (a.y * b.y + a.z * b.z + a.w * b.w) / (a.x * b.x);
  if (cos_theta > 1.) cos_theta = 1.0;
  if (cos_theta < -1.) cos_theta = -1.0;
  theta = acos(cos_theta);
}
// geo = (d, dx, dy, dz); outputs derivative wrt end atom of a ("di"), centre ("dj"), end atom of b ("dk")
__device__ __forceinline__ void calc_dcos(const double4& a, const double4& b, double* di, double* dj, double* dk) {
  const double sqr_d_ji = a.x * a.x, sqr_d_jk = b.x * b.x;
  const double inv_dists = 1.0 / (a.x * b.x);
  const double inv_dists3 = inv_dists * inv_dists * inv_dists;
  const double dot = a.y * b.y + a.z * b.z + a.w * b.w;
  const double Cdot_inv3 = dot * inv_dists3;
  const double av[3] = {a.y, a.z, a.w}, bv[3] = {b.y, b.z, b.w};
#pragma unroll
  for (int t = 0; t < 3; t++) {
    di[t] = bv[t] * inv_dists - Cdot_inv3 * sqr_d_jk * av[t];
    dj[t] = -(bv[t] + av[t]) * inv_dists + Cdot_inv3 * (sqr_d_jk * av[t] + sqr_d_ji * bv[t]);
    dk[t] = av[t] * inv_dists - Cdot_inv3 * sqr_d_ji * bv[t];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Scalars of the dihedral i-j-k-l (reaxc_torsion_angles_sunway.cpp:95-170 Calculate_Omega).  The reference takes the two
// valence angles, calls sin/cos on them, forms omega with atan2 and then only ever uses cos(omega), cos(2 omega),
// cos(3 omega) (:1106-1130).  The angles themselves are acos() of a clamped cosine, so sin = sqrt((1-c)(1+c)) >= 0 and
// cos = c reproduce them to an ulp, and cos(atan2(s, c)) = c / hypot(s, c): no fp64 trigonometric call (and none of their
// 512-byte slow-path stack) is left.  The derivative vectors of the reference (dcos_omega_di .. dl) are linear
// combinations of the three bond vectors, r_li and the valence-angle derivatives; only the five coefficients a1..a5 and
// the scale 2/poem are returned, the torsion kernel folds them into one coefficient per (atom, base vector).
struct Omega { double cos_omega, a1, a2, a3, a4, a5, sc; };

__device__ __forceinline__ Omega calc_omega(const double4& gij, const double4& gjk, const double4& gkl, double r_li,
                                            double sin_ijk, double cos_ijk, double sin_jkl, double cos_jkl) {
  Omega o;
  const double r_ij = gij.x, r_jk = gjk.x, r_kl = gkl.x;
  const double vij[3] = {gij.y, gij.z, gij.w}, vjk[3] = {gjk.y, gjk.z, gjk.w}, vkl[3] = {gkl.y, gkl.z, gkl.w};
  auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  const double unnorm_cos_omega = -dot(vij, vjk) * dot(vjk, vkl) + sqr(r_jk) * dot(vij, vkl);
  const double cr[3] = {vjk[1] * vkl[2] - vjk[2] * vkl[1], vjk[2] * vkl[0] - vjk[0] * vkl[2], vjk[0] * vkl[1] - vjk[1] * vkl[0]};
  const double unnorm_sin_omega = -r_jk * dot(vij, cr);
  const double hyp = sqrt(unnorm_sin_omega * unnorm_sin_omega + unnorm_cos_omega * unnorm_cos_omega);
  o.cos_omega = hyp > 0 ? unnorm_cos_omega / hyp : (signbit(unnorm_cos_omega) ? -1.0 : 1.0);   // atan2(0, +-0) = 0, pi
  const double htra = r_ij + cos_ijk * (r_kl * cos_jkl - r_jk);
  const double htrb = r_jk - r_ij * cos_ijk - r_kl * cos_jkl;
  const double htrc = r_kl + cos_jkl * (r_ij * cos_ijk - r_jk);
  const double hthd = r_ij * sin_ijk * (r_jk - r_kl * cos_jkl);
  const double hthe = r_kl * sin_jkl * (r_jk - r_ij * cos_ijk);
  const double hnra = r_kl * sin_ijk * sin_jkl;
  const double hnrc = r_ij * sin_ijk * sin_jkl;
  const double hnhd = r_ij * r_kl * cos_ijk * sin_jkl;
  const double hnhe = r_ij * r_kl * sin_ijk * cos_jkl;
  double poem = 2.0 * r_ij * r_kl * sin_ijk * sin_jkl;
  if (poem < 1e-20) poem = 1e-20;
  const double tel = sqr(r_ij) + sqr(r_jk) + sqr(r_kl) - sqr(r_li) -
                     2.0 * (r_ij * r_jk * cos_ijk - r_ij * r_kl * cos_ijk * cos_jkl + r_jk * r_kl * cos_jkl);
  double arg = tel / poem;
  if (arg > 1.0) arg = 1.0;
  if (arg < -1.0) arg = -1.0;
  if (sin_ijk >= 0 && sin_ijk <= kMinSine) sin_ijk = kMinSine;
  else if (sin_ijk <= 0 && sin_ijk >= -kMinSine) sin_ijk = -kMinSine;
  if (sin_jkl >= 0 && sin_jkl <= kMinSine) sin_jkl = kMinSine;
  else if (sin_jkl <= 0 && sin_jkl >= -kMinSine) sin_jkl = -kMinSine;
  o.a1 = (htra - arg * hnra) / r_ij;
  o.a2 = (hthd - arg * hnhd) / sin_ijk;
  o.a3 = (hthe - arg * hnhe) / sin_jkl;
  o.a4 = (htrc - arg * hnrc) / r_kl;
  o.a5 = htrb / r_jk;
  o.sc = 2.0 / poem;
  return o;
}

// cos(theta) of the angle between two bond vectors of one centre, clamped like calc_theta, plus the three coefficients
// that express the derivative vectors of calc_dcos in the bond vectors themselves:
//   d/d(end of a) = P b - Q a,   d/d(end of b) = P a - R b,   d/d(centre) = -(both)
struct DCos { double c, P, Q, R; };
__device__ __forceinline__ DCos calc_dcos_coef(const double4& a, const double4& b) {
  DCos o;
  const double dot = a.y * b.y + a.z * b.z + a.w * b.w;
  o.P = 1.0 / (a.x * b.x);
  o.c = fmin(1.0, fmax(-1.0, dot / (a.x * b.x)));
  const double Cdot_inv3 = dot * (o.P * o.P * o.P);
  o.Q = Cdot_inv3 * (b.x * b.x);
  o.R = Cdot_inv3 * (a.x * a.x);
  return o;
}

// ------------------------------------------------------------------------------------------------------------
// K-enum: light integer pass, one warp per local centre atom.  Emits dense work lists so that the heavy fp64 kernels
// run one thread per angle / torsion / hydrogen bond with every lane busy (the reference walks 4-deep nested loops per
// centre atom on one CPE, reaxc_torsion_angles_cpe.h:192-760; on a GPU that shape leaves most lanes idle).
//   angle item   (j, pk, ph)        : strong bonds pk < ph of j with BO_jk*BO_hj > thb_cutsq      (valence filter :796-801)
//   torsion item (j, pk, ph, pw)    : bond j-k selected once by tag order, strong ph on j, strong pw on k, i != l,
//                                     a torsion parameter set exists, BO_hj*BO_jk*BO_kw > thb_cut   (:982-1066)
//   hbond item   (j, pi, k)         : H atom j, acceptor bond pi with BO >= 0.01, acceptor-type k within hbond_cut,
//                                     tag_i != tag_k, r0_hb > 0                                    (hydrogen_bonds :313-354)
// "strong" = BO > thb_cut; every filter of the reference needs it for both bonds of an angle and all three of a torsion.
// Also stores the per-centre SBO quantities the angle items need (SBO2, CSBO2, dSBO1, dSBO2; :745-787).
// Eight lanes per centre, four centres per warp: a bond row has ~7.5 entries and a centre 0 - 16 ordered strong pairs, so a
// warp per centre left most lanes idle (and lane 0 alone did the pow() calls of the SBO block for the whole warp).
__global__ void __launch_bounds__(kWarps * 32)
k_enum(DevView v, DevParams P, BondedWork W) {
  extern __shared__ int s_strong[];   // [kWarps][4][strong_cap]: the strong bonds of each centre (grown on overflow bit 8)
  const int strong_cap = v.strong_cap;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, sub = lane & 7, grp = lane >> 3;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
  const double thb_cut = P.ctl.thb_cut, thb_cutsq = P.ctl.thb_cutsq;
  const double p_val8 = P.gp[33], p_val9 = P.gp[16];
  const int nt = P.nt;
  const unsigned lt_mask = (1u << lane) - 1, lt_sub = (1u << sub) - 1;
  int* strong_list = s_strong + (size_t)(wib * 4 + grp) * strong_cap;
  for (int j0 = 4 * wg; j0 < v.n; j0 += 4 * nwg) {   // warp-uniform trip count: the ballots below need all 32 lanes
    const int j = j0 + grp;
    const int type_j = j < v.n ? v.type[j] : -1;
    const int start_j = type_j >= 0 ? v.b_start[j] : 0, cnt_j = type_j >= 0 ? max(v.b_cnt[j], 0) : 0;
    // ---- strong list, SBO sums ----
    double SBOp = 0, prod_SBO = 1;
    int ns = 0;
    const int max_cnt = __reduce_max_sync(0xffffffffu, cnt_j);
    for (int e0 = 0; e0 < max_cnt; e0 += 8) {
      const int e = e0 + sub;
      bool strong = false;
      if (e < cnt_j) {
        const double4 bo = v.b_bo[start_j + e];
        SBOp += bo.z + bo.w;
        double t8 = bo.x * bo.x; t8 *= t8; t8 *= t8;
        prod_SBO *= exp(-t8);
        strong = bo.x > thb_cut;
      }
      const unsigned m = (__ballot_sync(0xffffffffu, strong) >> (8 * grp)) & 0xffu;   // this centre's 8 lanes
      if (strong) { const int slot = ns + __popc(m & lt_sub); if (slot < strong_cap) strong_list[slot] = start_j + e; }
      ns += __popc(m);
    }
    __syncwarp();
    if (ns > strong_cap) { if (sub == 0) { atomicOr(v.overflow, 8); atomicMax(v.need_row + 1, ns); } ns = strong_cap; }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      SBOp += __shfl_xor_sync(0xffffffffu, SBOp, o, 8);
      prod_SBO *= __shfl_xor_sync(0xffffffffu, prod_SBO, o, 8);
    }
    if (sub == 0 && cnt_j > 0) {
      const double Delta_boc_j = v.Delta_boc[j];
      double vlpadj, dSBO2;
      if (v.vlpex[j] >= 0) { vlpadj = 0; dSBO2 = prod_SBO - 1; }
      else { vlpadj = v.nlp[j]; dSBO2 = (prod_SBO - 1) * (1 - p_val8 * v.dDelta_lp[j]); }
      const double SBO = SBOp + (1 - prod_SBO) * (-Delta_boc_j - p_val8 * vlpadj);
      const double dSBO1 = -8 * prod_SBO * (Delta_boc_j + p_val8 * vlpadj);
      double SBO2, CSBO2;
      if (SBO <= 0) { SBO2 = 0; CSBO2 = 0; }
      else if (SBO < 2) {   // s^(p-1) = s^p / s, s in (0, 1]: one pow() for both
        const double sb = SBO <= 1 ? SBO : 2 - SBO, pw = pow(sb, p_val9);
        SBO2 = SBO <= 1 ? pw : 2 - pw;
        CSBO2 = p_val9 * (pw / sb);
      }
      else { SBO2 = 2; CSBO2 = 0; }
      W.sbo[j] = make_double4(SBO2, CSBO2, dSBO1, dSBO2);
    }

    // ---- angles and torsions from ordered strong pairs (a, b), a != b ----
    const int npair = ns >= 2 ? ns * ns : 0;
    const int max_np = __reduce_max_sync(0xffffffffu, npair);
    double4 xj = make_double4(0, 0, 0, 0);
    int tag_j = 0;
    if (npair > 0) { xj = v.xq[j]; tag_j = v.tag[j]; }
    for (int q0 = 0; q0 < max_np; q0 += 8) {
      const int q = q0 + sub;
      bool ang = false;
      int pk = -1, ph = -1, start_k = 0;
      // torsion scan of k's row (done below, 64 entries per round): what this lane needs for it
      int t_rounds = 0, t_cnt_k = 0, t_pj = -1, t_h = -1, t_base = 0;
      double t_bo2 = 0.0;   // bo_hj * bo_jk
      if (q < npair) {
        const int a = q / ns, b = q - a * ns;
        if (a != b) {
          pk = strong_list[a]; ph = strong_list[b];
          const int k = v.b_nbr[pk], h = v.b_nbr[ph];
          const int type_k = v.type[k], type_h = v.type[h];
          const int cnt_k = v.b_cnt[k];
          start_k = v.b_start[k];
          if (type_k >= 0 && type_h >= 0 && cnt_k > 0) {
            const double bo_jk = v.b_bo[pk].x, bo_hj = v.b_bo[ph].x;
            ang = (ph > pk) && (bo_jk * bo_hj > thb_cutsq);
            const int pj = v.b_sym[pk];
            if (pj >= 0 && half_select(tag_j, v.tag[k], xj, v.xq[k])) {
              t_rounds = (cnt_k + 63) >> 6; t_cnt_k = cnt_k; t_pj = pj; t_h = h; t_bo2 = bo_hj * bo_jk;
              t_base = ((type_h * nt + type_j) * nt + type_k) * nt;
            }
          }
        }
      }
      // angles: one atomic for the four centres of the warp
      const unsigned m = __ballot_sync(0xffffffffu, ang);
      if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(W.n_ang, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (ang) {
          const int o = base + __popc(m & lt_mask);
          if (o < W.cap_ang) W.ang[o] = make_int4(j, pk, ph, 0);
        }
      }
      // torsions: k's row is scanned 64 entries per round (a 64-bit match mask per lane; rows of more than 64 bonds
      // simply take more rounds), then a warp exclusive scan of the per-lane counts carves the output slots
      const int rounds = __reduce_max_sync(0xffffffffu, t_rounds);
      for (int c = 0; c < rounds; c++) {
        unsigned long long tmask = 0ull;  // matching pw offsets in this 64-entry window of k's row
        const int e0 = c << 6;
        if (c < t_rounds) {
          const int ne = min(t_cnt_k - e0, 64);
          for (int e = 0; e < ne; e++) {
            const int pw = start_k + e0 + e;
            if (pw == t_pj) continue;
            const double bo_kl = v.b_bo[pw].x;
            if (!(bo_kl > thb_cut)) continue;
            const int l = v.b_nbr[pw];
            if (l == t_h) continue;
            const int type_l = v.type[l];
            if (type_l < 0) continue;
            if (!P.tors[t_base + type_l].cnt) continue;
            if (!(t_bo2 * bo_kl > thb_cut)) continue;
            tmask |= 1ull << e;
          }
        }
        const int mine = __popcll(tmask);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
          int base = 0;
          if (lane == 0) base = atomicAdd(W.n_tor, total);
          base = __shfl_sync(0xffffffffu, base, 0);
          int o = base + incl - mine;
          while (tmask) {
            const int e = __ffsll((long long)tmask) - 1;
            tmask &= tmask - 1;
            if (o < W.cap_tor) W.tor[o] = make_int4(j, pk, ph, start_k + e0 + e);
            o++;
          }
        }
      }
    }
    __syncwarp();
  }
}

constexpr int kItemThreads = 128;

// ------------------------------------------------------------------------------------------------------------
// K-hb: one thread per (H atom j, partner k) candidate emitted by K-farH; loops over the acceptor bonds of j
// (p_hbond == 2, BO >= HB_THRESHOLD; usually exactly one)   reaxc_hydrogen_bonds_sunway.cpp:313-436
__global__ void __launch_bounds__(kItemThreads)
k_hbond_items(DevView v, DevParams P, BondedWork W) {
  const int nitems = min(*W.n_hb, W.cap_hb);
  const int nt = P.nt;
  double e_hb = 0;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < nitems; it += gridDim.x * blockDim.x) {
    const int4 w = W.hb[it];
    const int j = w.x, k = w.y;
    const int start = v.b_start[j], cnt = v.b_cnt[j];
    const int tj = v.type[j], tk = v.type[k], tag_k = v.tag[k];
    const double4 xj = v.xq[j], xk = v.xq[k];
    const double dx = xk.x - xj.x, dy = xk.y - xj.y, dz = xk.z - xj.z;
    const double r_jk = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    const double4 gjk = make_double4(r_jk, dx, dy, dz);
    double fjx = 0, fjy = 0, fjz = 0, fkx = 0, fky = 0, fkz = 0;
    for (int pi = start; pi < start + cnt; pi++) {
      const double BOij = v.b_bo[pi].x;
      if (!(BOij >= kHbThreshold)) continue;
      const int i = v.b_nbr[pi];
      const int ti = v.type[i];
      if (ti < 0 || P.atom[ti].p_hbond != 2) continue;
      if (v.tag[i] == tag_k) continue;
      const HbPar hp = P.hb[(ti * nt + tj) * nt + tk];
      if (hp.r0_hb <= 0.0) continue;
      const double4 gij = v.b_geo[pi];
      double theta, cos_theta, di[3], dj[3], dk[3];
      calc_theta(gij, gjk, theta, cos_theta);
      calc_dcos(gij, gjk, di, dj, dk);
      const double cos_xhz1 = (1.0 - cos_theta);
      const double sin_xhz4 = sqr(0.5 * cos_xhz1);   // sin^4(theta/2) = ((1 - cos theta)/2)^2: no sin(), no acos()
      const double exp_hb2 = exp(-hp.p_hb2 * BOij);
      const double exp_hb3 = exp(-hp.p_hb3 * (hp.r0_hb / r_jk + r_jk / hp.r0_hb - 2.0));
      const double ehb = hp.p_hb1 * (1.0 - exp_hb2) * exp_hb3 * sin_xhz4;
      e_hb += ehb;
      const double CEhb1 = hp.p_hb1 * hp.p_hb2 * exp_hb2 * exp_hb3 * sin_xhz4;
      const double CEhb2 = -hp.p_hb1 / 2.0 * (1.0 - exp_hb2) * exp_hb3 * cos_xhz1;
      const double CEhb3 = -hp.p_hb3 * (-hp.r0_hb / sqr(r_jk) + 1.0 / hp.r0_hb) * ehb;
      const double c3 = CEhb3 / r_jk;
      atomicAdd(&v.b_Cdbo[pi], CEhb1);
      // reference accumulates -force in fCdDelta; f is the true force here
      fadd3(v.f, i, -CEhb2, di[0], di[1], di[2]);
      fjx -= CEhb2 * dj[0] - c3 * dx; fjy -= CEhb2 * dj[1] - c3 * dy; fjz -= CEhb2 * dj[2] - c3 * dz;
      fkx -= CEhb2 * dk[0] + c3 * dx; fky -= CEhb2 * dk[1] + c3 * dy; fkz -= CEhb2 * dk[2] + c3 * dz;
    }
    if (fjx != 0.0 || fjy != 0.0 || fjz != 0.0) {
      atomicAdd(&v.f[3 * j], fjx); atomicAdd(&v.f[3 * j + 1], fjy); atomicAdd(&v.f[3 * j + 2], fjz);
      atomicAdd(&v.f[3 * k], fkx); atomicAdd(&v.f[3 * k + 1], fky); atomicAdd(&v.f[3 * k + 2], fkz);
    }
  }
  const int slots[1] = {E_HB};
  double vals[1] = {e_hb};
  block_commit<1>(v.en, slots, vals);
}

// ------------------------------------------------------------------------------------------------------------
// K-angle: one thread per valence angle k-j-h   reaxc_torsion_angles_sunway.cpp:803-979
template <int MINB>
__global__ void __launch_bounds__(kItemThreads, MINB)
k_angle_items(DevView v, DevParams P, BondedWork W) {
  const int nitems = min(*W.n_ang, W.cap_ang);
  const double p_val6 = P.gp[14], p_val10 = P.gp[17];
  const double p_pen2 = P.gp[19], p_pen3 = P.gp[20], p_pen4 = P.gp[21];
  const double p_coa2 = P.gp[2], p_coa3 = P.gp[38], p_coa4 = P.gp[30];
  const double thb_cut = P.ctl.thb_cut;
  const int nt = P.nt;
  double e_ang = 0, e_pen = 0, e_coa = 0;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < nitems; it += gridDim.x * blockDim.x) {
    const int4 w = W.ang[it];
    const int j = w.x, pk = w.y, ph = w.z;
    const int k = v.b_nbr[pk], h = v.b_nbr[ph];
    const int type_j = v.type[j], type_k = v.type[k], type_h = v.type[h];
    const double4 gjk = v.b_geo[pk], ghj = v.b_geo[ph];
    const double BOA_jk = v.b_bo[pk].x - thb_cut, BOA_hj = v.b_bo[ph].x - thb_cut;
    const double4 sb = W.sbo[j];
    const double SBO2 = sb.x, CSBO2 = sb.y, dSBO1 = sb.z, dSBO2 = sb.w;
    const double Delta_boc_j = v.Delta_boc[j], Delta_j = v.Delta[j], Delta_val_j = v.Delta_val[j];
    const double p_val3 = P.atom[type_j].p_val3, p_val5 = P.atom[type_j].p_val5;
    const double expval6 = exp(p_val6 * Delta_boc_j);
    double theta_hjk, cos_theta_hjk;
    calc_theta(gjk, ghj, theta_hjk, cos_theta_hjk);
    double sin_theta_hjk = sqrt((1.0 - cos_theta_hjk) * (1.0 + cos_theta_hjk));   // theta = acos(c) in [0, pi]
    if (sin_theta_hjk < 1.0e-5) sin_theta_hjk = 1.0e-5;
    const AngleSet& as = P.angle[(type_k * nt + type_j) * nt + type_h];
    const double tbo_k = v.total_bo[k], tbo_h = v.total_bo[h];
    double cdbo_k = 0, cdbo_h = 0, cdd_j = 0, cdd_k = 0, cdd_h = 0, s5 = 0, s6 = 0, c8 = 0;
    for (int c = 0; c < as.cnt && c < kMaxAngleSets; c++) {
      const AnglePar tp = as.prm[c];
      if (!(fabs(tp.p_val1) > 0.001)) continue;
      const double p_val1 = tp.p_val1, p_val2 = tp.p_val2, p_val4 = tp.p_val4, p_val7 = tp.p_val7, theta_00 = tp.theta_00;
      // BOA > 0 (both bonds are "strong"), so BOA^(p-1) = BOA^p / BOA: two pow() calls instead of four
      const double pow_jk = pow(BOA_jk, p_val4), pow_hj = pow(BOA_hj, p_val4);
      const double exp3jk = exp(-p_val3 * pow_jk);
      const double f7_jk = 1.0 - exp3jk;
      const double Cf7jk = p_val3 * p_val4 * (pow_jk / BOA_jk) * exp3jk;
      const double exp3hj = exp(-p_val3 * pow_hj);
      const double f7_hj = 1.0 - exp3hj;
      const double Cf7hj = p_val3 * p_val4 * (pow_hj / BOA_hj) * exp3hj;
      const double expval7 = exp(-p_val7 * Delta_boc_j);
      const double trm8 = 1.0 + expval6 + expval7;
      const double f8_Dj = p_val5 - ((p_val5 - 1.0) * (2.0 + expval6) / trm8);
      const double Cf8j = ((1.0 - p_val5) / sqr(trm8)) *
                          (p_val6 * expval6 * trm8 - (2.0 + expval6) * (p_val6 * expval6 - p_val7 * expval7));
      const double ex10 = exp(-p_val10 * (2.0 - SBO2));
      double theta_0 = 180.0 - theta_00 * (1.0 - ex10);
      theta_0 = deg2rad(theta_0);
      const double expval2theta = exp(-p_val2 * sqr(theta_0 - theta_hjk));
      const double expval12theta = (p_val1 >= 0) ? p_val1 * (1.0 - expval2theta) : p_val1 * -expval2theta;
      const double CEval1 = Cf7jk * f7_hj * f8_Dj * expval12theta;
      const double CEval2 = Cf7hj * f7_jk * f8_Dj * expval12theta;
      const double CEval3 = Cf8j * f7_hj * f7_jk * expval12theta;
      const double CEval4 = -2.0 * p_val1 * p_val2 * f7_jk * f7_hj * f8_Dj * expval2theta * (theta_0 - theta_hjk);
      const double Ctheta_0 = p_val10 * deg2rad(theta_00) * ex10;
      const double CEval5 = -CEval4 * Ctheta_0 * CSBO2;
      const double CEval6 = CEval5 * dSBO1;
      const double CEval7 = CEval5 * dSBO2;
      const double CEval8 = -CEval4 / sin_theta_hjk;
      e_ang += f7_jk * f7_hj * f8_Dj * expval12theta;

      const double exp_pen2jk = exp(-p_pen2 * sqr(BOA_jk - 2.0));
      const double exp_pen2hj = exp(-p_pen2 * sqr(BOA_hj - 2.0));
      const double exp_pen3 = exp(-p_pen3 * Delta_j);
      const double exp_pen4 = exp(p_pen4 * Delta_j);
      const double trm_pen34 = 1.0 + exp_pen3 + exp_pen4;
      const double f9_Dj = (2.0 + exp_pen3) / trm_pen34;
      const double Cf9j = (-p_pen3 * exp_pen3 * trm_pen34 - (2.0 + exp_pen3) * (-p_pen3 * exp_pen3 + p_pen4 * exp_pen4)) / sqr(trm_pen34);
      const double epen = tp.p_pen1 * f9_Dj * exp_pen2jk * exp_pen2hj;
      e_pen += epen;
      const double CEpen1 = epen * Cf9j / f9_Dj;
      const double tpen = -2.0 * p_pen2 * epen;
      const double CEpen2 = tpen * (BOA_jk - 2.0);
      const double CEpen3 = tpen * (BOA_hj - 2.0);

      const double exp_coa2 = exp(p_coa2 * Delta_val_j);
      const double ecoa = tp.p_coa1 / (1. + exp_coa2) * exp(-p_coa3 * sqr(tbo_k - BOA_jk)) * exp(-p_coa3 * sqr(tbo_h - BOA_hj)) *
                          exp(-p_coa4 * sqr(BOA_jk - 1.5)) * exp(-p_coa4 * sqr(BOA_hj - 1.5));
      e_coa += ecoa;
      const double CEcoa1 = -2 * p_coa4 * (BOA_jk - 1.5) * ecoa;
      const double CEcoa2 = -2 * p_coa4 * (BOA_hj - 1.5) * ecoa;
      const double CEcoa3 = -p_coa2 * exp_coa2 * ecoa / (1 + exp_coa2);
      const double CEcoa4 = -2 * p_coa3 * (tbo_k - BOA_jk) * ecoa;
      const double CEcoa5 = -2 * p_coa3 * (tbo_h - BOA_hj) * ecoa;

      cdbo_k += (CEval1 + CEpen2 + (CEcoa1 - CEcoa4));
      cdbo_h += (CEval2 + CEpen3 + (CEcoa2 - CEcoa5));
      cdd_j += ((CEval3 + CEval7) + CEpen1 + CEcoa3);
      cdd_k += CEcoa4;
      cdd_h += CEcoa5;
      s6 += CEval6;  // every bond t of j: Cdbo[t] += CEval6 * BO_t^7          (applied in K-dbond through W.sum56)
      s5 += CEval5;  //                    Cdbopi[t], Cdbopi2[t] += CEval5
      c8 += CEval8;
    }
    if (cdbo_k != 0.0) atomicAdd(&v.b_Cdbo[pk], cdbo_k);
    if (cdbo_h != 0.0) atomicAdd(&v.b_Cdbo[ph], cdbo_h);
    if (cdd_j != 0.0) atomicAdd(&v.CdDelta[j], cdd_j);
    if (cdd_k != 0.0) atomicAdd(&v.CdDelta[k], cdd_k);
    if (cdd_h != 0.0) atomicAdd(&v.CdDelta[h], cdd_h);
    if (s5 != 0.0) atomicAdd(&W.sum56[j].x, s5);
    if (s6 != 0.0) atomicAdd(&W.sum56[j].y, s6);
    if (c8 != 0.0) {
      // d cos(theta) / d(k, j, h) as coefficients on the two bond vectors (see calc_dcos_coef), formed only now from the
      // re-read geometry (an L1 hit) so that nine doubles do not stay live across the parameter-set loop above
      const double4 a = v.b_geo[pk], b = v.b_geo[ph];
      const DCos dc = calc_dcos_coef(a, b);
      const double kA = -c8 * -dc.Q, kB = -c8 * dc.P, hA = -c8 * dc.P, hB = -c8 * -dc.R;
      const double fk[3] = {kA * a.y + kB * b.y, kA * a.z + kB * b.z, kA * a.w + kB * b.w};
      const double fh[3] = {hA * a.y + hB * b.y, hA * a.z + hB * b.z, hA * a.w + hB * b.w};
      fadd3(v.f, k, 1.0, fk[0], fk[1], fk[2]);
      fadd3(v.f, h, 1.0, fh[0], fh[1], fh[2]);
      fadd3(v.f, j, -1.0, fk[0] + fh[0], fk[1] + fh[1], fk[2] + fh[2]);
    }
  }
  const int slots[3] = {E_ANG, E_PEN, E_COA};
  double vals[3] = {e_ang, e_pen, e_coa};
  block_commit<3>(v.en, slots, vals);
}

// ------------------------------------------------------------------------------------------------------------
// K-tors: one thread per torsion h-j-k-l (= i-j-k-l)   reaxc_torsion_angles_sunway.cpp:994-1290
template <int MINB>
__global__ void __launch_bounds__(kItemThreads, MINB)
k_torsion_items(DevView v, DevParams P, BondedWork W) {
  const int nitems = min(*W.n_tor, W.cap_tor);
  const double p_tor2 = P.gp[23], p_tor3 = P.gp[24], p_tor4 = P.gp[25], p_cot2 = P.gp[27];
  const double thb_cut = P.ctl.thb_cut;
  const int nt = P.nt;
  double e_tor = 0, e_con = 0;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < nitems; it += gridDim.x * blockDim.x) {
    const int4 w = W.tor[it];
    const int j = w.x, pk = w.y, ph = w.z, pw = w.w;
    const int k = v.b_nbr[pk], i = v.b_nbr[ph], l = v.b_nbr[pw];
    const int pj = v.b_sym[pk];
    const double4 gjk = v.b_geo[pk], ghj = v.b_geo[ph], gkj = v.b_geo[pj], gkl = v.b_geo[pw];
    const double4 bo_jk = v.b_bo[pk];
    const double bo_hj = v.b_bo[ph].x, bo_kl = v.b_bo[pw].x;
    const double BOA_jk = bo_jk.x - thb_cut, BOA_ij = bo_hj - thb_cut, BOA_kl = bo_kl - thb_cut;
    const TorsPar fp = P.tors[((v.type[i] * nt + v.type[j]) * nt + v.type[k]) * nt + v.type[l]];
    const DCos hjk = calc_dcos_coef(gjk, ghj);   // a = j->k, b = j->i
    const DCos jkl = calc_dcos_coef(gkj, gkl);   // a = k->j, b = k->l
    const double cos_theta_hjk = hjk.c, cos_theta_jkl = jkl.c;
    const double cos_ijk = cos_theta_hjk, sin_ijk = sqrt((1.0 - cos_ijk) * (1.0 + cos_ijk));
    double tan_ijk_i;
    if (sin_ijk >= 0 && sin_ijk <= kMinSine) tan_ijk_i = cos_ijk / kMinSine;
    else if (sin_ijk <= 0 && sin_ijk >= -kMinSine) tan_ijk_i = cos_ijk / -kMinSine;
    else tan_ijk_i = cos_ijk / sin_ijk;
    const double cos_jkl = cos_theta_jkl, sin_jkl = sqrt((1.0 - cos_jkl) * (1.0 + cos_jkl));
    double tan_jkl_i;
    if (sin_jkl >= 0 && sin_jkl <= kMinSine) tan_jkl_i = cos_jkl / kMinSine;
    else if (sin_jkl <= 0 && sin_jkl >= -kMinSine) tan_jkl_i = cos_jkl / -kMinSine;
    else tan_jkl_i = cos_jkl / sin_jkl;
    const double exp_tor2_ij = exp(-p_tor2 * BOA_ij);
    const double exp_cot2_ij = exp(-p_cot2 * sqr(BOA_ij - 1.5));
    const double exp_tor2_jk = exp(-p_tor2 * BOA_jk);
    const double exp_cot2_jk = exp(-p_cot2 * sqr(BOA_jk - 1.5));
    const double exp_tor2_kl = exp(-p_tor2 * BOA_kl);
    const double exp_cot2_kl = exp(-p_cot2 * sqr(BOA_kl - 1.5));
    const double DjDk = v.Delta_boc[j] + v.Delta_boc[k];
    const double exp_tor3_DjDk = exp(-p_tor3 * DjDk);
    const double exp_tor4_DjDk = exp(p_tor4 * DjDk);
    const double exp_tor34_inv = 1.0 / (1.0 + exp_tor3_DjDk + exp_tor4_DjDk);
    const double f11_DjDk = (2.0 + exp_tor3_DjDk) * exp_tor34_inv;
    const double4 xi = v.xq[i], xl = v.xq[l];
    const double dvec_li[3] = {xi.x - xl.x, xi.y - xl.y, xi.z - xl.z};
    const double r_li = sqrt(dvec_li[0] * dvec_li[0] + dvec_li[1] * dvec_li[1] + dvec_li[2] * dvec_li[2]);
    const Omega om = calc_omega(ghj, gjk, gkl, r_li, sin_ijk, cos_ijk, sin_jkl, cos_jkl);
    const double cos_omega = om.cos_omega, cos2omega = 2.0 * sqr(cos_omega) - 1.0;
    const double cos3omega = cos_omega * (4.0 * sqr(cos_omega) - 3.0);
    const double exp_tor1 = exp(fp.p_tor1 * sqr(2.0 - bo_jk.z - f11_DjDk));
    const double fn10 = (1.0 - exp_tor2_ij) * (1.0 - exp_tor2_jk) * (1.0 - exp_tor2_kl);
    const double CV = 0.5 * (fp.V1 * (1.0 + cos_omega) + fp.V2 * exp_tor1 * (1.0 - cos2omega) + fp.V3 * (1.0 + cos3omega));
    e_tor += fn10 * sin_ijk * sin_jkl * CV;
    const double dfn11 = (-p_tor3 * exp_tor3_DjDk + (p_tor3 * exp_tor3_DjDk - p_tor4 * exp_tor4_DjDk) * (2.0 + exp_tor3_DjDk) * exp_tor34_inv) * exp_tor34_inv;
    const double CEtors1 = sin_ijk * sin_jkl * CV;
    const double CEtors2 = -fn10 * 2.0 * fp.p_tor1 * fp.V2 * exp_tor1 * (2.0 - bo_jk.z - f11_DjDk) * (1.0 - sqr(cos_omega)) * sin_ijk * sin_jkl;
    const double CEtors3 = CEtors2 * dfn11;
    const double CEtors4 = CEtors1 * p_tor2 * exp_tor2_ij * (1.0 - exp_tor2_jk) * (1.0 - exp_tor2_kl);
    const double CEtors5 = CEtors1 * p_tor2 * (1.0 - exp_tor2_ij) * exp_tor2_jk * (1.0 - exp_tor2_kl);
    const double CEtors6 = CEtors1 * p_tor2 * (1.0 - exp_tor2_ij) * (1.0 - exp_tor2_jk) * exp_tor2_kl;
    const double cmn = -fn10 * CV;
    const double CEtors7 = cmn * sin_jkl * tan_ijk_i;
    const double CEtors8 = cmn * sin_ijk * tan_jkl_i;
    const double CEtors9 = fn10 * sin_ijk * sin_jkl * (0.5 * fp.V1 - 2.0 * fp.V2 * exp_tor1 * cos_omega + 1.5 * fp.V3 * (cos2omega + 2.0 * sqr(cos_omega)));
    const double fn12 = exp_cot2_ij * exp_cot2_jk * exp_cot2_kl;
    const double cterm = (1.0 + (sqr(cos_omega) - 1.0) * sin_ijk * sin_jkl);
    e_con += fp.p_cot1 * fn12 * cterm;
    const double Cconj = -2.0 * fn12 * fp.p_cot1 * p_cot2 * cterm;
    const double CEconj1 = Cconj * (BOA_ij - 1.5e0);
    const double CEconj2 = Cconj * (BOA_jk - 1.5e0);
    const double CEconj3 = Cconj * (BOA_kl - 1.5e0);
    const double CEconj4 = -fp.p_cot1 * fn12 * (sqr(cos_omega) - 1.0) * sin_jkl * tan_ijk_i;
    const double CEconj5 = -fp.p_cot1 * fn12 * (sqr(cos_omega) - 1.0) * sin_ijk * tan_jkl_i;
    const double CEconj6 = 2.0 * fp.p_cot1 * fn12 * cos_omega * sin_ijk * sin_jkl;

    atomicAdd(&v.b_Cdbopi[pk], CEtors2);
    atomicAdd(&v.CdDelta[j], CEtors3);
    atomicAdd(&v.CdDelta[k], CEtors3);
    atomicAdd(&v.b_Cdbo[ph], (CEtors4 + CEconj1));
    atomicAdd(&v.b_Cdbo[pk], (CEtors5 + CEconj2));
    atomicAdd(&v.b_Cdbo[pw], (CEtors6 + CEconj3));
    const double c74 = CEtors7 + CEconj4, c85 = CEtors8 + CEconj5, c96 = CEtors9 + CEconj6;
    // -force on i, j, k, l = c74 dcos(hjk) + c85 dcos(jkl) + c96 dcos(omega) (:1190-1290), every vector of which is a
    // combination of A = j->k, B = j->i, C = k->j, D = k->l and r_li: one coefficient per (atom, base vector)
    const double s96 = c96 * om.sc, uc = c74 - s96 * om.a2, wc = c85 - s96 * om.a3;
    const double iA = uc * hjk.P, iB = s96 * om.a1 - uc * hjk.R;
    const double lC = wc * jkl.P, lD = s96 * om.a4 - wc * jkl.R;
    const double jA = uc * (hjk.Q - hjk.P) - s96 * om.a5, jB = uc * (hjk.R - hjk.P) - s96 * om.a1, jC = -wc * jkl.Q, jD = wc * jkl.P;
    const double kA = s96 * om.a5 - uc * hjk.Q, kB = uc * hjk.P, kC = wc * (jkl.Q - jkl.P), kD = wc * (jkl.R - jkl.P) - s96 * om.a4;
    const double A[3] = {gjk.y, gjk.z, gjk.w}, B[3] = {ghj.y, ghj.z, ghj.w}, C[3] = {gkj.y, gkj.z, gkj.w}, D[3] = {gkl.y, gkl.z, gkl.w};
#pragma unroll
    for (int t = 0; t < 3; t++) {
      atomicAdd(&v.f[3 * i + t], -(iA * A[t] + iB * B[t] - s96 * dvec_li[t]));
      atomicAdd(&v.f[3 * j + t], -(jA * A[t] + jB * B[t] + jC * C[t] + jD * D[t]));
      atomicAdd(&v.f[3 * k + t], -(kA * A[t] + kB * B[t] + kC * C[t] + kD * D[t]));
      atomicAdd(&v.f[3 * l + t], -(lC * C[t] + lD * D[t] + s96 * dvec_li[t]));
    }
  }
  const int slots[2] = {E_TOR, E_CON};
  double vals[2] = {e_tor, e_con};
  block_commit<2>(v.en, slots, vals);
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32)
k_dbond(DevView v, BondedWork W) {
  // 8 lanes per atom, 4 atoms per warp: a bond row has ~7.5 entries (20 when hot-compressed), so a warp per atom left
  // three quarters of the lanes idle
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwg = (gridDim.x * blockDim.x) >> 5;
  for (int i0 = 4 * wg; i0 < v.N; i0 += 4 * nwg) {
    const int i = i0 + grp;
    const bool live = i < v.N;
    const int start = live ? v.b_start[i] : 0, cnt = live ? v.b_cnt[i] : 0;
    double fx = 0, fy = 0, fz = 0, buf_c = 0;
    if (cnt > 0) {
      const double dsx = v.dDeltap_self[3 * i], dsy = v.dDeltap_self[3 * i + 1], dsz = v.dDeltap_self[3 * i + 2];
      const double cdd_i = v.CdDelta[i];
      const double2 s56_i = W.sum56[i];  // (sum CEval5, sum CEval6) of the angles centred on i (zero for ghosts)
      for (int e = sub; e < cnt; e += 8) {
        const int p = start + e;
        const int j = v.b_nbr[p];
        const int sym = v.b_sym[p];
        if (sym < 0) continue;
        const double2 s56_j = W.sum56[j];
        double Cdbo = v.b_Cdbo[p] + v.b_Cdbo[sym];
        double Cdbopi = v.b_Cdbopi[p] + v.b_Cdbopi[sym];
        double Cdbopi2 = v.b_Cdbopi2[p] + v.b_Cdbopi2[sym];
        if (s56_i.x != 0.0 || s56_i.y != 0.0 || s56_j.x != 0.0 || s56_j.y != 0.0) {
          const double bo_p = v.b_bo[p].x, bo_s = v.b_bo[sym].x;
          const double p3 = bo_p * bo_p * bo_p, q3 = bo_s * bo_s * bo_s;
          Cdbo += s56_i.y * (p3 * p3 * bo_p) + s56_j.y * (q3 * q3 * bo_s);
          Cdbopi += s56_i.x + s56_j.x;
          Cdbopi2 += s56_i.x + s56_j.x;
        }
        const double cdd = cdd_i + v.CdDelta[j];
        if (Cdbo == 0.0 && Cdbopi == 0.0 && Cdbopi2 == 0.0 && cdd == 0.0) continue;
        const double4 c1 = v.b_c1[p], c2 = v.b_c2[p], c3 = v.b_c3[p], der = v.b_der[p], geo = v.b_geo[p];
        const double C1dbo = c1.x * Cdbo, C2dbo = c1.y * Cdbo;
        const double C1dbopi = c1.w * Cdbopi, C2dbopi = c2.x * Cdbopi, C3dbopi = c2.y * Cdbopi;
        const double C1dbopi2 = c2.w * Cdbopi2, C2dbopi2 = c3.x * Cdbopi2, C3dbopi2 = c3.y * Cdbopi2;
        const double C1dDelta = c1.x * cdd, C2dDelta = c1.y * cdd;
        // temp = sum coef * vector; dBOp = der.x*dvec, dln_BOp_pi = der.y*dvec, dln_BOp_pi2 = der.z*dvec
        const double a_dbop = C1dbo + C1dDelta + C2dbopi + C2dbopi2;
        const double a_self = C2dbo + C2dDelta + C3dbopi + C3dbopi2;
        const double along = a_dbop * der.x + C1dbopi * der.y + C1dbopi2 * der.z;
        // reference adds temp to fCdDelta (= -force)
        fx -= along * geo.y + a_self * dsx;
        fy -= along * geo.z + a_self * dsy;
        fz -= along * geo.w + a_self * dsz;
        buf_c += -a_self;
      }
    }
    // sums over the 8 lanes of the atom (all 32 lanes take part)
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      buf_c += __shfl_xor_sync(0xffffffffu, buf_c, o, 8);
      fx += __shfl_xor_sync(0xffffffffu, fx, o, 8);
      fy += __shfl_xor_sync(0xffffffffu, fy, o, 8);
      fz += __shfl_xor_sync(0xffffffffu, fz, o, 8);
    }
    if (sub == 0 && (fx != 0.0 || fy != 0.0 || fz != 0.0)) {
      atomicAdd(&v.f[3 * i], fx); atomicAdd(&v.f[3 * i + 1], fy); atomicAdd(&v.f[3 * i + 2], fz);
    }
    if (buf_c != 0.0)
      for (int e = sub; e < cnt; e += 8) {
        const int p = start + e;
        const double4 geo = v.b_geo[p];
        const double c = -buf_c * v.b_der[p].x;
        fadd3(v.f, v.b_nbr[p], c, geo.y, geo.z, geo.w);
      }
  }
}

}  // namespace

// part 1 needs only the bond list; part 2 also needs this step's far list (hydrogen-bond partners)
void launch_bonded_part1(System& s, DevView& v, const DevParams& P, cudaStream_t st) {
  if (v.n == 0) return;
  BondedWork W = s.bonded_work();
  RXB_CUDA(cudaMemsetAsync(W.n_ang, 0, 2 * sizeof(int), st));   // n_ang, n_tor (n_hb belongs to K-farH)
  RXB_CUDA(cudaMemsetAsync(W.sum56, 0, (size_t)v.N * sizeof(double2), st));
  const int t = s.tick(StepTimers::MULTI, st);
  k_multi<<<(v.n + kWarps * 32 - 1) / (kWarps * 32), kWarps * 32, 0, st>>>(v, P);
  s.tock(t, st);
  s.kernel_launches += 1;
}
void launch_bonded_part2(System& s, DevView& v, const DevParams& P, cudaStream_t st) {
  if (v.n == 0) return;
  BondedWork W = s.bonded_work();
  int t = s.tick(StepTimers::ENUM, st);
  static int occ_enum = 0, occ_hb = 0, occ_ang = 0, occ_tor = 0;
  // one wave of resident CTAs each (measured for the item kernels: 1 / 2 / 3 waves 1.81 / 1.82 / 1.84 ms for the chain;
  // the former fixed 148 x 8 grids were 1.6 - 2.7 waves: 1.96 ms)
  constexpr int kItemWaves = 1;
  const size_t smem_enum = (size_t)kWarps * 4 * v.strong_cap * sizeof(int);
  static size_t smem_enum_set = 0;
  if (smem_enum > 48 * 1024 && smem_enum > smem_enum_set) {
    RXB_CUDA(cudaFuncSetAttribute(k_enum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_enum));
    smem_enum_set = smem_enum;
  }
  k_enum<<<wave_grid(k_enum, kWarps * 32, chain_waves(), occ_enum), kWarps * 32, smem_enum, st>>>(v, P, W);
  s.tock(t, st);
  t = s.tick(StepTimers::HBOND, st);
  k_hbond_items<<<wave_grid(k_hbond_items, kItemThreads, kItemWaves * chain_waves(), occ_hb), kItemThreads, 0, st>>>(v, P, W);
  s.tock(t, st);
  t = s.tick(StepTimers::VALTOR, st);
  // resident CTAs per SM the register allocation is bounded for (dev knobs RXB_ANG_OCC / RXB_TOR_OCC, A/B in
  // profiles/r02_item_occ_ab.txt)
  static const int ang_minb = getenv("RXB_ANG_OCC") ? atoi(getenv("RXB_ANG_OCC")) : 3;
  static const int tor_minb = getenv("RXB_TOR_OCC") ? atoi(getenv("RXB_TOR_OCC")) : 3;
#define RXB_ITEM_LAUNCH(K, MINB, OCC) \
  K<MINB><<<wave_grid(K<MINB>, kItemThreads, kItemWaves * chain_waves(), OCC), kItemThreads, 0, st>>>(v, P, W)
  switch (ang_minb) {
    case 4: RXB_ITEM_LAUNCH(k_angle_items, 4, occ_ang); break;
    case 5: RXB_ITEM_LAUNCH(k_angle_items, 5, occ_ang); break;
    default: RXB_ITEM_LAUNCH(k_angle_items, 3, occ_ang); break;
  }
  switch (tor_minb) {
    case 4: RXB_ITEM_LAUNCH(k_torsion_items, 4, occ_tor); break;
    case 5: RXB_ITEM_LAUNCH(k_torsion_items, 5, occ_tor); break;
    default: RXB_ITEM_LAUNCH(k_torsion_items, 3, occ_tor); break;
  }
#undef RXB_ITEM_LAUNCH
  s.tock(t, st);
  s.kernel_launches += 4;
}
void launch_bonded(System& s, DevView& v, const DevParams& P, cudaStream_t st) {
  launch_bonded_part1(s, v, P, st);
  launch_bonded_part2(s, v, P, st);
}

void launch_dbond(System& s, DevView& v, const DevParams& P, cudaStream_t st) {
  (void)P;
  if (v.N == 0) return;
  static int occ_dbond = 0;
  k_dbond<<<wave_grid(k_dbond, kWarps * 32, chain_waves(), occ_dbond), kWarps * 32, 0, st>>>(v, s.bonded_work());
  s.kernel_launches++;
}

}  // namespace rxb
