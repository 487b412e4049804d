// Device-side data model of the B200 ReaxFF path (sm_100a).
//
// Everything the hot path touches lives in HBM for the whole run (the reference has no residency concept:
// MPE and CPEs share DDR and every kernel DMA-tiles its inputs).  Layout decisions:
//   * atoms:  xq[N] = (x,y,z,q) one 32-byte record per atom, local atoms [0,n) then ghosts [n,N)
//             (replaces atom_pack_t AoS 40 B, reaxc_ctypes_sunway.h:233-238)
//   * lists:  CSR, int32 columns, int64 row offsets.  No 64-byte far_neighbor_data_full records
//             (reaxc_ctypes_sunway.h:575-583): distances are recomputed from xq where needed.
//   * bonds:  rows allocated from one atomic cursor, SoA in 32-byte groups (replaces bond_data 56 B +
//             bond_order_data 200 B + BO_list/BOpi_list/Cdbo* side arrays, reaxc_list_sunway.cpp:41-141)
#pragma once
#include <cuda_runtime.h>

#include "rxb_params.h"

namespace rxb {

struct DevParams {  // pointers into one HBM blob
  int nt;
  Control ctl;
  const double* gp;
  const AtomPar* atom;
  const PairPar* pair;
  const AngleSet* angle;
  const TorsPar* tors;
  const HbPar* hb;
  // tabulated long-range mode (ctl.tabulate > 0): 128-byte records [type pair][interval]{CEvd, CEclmb, e_vdW, e_ele}
  const double4* lut;
  int lut_n;
  double lut_dx, lut_inv_dx;
};

// energy accumulator slots (simulation_data::my_en order used by pvector, pair_reaxc_sunway.cpp:657-670)
enum EnSlot { E_BOND = 0, E_OV, E_UN, E_LP, E_ANG, E_PEN, E_COA, E_HB, E_TOR, E_CON, E_VDW, E_ELE, E_POL, E_NUM };

struct DevView {
  int n, N;            // local, local + ghost
  int cap_bonds;       // capacity of the bond arrays (directed bonds)
  // atoms
  double4* xq;         // N
  const float4* xf;    // N: fp32 shadow (x,y,z relative to the list origin, element type bits in .w), refreshed every step
  float far_band, bond_band;   // fp32 rounding bands (in r^2) around nonb_cut^2 / bond_cut^2, see CellList::fp32_band
  const int* type;     // N element index (-1 = NULL)
  const int* tag;      // N
  double* f;           // N*3, true forces
  double* CdDelta;     // N
  // Verlet list r <= cutneigh, local rows          (a1)
  const long long* vl_off; const int* vl_idx; const int* vl_cnt;   // row i: vl_idx[vl_off[i] .. + vl_cnt[i])
  // inner partition of the Verlet rows: the first vl_cnt_in[i] entries were within vl_cut_in (> the far cut-off) when the
  // list was built; *disp2 = max |x - x_build|^2 over all atoms this step.  A pair inside the far cut-off now was within
  // far + 2 sqrt(*disp2) at the build, so while that is <= vl_cut_in the far-list sweep reads only the inner block.
  const int* vl_cnt_in; const double* disp2; double vl_cut_in;
  // bond candidates r <= bond_cut + skin, all rows (a1, ghost rows included)
  const long long* bc_off; const int* bc_idx; const int* bc_cnt;
  // hbond candidates r <= hbond_cut + skin, local H rows only
  const long long* hc_off; const int* hc_idx;
  // far list == H sparsity pattern: r <= nonb_cut / swb, local rows, slots vl_off[i] .. +far_num[i]
  int* far_num; int* far_idx; double* H_val;
  // bonds
  int* b_start; int* b_cnt; int* b_cursor; int* overflow;
  int* b_nbr; int* b_sym; int* b_owner;
  double4* b_geo;      // d, dx, dy, dz          (dvec = x_nbr - x_i)
  double4* b_bo;       // BO, BO_s, BO_pi, BO_pi2
  double4* b_der;      // cBOp, cPi, cPi2, unused :  dBOp = cBOp*dvec, dln_BOp_pi = cPi*dvec, dln_BOp_pi2 = cPi2*dvec
  double4* b_c1;       // C1dbo, C2dbo, C3dbo, C1dbopi
  double4* b_c2;       // C2dbopi, C3dbopi, C4dbopi, C1dbopi2
  double4* b_c3;       // C2dbopi2, C3dbopi2, C4dbopi2, unused
  double* b_Cdbo; double* b_Cdbopi; double* b_Cdbopi2;
  // per-atom workspace (storage, reaxc_ctypes_sunway.h:675-738)
  double* total_bop;   // sum BO' (uncorrected)
  double2* Deltap;     // (Deltap, Deltap_boc)
  double* dDeltap_self;// N*3
  double* total_bo; double* Delta_boc; double* Delta; double* Delta_val; double* vlpex; double* nlp;
  double* Delta_lp; double* dDelta_lp; double* Delta_lp_temp;
  // energies / virial
  double* en;          // E_NUM
  double* virial;      // 6
};

// dense work lists of the bonded terms (rxb_bonded.cu: K-enum fills, the item kernels consume)
struct BondedWork {
  int4* ang;      // (j, pk, ph, -)
  int4* tor;      // (j, pk, ph, pw)
  int4* hb;       // (j, pi, k, -)
  int* n_ang; int* n_tor; int* n_hb;   // device cursors (contiguous: n_ang, n_tor, n_hb, pad)
  int cap_ang, cap_tor, cap_hb;
  double4* sbo;   // per local centre: SBO2, CSBO2, dSBO1, dSBO2
  double2* sum56; // per atom: sum CEval5, sum CEval6 over the angles centred on it
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double4 ld4(const double4* p) { return *p; }

// block-level accumulation of a few energy terms into global slots: one atomic per warp per slot
__device__ __forceinline__ void warp_commit(double* slot, double v) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(slot, v);
}

}  // namespace rxb
