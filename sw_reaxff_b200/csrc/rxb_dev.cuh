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
  // rows sit at a fixed stride: row r = vl_idx[r * vl_stride .. + vl_cnt[r]); rows = local atoms in S order, columns = sorted
  // positions (see "S space" below)
  const long long* vl_off; const int* vl_idx; const int* vl_cnt; int vl_stride;
  // inner partition of the Verlet rows: the first vl_cnt_in[i] entries were within vl_cut_in (> the far cut-off) when the
  // list was built; *disp2 = max |x - x_build|^2 over all atoms this step.  A pair inside the far cut-off now was within
  // far + 2 sqrt(*disp2) at the build, so while that is <= vl_cut_in the far-list sweep reads only the inner block.
  const int* vl_cnt_in; const double* disp2; double vl_cut_in;
  // bond candidates r <= bond_cut + skin, all rows (a1, ghost rows included)
  const long long* bc_off; const int* bc_idx; const int* bc_cnt;
  // hbond candidates r <= hbond_cut + skin, local H rows only
  const long long* hc_off; const int* hc_idx;
  // ---- S space: the cell-sorted order of the last neighbour build.  The long-range machinery (Verlet list, far list, H,
  // nonbonded sweep, CG vectors) lives in it: row r = the r-th LOCAL atom in sorted order, columns = sorted positions of
  // all atoms (locals and ghosts interleaved as they lie in space).  Neighbours of a row are then runs of consecutive
  // integers, so the per-pair gathers of positions / charges / CG vectors touch whole cache lines instead of one sector
  // per lane.  The bonded machinery keeps the caller's atom order; s2a / row_atom translate.
  const int* s2a;        // N: sorted position -> atom index
  const int* rowpos;     // n: row -> sorted position
  const int* row_atom;   // n: row -> atom index
  const float4* xs;      // N: fp32 shadow in S order (origin-relative x,y,z; element type bits in .w), refreshed every step
  double4* xqs;          // N: exact (x,y,z,q) in S order, refreshed every step (q again after the QEq solve)
  const int* type_s;     // N: element index in S order
  // fix qeq/reax <param file> (fix_qeq_reax_sunway.cpp:198-245): chi / eta / gamma per LAMMPS type instead of the pair
  // style's per-element values.  null = the "reax/c" mode.  shld_lt[(lt_i) * nlt + lt_j] = (gamma_i gamma_j)^-1.5
  const int* ltype_s;    // N: LAMMPS type in S order
  const double* shld_lt; const double* chi_lt; const double* eta_lt; int nlt;
  // far list == H sparsity pattern: r <= nonb_cut / swb, row r in slots vl_off[r] .. +far_num[r].  Two storage formats:
  //  packed (default): one 64-bit word per entry = column (22 bits) << 42 | round(H * 2^h_shift) (42-bit fixed point);
  //  exact           : int32 column in far_idx + fp64 value in H_val (systems beyond 2^22 atoms per GPU, parity tests).
  int* far_num; int* far_idx; double* H_val;
  unsigned long long* hpk;
  double h_quant;        // 2^h_shift (packed format); the SpMV multiplies its row sums by 1 / h_quant
  // bonds
  int* b_start; int* b_cnt; int* b_cursor; int* overflow;
  // per-atom staging capacities of the bond-list / enumeration kernels (grown by the host on overflow bits 1 / 8, then the
  // force phase is replayed) and the largest need seen: need_row[0] bonds of one atom, need_row[1] strong bonds of one centre
  int row_cap, strong_cap; int* need_row;
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

// packed H entry (rxb_nonbonded.cu writes, rxb_qeq.cu / rxb_nonbonded.cu read)
constexpr int kHColShift = 42;
constexpr unsigned long long kHValMask = (1ULL << kHColShift) - 1;

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

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialisation attribute may be scheduled
// while its predecessor in the stream drains; pdl_wait() returns once the predecessor grid has completed and its writes
// are visible (a no-op without the attribute), pdl_release() lets the NEXT kernel's CTAs be scheduled as soon as every CTA
// of this grid has started.  Both are the first statements of the CG-loop kernels: the launch latency and the tail of
// one kernel overlap the ramp-up of the next (3 dependent launches per CG iteration, ~35 iterations per step).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
