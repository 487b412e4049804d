// Host-side owner of all device-resident state of the B200 ReaxFF path and the launch sequence of one timestep.
// Plays the role of the reference's reax_system + storage + reax_list[] + the *_sunway.cpp drivers
// (pair_reaxc_sunway.cpp:434-793, reaxc_forces_sunway.cpp:1297-1365, fix_qeq_reax_sunway.cpp:539-600),
// re-designed for residency: buffers are grown, never re-allocated per step, and nothing returns to the host
// inside a step except a few scalars.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "rxb_dev.cuh"
#include "rxb_params.h"

#define RXB_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " __FILE__ ":" + \
                               std::to_string(__LINE__));                                                \
  } while (0)

// every host-side wait on a stream goes through this: the count is what "no host in the loop" claims are checked against
// (rxb_host_sync_count, reported per step by bench.py)
inline long long& host_sync_counter() { static long long n = 0; return n; }
#define RXB_SYNC(stream)                       \
  do {                                         \
    ++::host_sync_counter();                   \
    RXB_CUDA(cudaStreamSynchronize(stream));   \
  } while (0)

namespace rxb {

template <class T>
struct DBuf {  // grow-only device buffer (contents are NOT preserved across growth unless resize_keep is used)
  T* p = nullptr;
  size_t cap = 0, n = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { if (p) cudaFree(p); }
  void resize(size_t m) {
    if (m > cap) {
      if (p) cudaFree(p);
      size_t want = m + m / 8 + 64;
      RXB_CUDA(cudaMalloc(&p, want * sizeof(T)));
      cap = want;
    }
    n = m;
  }
  void resize_keep(size_t m) {
    if (m > cap) {
      size_t want = m + m / 8 + 64;
      T* q = nullptr;
      RXB_CUDA(cudaMalloc(&q, want * sizeof(T)));
      if (p) { RXB_CUDA(cudaMemcpy(q, p, n * sizeof(T), cudaMemcpyDeviceToDevice)); cudaFree(p); }
      p = q;
      cap = want;
    }
    n = m;
  }
  size_t bytes() const { return cap * sizeof(T); }
};

// Neighbour rows at a fixed stride: row i occupies idx[off[i] .. off[i] + cnt[i]) with off[i] = i * stride.  The stride
// (max row length of the previous build + slack, a multiple of 32 ints so rows start on 128-byte lines) lets a rebuild run
// ONE pass — no count pass, no scan; a row that outgrows it triggers a re-run with a larger stride.
constexpr double kInnerSkin = 0.5;   // Angstrom: inner block of a Verlet row = far cut-off + this (adaptive skin, rxb_dev.cuh)

struct Csr {
  DBuf<long long> off;
  DBuf<int> idx, cnt;
  DBuf<int> cnt_in;        // rows are partitioned: the first cnt_in[i] entries were within cut_in at the build
  double cut_in = 0.0;     // 0: not partitioned
  DBuf<long long> stats;   // device: [0] max row length, [1] sum of row lengths
  long long nnz = 0;       // sum of cnt
  long long slots = 0;     // nrows * stride
  int nrows = 0;
  int stride = 0;
};

struct CellList {
  DBuf<double> bounds;
  DBuf<int> key, key2, val, sorted_idx, bin_count, bin_start, cnt;
  DBuf<long long> cntll;
  DBuf<double4> spos;
  DBuf<float4> sposf;
  DBuf<char> temp;
  char grid_blob[96];
  int num_bins = 0, reach = 2;
  double origin[3] = {0, 0, 0}, extent = 0.0;   // bounding box of the last bin() (origin of the fp32 shadows)
  float fp32_band(double cut) const;
  void bin(const double4* xq, int N, double bin_size, int reach, cudaStream_t st);
  // cut_in < cut: every row is partitioned, entries within cut_in (at the build) first; cut_in >= cut: no partition
  // rowpos != null: "S space" output (rxb_dev.cuh): row i = the local atom at sorted position rowpos[i], columns = sorted positions
  void build(const double4* xq, int nrows, double cut, double cut_in, Csr& out, cudaStream_t st, const int* rowpos = nullptr);
};

struct Box {  // LAMMPS triclinic box, lo = 0
  double h[6] = {1, 1, 1, 0, 0, 0};  // xprd, yprd, zprd, yz, xz, xy
  double h_inv[6];
  void set(double xprd, double yprd, double zprd, double xy, double xz, double yz);
};

struct Dist;  // multi-GPU state, rxb_dist.cu

struct StepTimers {
  // SPMV_B: the boundary-row half of a split SpMV (multi-GPU: interior rows run while the halo is in flight); its time
  // belongs to the SPMV launch of the same iteration, which alone counts the calls
  enum { NEIGH, QEQ_H, QEQ_CG, BONDS, BO, BONDED, NONB, DBOND, SPMV, HBOND, VALTOR, MULTI, ENUM, SPMV_B, NUM };
  double ms[NUM] = {0};
  long calls[NUM] = {0};
};

class System {
 public:
  explicit System(int device);
  ~System();

  // ---- configuration (pair_style / pair_coeff / fix qeq/reax arguments) ----
  ForceField ff;
  void upload_params();
  double qeq_swa = 0.0, qeq_swb = 10.0, qeq_tol = 1e-6;
  int qeq_imax = 200;       // fix_qeq_reax_sunway.cpp:998
  int qeq_check_every = 4;  // host polls the device convergence flags every this many iterations
  double skin = 2.5;

  // ---- atoms ----
  int n = 0, N = 0;
  void set_atoms(int nlocal, int nghost, const double* x, const int* ltype, const int* tag, const double* q,
                 const int* ghost_owner);
  void set_positions(const double* x_host);  // all N, host pointer (H2D)
  void set_charges(const double* q_host);    // all N
  void get_forces(double* f_host);           // all N (caller does reverse_comm, as LAMMPS does)
  void get_charges(double* q_host);          // all N

  // ---- lists ----
  void build_neighbors();   // a1: Verlet (local rows) + bond-candidate (all rows) lists
  double cutneigh() const;
  double bond_reach() const;

  // ---- QEq (fix qeq/reax pre_force) ----
  void qeq_reset_history();
  void qeq_set_history(const double* s_hist, const double* t_hist);  // [n][5] host
  void qeq_get_history(double* s_hist, double* t_hist);
  // wait_for_convergence: poll the device flags before returning (the caller wants matvecs_s/t now); false: enqueue the
  // predicted number of iterations and settle with the end-of-step status (qeq_settle)
  void qeq_pre_force(bool wait_for_convergence = true);
  bool qeq_settle();                    // true: the solve had to be continued, charges changed, replay the force phase
  void qeq_settle_now() { (void)qeq_settle(); }
  void qeq_iteration(int it);
  void qeq_forward_S(double2* vecS);    // ghosts of an S-space vector <- owners (periodic images / peer ranks)
  void qeq_spmv(const double2* xS, double2* y_row, bool gated, int parity, int r0 = 0, int r1 = -1);   // rows r0 .. r1 of the row list
  // multi-GPU: rows whose columns are all local (or own periodic images) come first in q_rowlist: they are multiplied while
  // the halo of the search direction is in flight (rxb_dist.cu dist_classify_rows)
  DBuf<int> q_rowlist;
  int n_interior_ = 0;
  static bool qeq_split_rows() { static const bool on = getenv("RXB_SPLIT") && atoi(getenv("RXB_SPLIT")) != 0; return on; }
  void dist_classify_rows();
  void dist_push2(double2* vecS, double* dots, int ndots);
  void dist_pull2(double2* vecS, double* dots, int ndots);
  void qeq_finish(bool shift_hist);
  bool qeq_poll();
  void plugin_qeq_pre_force(bool wait_for_convergence = true);   // C ABI entry: QEq with the bonded chain started on the second stream
  void plugin_compute(bool eflag, bool vflag);
  int matvecs_s = 0, matvecs_t = 0;
  long qeq_iters_total = 0;  // dual-RHS iterations launched and active (M2 metric)
  DBuf<unsigned long long> spmv_active_d;   // [1]: SpMV launches that were not gated off by the convergence flags
  long qeq_replays = 0;      // solves that had to be continued after the end-of-step check (force phase replayed)

  // storage format of the off-diagonal H entries (rxb_dev.cuh): packed 8-byte words unless exact is requested, the taper
  // does not start at 0 (values then leave [0, bound]) or the atom count exceeds the 22-bit column field
  bool h_exact_request = false;
  // fix qeq/reax <param file>: per-LAMMPS-type chi / eta / gamma used by the QEq instead of the pair style's (empty = reax/c)
  std::vector<double> qeq_chi_lt, qeq_eta_lt, qeq_gamma_lt;
  DBuf<double> qeq_lt_d;     // [chi | eta | shld matrix]
  DBuf<int> ltype_s;
  void qeq_set_type_params(int ntypes, const double* chi, const double* eta, const double* gamma);
  int h_bytes_per_entry() const { return h_packed_ ? 8 : 12; }
  const char* h_format_name() const {
    return h_packed_ ? "22-bit column + 42-bit fixed-point value in one 64-bit word (8 B)" : "fp64 value + int32 column (12 B)";
  }
  // tests: shrink the list / staging capacities (values <= 0 are left alone) so that every grow-and-replay branch of
  // compute() can be driven on an ordinary cell; read them back (row_cap, strong_cap, cap_bonds, cap_ang, cap_tor, cap_hb)
  void debug_set_caps(int row_cap, int strong_cap, int cap_bonds_, int cap_ang_, int cap_tor_, int cap_hb_) {
    if (row_cap > 0) row_cap_ = row_cap;
    if (strong_cap > 0) strong_cap_ = strong_cap;
    if (cap_bonds_ > 0) cap_bonds = cap_bonds_;
    if (cap_ang_ > 0) cap_ang = cap_ang_;
    if (cap_tor_ > 0) cap_tor = cap_tor_;
    if (cap_hb_ > 0) cap_hb = cap_hb_;
  }
  void debug_get_caps(int* o) const { o[0] = row_cap_; o[1] = strong_cap_; o[2] = cap_bonds; o[3] = cap_ang; o[4] = cap_tor; o[5] = cap_hb; }
  // introspection in the caller's index space (tests): compact CSR over local atoms, columns = atom indices
  void export_verlet(long long* off, int* idx);
  void export_far(int* num, int* idx, double* val);

  // ---- pair compute ----
  void compute(bool eflag, bool vflag);
  double energies[E_NUM] = {0};
  double virial[6] = {0};
  int overflow_flag = 0;
  int num_bonds = 0;         // directed bonds this step

  // ---- device-resident MD (mini LAMMPS core: fix nve + periodic ghosts) ----
  Box box;
  void md_setup(const double* box6, int nlocal, const double* x, const double* v, const int* ltype, const int* tag,
                const double* mass_by_type, int ntypes, double dt, int every);
  void md_run(int nsteps);
  void md_get(double* x, double* v, double* f, double* q);  // local atoms
  double md_kinetic();
  long ntimestep = 0;
  int md_every = 5, md_ago = 0;
  double md_dt = 0.0625;
  bool qeq_on = true;
  bool overlap = true;       // run the bond-order chain on a second stream concurrently with the QEq solve
  int md_thermo = 5;         // energies are reduced every md_thermo steps (thermo 5, in.reaxc.lattice:832)

  // ---- multi-GPU (rxb_dist.cu): brick decomposition + NCCL ----
  static void dist_unique_id(char* out128);
  void dist_init(int rank, int world, const char* id128, int px, int py, int pz);
  void dist_destroy();
  int dist_world() const;
  void dist_allreduce(double* dev_ptr, int count);
  int dist_rank() const;
  void dist_allreduce_int(int* dev_ptr, size_t count);
  void dist_sum_small(double* dev_ptr, int count);   // <= 8 scalars; through the peer windows when they are active
  void dist_peer_setup();
  bool dist_peer_active() const;
  const int* dist_peer_err_ptr() const;   // device flag set when a wait of the peer exchange timed out
  void dist_allreduce_max_int(int* dev_ptr, size_t count);
  void dist_allgather_int(const int* send, int* recv, size_t count_per_rank);
  void dist_exchange();          // exchange + borders at reneighbouring

  void dist_set_p2p(bool on);    // false: whole-slab all-gather halos (debug / comparison)
  void dist_forward_xq();        // ghosts <- owners (x + image shift, q)
  void dist_forward2(double2* vec);
  void dist_forward2_dots(double2* vec, double* dots);   // halo of vec + all-reduce of 4 dot products in one exchange
  void dist_reverse_f();
  size_t slab() const;           // (0: no all-gathered arrays any more)
  // host-planned halo: a multi-rank LAMMPS keeps its own decomposition and drives one handle per rank through the plugin
  // calls; the ranks share an NCCL communicator and each tells the library, per ghost, who owns the real atom
  void comm_init(int rank, int world, const char* id128);
  void comm_set_ghosts(int nghost, const int* owner_rank, const int* owner_index);   // collective, after set_atoms
  bool dist_external() const;
  const int* dist_gs() const;    // sorted position of the ghost behind each slot of the exchange plan
  size_t dist_last_recv_bytes() const;   // payload received at the last reneighbouring (migrants + ghost records)

  // ---- analysis outputs (rxb_analysis.cu): fix reax/c/bonds table, fix reax/c/species molecules ----
  struct BondTable { int n = 0, entries = 0, max_nb = 0; };
  BondTable bond_table_build(double bo_cut);   // bo_cut < 0: control file bond_graph_cutoff
  void bond_table_get(int* tag_out, int* type_out, int* off_out, int* nbr_tag, double* bo, double* abo, double* nlp_out,
                      double* q_out);
  struct Species {
    bool on = false;
    int nevery = 1, nrepeat = 1, nfreq = 1, ntypes = 0;
    long nvalid_ave = -1, nvalid_out = -1;   // next sampling step (fix ave/atom) and next output step (fix reax/c/species)
    int irepeat = 0;
    int nmole = 0;
    long natoms = 0;
    std::vector<int> composition;            // [nmole][ntypes], molecules in ascending order of their smallest atom ID
  } species;
  struct SpeciesRecord { long step; int nmole; std::vector<int> composition; };
  std::vector<SpeciesRecord> species_log;      // one record per output step reached inside md_run
  // returns 1 when the reneighbouring period had to be reset (the reference's warning, fix_reaxc_species_sunway.cpp:84-108)
  int species_config(int nevery, int nrepeat, int nfreq, int ntypes, const double* bocut, long natoms_total, long now = -1);
  bool species_step(long step);                // the post_integrate hook of timestep `step`; true when molecules were found
  void species_sample();
  void species_avg_qxyz(double* out4);     // the averaged q, x, y, z columns of the last complete window: [nlocal][4]
  void spec_atom_abo(double* abo_host);    // compute SPEC/ATOM abo columns, one sample: [nlocal][12]
  void species_find();
  void species_get_cluster(int* cluster_of_local);   // 1..nmole per local atom (vector_atom of the reference fix)
  DBuf<int> bt_cnt, bt_off, bt_tag, sp_id, sp_edges, sp_edges_all, sp_parent, sp_flag, sp_molidx, sp_comp, sp_misc;
  DBuf<double> bt_bo, sp_acc, sp_bocut, sp_qxyz;   // sp_qxyz: [n][4] sums of q, x, y, z over the samples of the window
  BondTable bt_last_;
  int sp_n_ = -1;

  // introspection for parity tests
  DevView view();
  std::vector<double> params_dump() const { return ff.dump(); }
  cudaStream_t stream() const { return st_; }
  StepTimers timers;
  int tick(int which, cudaStream_t st = nullptr);  // lazy CUDA-event timers on the launch stream: no sync inside a step
  void tock(int id, cudaStream_t st = nullptr);
  void resolve_timers();
  double last_run_ms = 0.0;     // device time of the last md_run (CUDA events on the launch stream)
  bool profile = false;
  long kernel_launches = 0;

  // raw buffers (public: the C ABI copies them out for tests / fix reax/c/bonds)
  DBuf<double4> xq;
  DBuf<float4> xf;           // fp32 shadow in atom order (bond-candidate prefilter)
  // S space (rxb_dev.cuh): cell-sorted order of the last neighbour build
  DBuf<int> s2a, a2s, rowpos, row_atom, type_s, gs_pos, gs_own;
  DBuf<int> row_of_atom, img_off, img_cur, img_pos;   // periodic images of each row (CSR of sorted positions), single-rank runs
  bool img_valid_ = false;
  DBuf<long long> row_flag, row_scan;
  DBuf<float4> xs;
  DBuf<double4> xqs;
  void build_sorted_space();
  DBuf<double4> x_build;     // S-space positions at the last neighbour build (adaptive inner skin, rxb_dev.cuh)
  DBuf<double> disp2_d;      // [1]: max squared displacement from x_build, refreshed with the shadow positions
  void update_shadow(cudaStream_t st);   // xq -> xf (atom order), xs / xqs (S order), max displacement since the build
  void positions_changed() { shadow_valid_ = false; }
  DBuf<int> type, tag, ltype_d, ghost_owner;
  DBuf<double> f, CdDelta;
  Csr vl, bc;
  DBuf<int> far_num, far_idx;
  DBuf<double> H_val;
  DBuf<unsigned long long> hpk;
  DBuf<int> b_start, b_cnt, b_cursor, overflow, b_nbr, b_sym, b_owner;
  DBuf<double4> b_geo, b_bo, b_der, b_c1, b_c2, b_c3;
  DBuf<double> b_Cdbo, b_Cdbopi, b_Cdbopi2;
  DBuf<double> total_bop, dDeltap_self, total_bo, Delta_boc, Delta, Delta_val, vlpex, nlp, Delta_lp, dDelta_lp, Delta_lp_temp;
  DBuf<double2> Deltap;
  DBuf<double> en_d, virial_d;
  // bonded work lists
  DBuf<int4> it_ang, it_tor, it_hb;
  DBuf<int> it_count;                 // n_ang, n_tor, n_hb, pad
  DBuf<double4> vt_sbo;
  DBuf<double2> sum56;
  int cap_ang = 0, cap_tor = 0, cap_hb = 0;
  int num_ang = 0, num_tor = 0, num_hb = 0;
  BondedWork bonded_work();
  // qeq
  DBuf<double> q_s_hist, q_t_hist;   // [n][5]
  DBuf<double2> q_x, q_r, q_u, q_w, q_p, q_ss, q_v, q_z, q_q, q_b, q_m;  // (s,t) interleaved row vectors (length n)
  DBuf<double2> q_xS, q_d;           // S-space vectors (length N): initial guess, search direction
  DBuf<double> q_eta;
  DBuf<double> q_Hdia_inv;
  DBuf<double> q_scal;               // device scalars
  DBuf<double> shld_d;               // nt*nt shielding (gamma_i gamma_j)^-1.5
  DBuf<double4> lut_d;               // spline tables of the tabulated long-range mode
  // md
  DBuf<double> v_d, mass_d, x_stage;
  DBuf<int> ghost_shift;             // per ghost 3 ints
  DBuf<long long> gcount, goff;
  DBuf<char> scan_temp;
  int cap_bonds = 0;
  DBuf<int> map_d;

 private:
  int device_;
  Dist* dist_ = nullptr;
  cudaStream_t st_ = nullptr, st2_ = nullptr;
  cudaEvent_t ev_fork_ = nullptr, ev_far_ = nullptr, ev_join_ = nullptr;
  bool hook_after_far_ = false;
  int chain_mode_ = 0;
  void after_far_hook();
  void md_force_overlapped(bool ev);
  void overlapped_front(bool wait_for_convergence);
  void choose_h_format();
  void dist_sorted_maps();               // S positions of the send lists / ghosts of the boundary exchange (rxb_dist.cu)
  void overlapped_back(bool eflag, bool vflag);
  void cancel_inflight();
  bool chain_inflight_ = false;          // the bonded chain of this step is queued on st2_ and not yet consumed
  std::vector<cudaEvent_t> ev_pool_;
  struct Pending { int which; int a, b; };
  std::vector<Pending> ev_pending_;
  size_t ev_used_ = 0;
  DBuf<char> param_blob_;
  DevParams dp_{};
  CellList cells_a_, cells_b_;
  bool qeq_ran_this_step_ = false;
  bool shadow_valid_ = false;            // xf / xs / xqs hold the current positions
  bool h_packed_ = false;
  double h_quant_ = 1.0;
  int qeq_predict_ = 0, qeq_it_ = 0;     // largest iteration count of the last 5 solves; loop index of the last launched sweep
  int qeq_recent_[5] = {0, 0, 0, 0, 0};
  unsigned qeq_recent_at_ = 0;
  void qeq_record_iterations();
  bool qeq_unsettled_ = false;
  double last_tap_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, last_swb_ = 0.0;   // what K-farH was last launched with (replays)
  cudaEvent_t run_ev_[2] = {nullptr, nullptr};
  double* h_pin_ = nullptr;  // pinned staging
  size_t h_pin_cap_ = 0;
  double* pin(size_t doubles);
  void h2d(void* dst, const void* src, size_t bytes, size_t stage_off_doubles);
  void step_forces(bool eflag, bool vflag);
  void read_step_status(bool ev, int* h2, int* wk4);
  DBuf<int> status_d_;
  int need_[11] = {0};                   // status slots of k_gather_status, maximum over all ranks
  int row_cap_ = 64, strong_cap_ = 32;   // shared-memory staging of the bond-list / enumeration kernels (grown on demand)
  DBuf<int> need_row_d;
  void grow_staging();
  void ensure_atom_capacity();
  void ensure_bond_capacity(int cap);
  void md_make_ghosts();
  void md_force();
  friend struct Launch;
};

// kernel launchers (one per translation unit)
void launch_bond_list(System& s, DevView& v, const DevParams& P, cudaStream_t st);
void launch_bond_orders(System& s, DevView& v, const DevParams& P, cudaStream_t st);
void launch_bonded(System& s, DevView& v, const DevParams& P, cudaStream_t st);
void launch_bonded_part1(System& s, DevView& v, const DevParams& P, cudaStream_t st);
void launch_bonded_part2(System& s, DevView& v, const DevParams& P, cudaStream_t st);
void launch_dbond(System& s, DevView& v, const DevParams& P, cudaStream_t st);
// Grid of a grid-stride kernel: a whole number of waves of the CTAs that fit on the 148 SMs at this kernel's register and
// shared-memory footprint (a fixed 148 x k grid is a fractional number of waves whenever the footprint changes: measured
// 13 % on the far-list kernel).  The occupancy is queried once per call site.
// launch with the programmatic-dependent-launch attribute (see pdl_wait in rxb_dev.cuh); RXB_PDL=0 launches plainly
inline bool pdl_enabled() { static const bool on = !(getenv("RXB_PDL") && atoi(getenv("RXB_PDL")) == 0); return on; }
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA launch failed: ") + cudaGetErrorString(e));
}

// RXB_CHAIN_WAVES (development knob, default 1): multiplies the wave count of the bonded-chain kernels - more, shorter CTAs,
// so that work queued on the low-priority stream hands the SMs back quickly when a solve kernel arrives
inline int chain_waves() { static const int w = getenv("RXB_CHAIN_WAVES") ? std::max(1, atoi(getenv("RXB_CHAIN_WAVES"))) : 1; return w; }
template <class Kernel>
inline int wave_grid(Kernel kernel, int threads, int waves, int& cache) {
  if (!cache) RXB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cache, kernel, threads, 0));
  return 148 * (cache > 0 ? cache : 1) * waves;
}

void launch_far_and_H(System& s, DevView& v, const DevParams& P, const double* qeq_tap, const double* shld, double swb,
                      cudaStream_t st);
void launch_nonbonded(System& s, DevView& v, const DevParams& P, bool evflag, cudaStream_t st);

}  // namespace rxb
