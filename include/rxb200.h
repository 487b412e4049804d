/* rxb200.h — C ABI of the B200-native ReaxFF force-and-charge path (librxb200.so).
 *
 * Drop-in boundary for the LAMMPS add-on styles of run-towards-the-future/SW_REAXFF.  In the reference every
 * accelerator call is an extern "C" void function taking one plain-C "param pack" (SURVEY.md §8b); the host owns all
 * buffers and the CPEs DMA them per call.  Here the device owns the state (atoms, lists, bonds, charges stay in HBM
 * across timesteps), so the entry points are coarser, one per LAMMPS style hook, and they return a status instead of
 * calling MPI_Abort:
 *      0  success        < 0  fatal, text via rxb_last_error()
 * All pointers are HOST pointers unless stated otherwise; arrays are C-contiguous, x/f are [nall][3] doubles as
 * LAMMPS' atom->x / atom->f, types are LAMMPS types (1-based), tags are 32-bit (LAMMPS_SMALLBIG).
 * One host thread per handle; a handle is bound to one CUDA device.  Not re-entrant per handle.
 */
#ifndef RXB200_H
#define RXB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rxb_handle rxb_handle;

/* ---- lifetime -------------------------------------------------------------------------------------------------- */
int rxb_create(int cuda_device, rxb_handle** out);
void rxb_destroy(rxb_handle* h);
const char* rxb_last_error(void);

/* ---- pair_style reax/c <control|NULL> [lgvdw yes/no] [enobonds yes/no]
 *      replaces PairReaxCSunway::settings, pair_reaxc_sunway.cpp:202-290 (Read_Control_File, reaxc_control_sunway.cpp:34) */
int rxb_pair_settings(rxb_handle* h, const char* control_file, int lgvdw, int enobonds);

/* ---- pair_coeff * * <ffield> <element per LAMMPS type | NULL>
 *      replaces PairReaxCSunway::coeff, pair_reaxc_sunway.cpp:294-362 (Read_Force_Field, reaxc_ffield_sunway.cpp:35) */
int rxb_pair_coeff(rxb_handle* h, const char* ffield_file, int ntypes, const char* const* elements);

/* ---- Pair::extract("chi"|"eta"|"gamma") : per LAMMPS type, out[0..ntypes] (index 0 unused)
 *      replaces PairReaxCSunway::extract, pair_reaxc_sunway.cpp:1106-1128 */
int rxb_pair_extract(rxb_handle* h, const char* name, double* out, int ntypes);

/* ---- fix qeq/reax <nevery> <swa> <swb> <tol> reax/c   (fix_qeq_reax_sunway.cpp:76-99); neighbor <skin> bin */
int rxb_fix_qeq(rxb_handle* h, double swa, double swb, double tolerance, int max_iter);
int rxb_neighbor_skin(rxb_handle* h, double skin);
/* fix qeq/reax ... <param file> instead of `reax/c` (FixQEqReaxSunway::pertype_parameters, fix_qeq_reax_sunway.cpp:198-245):
 * chi, eta, gamma per LAMMPS type, arrays of ntypes + 1 doubles (index 0 unused), used by the charge equilibration only
 * (H shielding, diagonal, right-hand side); the pair style keeps its force-field values.  ntypes = 0 or NULL arrays
 * return to the pair style's values. */
int rxb_fix_qeq_params(rxb_handle* h, int ntypes, const double* chi, const double* eta, const double* gamma);

/* canonical flat dump of every parsed parameter (parity tests); returns the count, writes min(count, cap) values */
long rxb_params_dump(rxb_handle* h, double* out, long cap);
/* host-only: parse control + ffield + element map without touching a GPU and dump as above; -1 on error */
long rxb_parse_dump(const char* control_file, const char* ffield_file, int ntypes, const char* const* elements,
                    int lgvdw, int enobonds, double* out, long cap);
/* Host-only: the spline tables of the tabulated long-range mode (control: tabulate_long_range N > 0; algorithm of the
 * reference's commented-out reaxc_lookup_sunway.cpp:157-285).  Layout [nt*nt][n = N+2][CEvd,CEclmb,e_vdW,e_ele][a,b,c,d];
 * returns the number of doubles (0 when tabulate is 0, -1 on error). */
long rxb_lookup_dump(const char* control_file, const char* ffield_file, int ntypes, const char* const* elements, int* n_out,
                     double* out, long cap);

/* ---- atoms: local [0,nlocal) then ghost [nlocal,nlocal+nghost)
 *      replaces write_reax_atoms_and_pack, pair_reaxc_sw64.c:114-190.
 *      ghost_owner[g] = local index whose image ghost g is (single-rank periodic images), or NULL to derive it from
 *      tags (atom->map); -1 = owned by another rank (the caller's comm layer must then refresh ghost values). */
int rxb_set_atoms(rxb_handle* h, int nlocal, int nghost, const double* x, const int* type, const int* tag,
                  const double* q, const int* ghost_owner);
/* every step, after forward_comm.  nall must be the nlocal + nghost of the last rxb_set_atoms (the index space the device
 * holds); a mismatch is an error, not a silent overrun. */
int rxb_set_positions(rxb_handle* h, int nall, const double* x);
int rxb_set_charges(rxb_handle* h, const double* q);

/* ---- neighbour build (every reneighbouring step)
 *      replaces NPairFullBin{Atomonly,Ghost}Sunway::build, npair_full_bin_atomonly_sunway.cpp:39-198,
 *      npair_full_bin_ghost_sw5.c:80-228: full list r <= cutmax+skin for local rows, plus the (bond_cut+skin) rows
 *      the bond-order kernels need for ghost atoms. */
int rxb_neigh_build(rxb_handle* h);

/* ---- FixQEqReaxSunway::pre_force (fix_qeq_reax_sunway.cpp:539-600): H build + dual-RHS pipelined CG + q.
 *      matvecs2[0..1] = iterations of the s and t solves (the reference's matvecs_s, matvecs_t). */
int rxb_qeq_pre_force(rxb_handle* h, int* matvecs2);
/* matvecs2 == NULL: the solve is only enqueued (as many iterations as the previous step took, plus a margin; iterations
 * past convergence are gated off on the device) and its convergence is checked together with the end-of-step status of
 * rxb_pair_compute, which continues it in the rare case the prediction fell short - no host round trip inside the step.
 * rxb_qeq_matvecs then returns the counts of that solve. */
int rxb_qeq_matvecs(rxb_handle* h, int* matvecs2);
/* H entry storage (SpMV stream): 0 (default) = packed 8-byte entries (22-bit column + 42-bit fixed-point value, absolute
 * quantisation 2^-39 ~ 1.8e-12 for TATB) whenever the taper starts at 0 and nall < 2^22; 1 = always fp64 value + int32
 * column (12 bytes).  Takes effect at the next rxb_qeq_pre_force. */
int rxb_set_h_exact(rxb_handle* h, int on);
int rxb_qeq_set_history(rxb_handle* h, const double* s_hist, const double* t_hist); /* [nlocal][5] */
int rxb_qeq_get_history(rxb_handle* h, double* s_hist, double* t_hist);
int rxb_get_charges(rxb_handle* h, double* q); /* nall */

/* ---- PairReaxCSunway::compute (pair_reaxc_sunway.cpp:541-793)
 *      f_out   : [nall][3] forces to ADD INTO atom->f by the caller (ghost rows included: caller reverse_comm's), or NULL
 *      pvector : 14 per-term energies in the reference's order (pair_reaxc_sunway.cpp:657-670), or NULL
 *      eng2    : eng_vdwl, eng_coul, or NULL ; virial6 : xx,yy,zz,xy,xz,yz, or NULL */
int rxb_pair_compute(rxb_handle* h, int nall, int eflag, int vflag, double* f_out, double* pvector, double* eng2,
                     double* virial6); /* nall: rows of f_out, checked against the device's atom count */

/* ---- device-resident run: fix nve (fix_nve_sw64.c:25-170) + periodic ghosts + the calls above, nothing leaves HBM.
 *      box6 = xprd,yprd,zprd,xy,xz,yz ; mass[0..ntypes] indexed by LAMMPS type */
int rxb_md_setup(rxb_handle* h, const double* box6, int nlocal, const double* x, const double* v, const int* type,
                 const int* tag, const double* mass, int ntypes, double dt, int reneigh_every, int thermo_every,
                 int qeq_on);
int rxb_md_run(rxb_handle* h, int nsteps);
double rxb_md_last_run_ms(rxb_handle* h); /* device time of the last rxb_md_run: CUDA events on the launch stream */
int rxb_md_get(rxb_handle* h, double* x, double* v, double* f, double* q); /* local atoms, any may be NULL */
int rxb_md_get_tags(rxb_handle* h, int* tags); /* tags of the current local atoms (they migrate in multi-GPU runs) */
int rxb_md_thermo(rxb_handle* h, double* pvector, double* pe, double* ke);

/* ---- multi-GPU: one process per GPU, bricks px*py*pz == world in lamda space, NCCL inside the library.
 *      Replaces what the reference gets from the LAMMPS core over MPI: exchange/borders, forward_comm(x), reverse_comm(f),
 *      forward_comm_fix of the CG direction and MPI_Allreduce of the dots (fix_qeq_reax_sunway.cpp:1043-1132).
 *      rank 0 calls rxb_dist_unique_id and broadcasts the 128 bytes; every rank then calls rxb_dist_init BEFORE
 *      rxb_md_setup, passing to rxb_md_setup only the atoms it currently holds (any initial assignment). */
int rxb_dist_unique_id(char* out128);
int rxb_dist_init(rxb_handle* h, int rank, int world, const char* id128, int px, int py, int pz);
/* halo transport: 1 (default) peer-to-peer boundary exchange with ncclSend/ncclRecv, 0 whole-slab all-gather */
int rxb_dist_set_p2p(rxb_handle* h, int on);

/* ---- multi-rank LAMMPS: the host keeps ITS decomposition and ghost order and drives one handle per MPI rank / GPU through
 *      the plugin calls above; the library runs what the reference does with comm->forward_comm_fix(this) and MPI_Allreduce
 *      inside the CG loop (fix_qeq_reax_sunway.cpp:1043-1140, pack/unpack_forward_comm :1300-1370) on the device, over
 *      NVLink peer windows or NCCL.  rxb_comm_init: once, instead of rxb_dist_init (id from rxb_dist_unique_id on rank 0,
 *      broadcast by the host, e.g. MPI_Bcast).  rxb_comm_set_ghosts: COLLECTIVE, after every rxb_set_atoms and before
 *      rxb_neigh_build; for ghost g (0 .. nghost-1, the host's order) owner_rank[g] is the rank that owns the real atom and
 *      owner_index[g] its local index THERE (periodic images of own atoms: owner_rank = this rank).  LAMMPS obtains both
 *      with one forward communication of (me, i) after borders(), see INTEGRATION.md.  In this mode rxb_pair_compute
 *      returns this rank's partial energies/virial and the forces on its local AND ghost atoms (the host reverse-
 *      communicates and reduces, as it does for the reference); rxb_get_charges returns q of local and ghost atoms. */
int rxb_comm_init(rxb_handle* h, int rank, int world, const char* id128);
int rxb_comm_set_ghosts(rxb_handle* h, int nghost, const int* owner_rank, const int* owner_index);

/* ---- introspection (tests, fix reax/c/bonds, fix reax/c/species) ---- */
/* counts[0..7] = nlocal, nall, verlet nnz, bond-candidate nnz, directed bonds, far nnz(sum), kernel launches, qeq iterations */
int rxb_get_counts(rxb_handle* h, long long* counts8);
int rxb_get_neighbors(rxb_handle* h, int which /*0 verlet, 1 bond candidates*/, long long* off, int* idx);
/* list cut-offs in use: out3 = Verlet list (cutmax + skin), bond-candidate list (bond reach + skin), bond reach = the
 * largest distance at which any element pair can still have BO' >= bo_cut (<= the control file's bond cutoff) */
int rxb_get_cutoffs(rxb_handle* h, double* out3);
/* bonds, CSR by atom (row = b_start[i] .. +b_cnt[i], ascending neighbour index), 31 doubles per directed bond:
 * d,dvec3,BO,BO_s,BO_pi,BO_pi2,dBOp3,dln_BOp_pi3,dln_BOp_pi2_3,C1..3dbo,C1..4dbopi,C1..4dbopi2,Cdbo,Cdbopi,Cdbopi2 */
int rxb_get_bonds(rxb_handle* h, int* b_start, int* b_cnt, int* nbr, int* sym, double* fields31);
/* per atom 16 doubles: total_bo,Delta_boc,Deltap,Deltap_boc,Delta,Delta_e,Delta_val,vlpex,nlp,Delta_lp,Clp,dDelta_lp,
 * nlp_temp,Delta_lp_temp,dDelta_lp_temp,CdDelta */
int rxb_get_workspace(rxb_handle* h, double* w16);
/* hydrogen-bond candidates of the last far-list sweep (the hbond list of Init_Forces_noQEq_HB_Full_C, reaxc_forces_sw64.c:
 * 787-863, which this path never stores as a list): pairs2[2k], pairs2[2k+1] = H atom, acceptor-type partner within
 * hbond_cut; unordered.  n_out = their number. */
int rxb_get_hbond_pairs(rxb_handle* h, int* n_out, int* pairs2, int cap);
/* far list == H pattern: num[nlocal], and for row i the entries off_verlet[i] .. +num[i] of idx/val */
int rxb_get_far(rxb_handle* h, int* num, int* idx, double* val);
/* ---- optional: page-lock the caller's per-atom arrays (atom->x, the force buffer) so that rxb_set_positions /
 * rxb_set_atoms / rxb_pair_compute copy them directly over PCIe instead of staging through an internal pinned buffer.
 * Pageable pointers keep working.  A page-locked x passed to rxb_set_positions must stay unchanged until the next
 * rxb_qeq_pre_force / rxb_pair_compute returns (LAMMPS' Verlet order guarantees that). */
int rxb_host_register(void* p, size_t bytes);
int rxb_host_unregister(void* p);

/* ---- fix reax/c/bonds (replaces FixReaxCBondsSunway::FindBond + PassBuffer, fix_reaxc_bonds_sunway.cpp:187-260) ----
 * Builds, on the device, the connection table of the local atoms from the bond list of the last force evaluation:
 * neighbours with BO > bo_cut (bo_cut < 0: the control file's bond_graph_cutoff, as the reference uses), in bond-row
 * order.  rxb_bond_table returns the sizes; rxb_bond_table_get copies tag/type[nlocal], CSR offsets off[nlocal+1],
 * neighbour IDs and bond orders [nentries], and abo (= total bond order), nlp, q [nlocal].  Any pointer may be NULL. */
int rxb_bond_table(rxb_handle* h, double bo_cut, int* nlocal, int* nentries, int* max_per_atom);
int rxb_bond_table_get(rxb_handle* h, int* tag, int* type, int* off, int* nbr_tag, double* bo, double* abo, double* nlp,
                       double* q);

/* ---- fix reax/c/species nevery nrepeat nfreq (fix_reaxc_species_sunway.cpp:60-117, 425-717; tmpid/tmpbo of
 * pair_reaxc_sunway.cpp:1170-1198; the averaging of the hidden fix ave/atom SPECBOND) ----
 * bocut = (ntypes+1)^2 matrix of BOCut[itype][jtype] (default 0.30 everywhere), natoms = global atom count (IDs must
 * be 1..natoms).  *reneighbor_reset = 1 when the reneighbouring period of the resident run had to be changed so that
 * lists stay frozen inside an averaging window (the reference prints "Resetting reneighboring criteria ...").
 * rxb_md_run calls the post_integrate hook itself and appends one record per output step to a log;
 * a host-driven loop calls rxb_species_step(ntimestep) after initial_integrate instead.
 * Result: nmole molecules ordered by their smallest atom ID; composition[m*ntypes + t] = atoms of type t+1 in molecule m
 * (summed over ranks when decomposed); cluster_of_local = molecule number 1..nmole per local atom. */
/* ---- compute SPEC/ATOM, abo columns (compute_spec_atom_sunway.cpp:35-170 reading PairReaxCSunway::tmpbo, filled by FindBond,
 * pair_reaxc_sunway.cpp:1170-1198): one sample, abo12[i*12 + k] = bond order (>= 0.10) of the k-th bond of local atom i to
 * a partner of higher index, in bond-row order, zero padded.  (The q/x/v columns of the compute are host data.) */
int rxb_spec_atom_abo(rxb_handle* h, double* abo12);
int rxb_species_config(rxb_handle* h, int nevery, int nrepeat, int nfreq, int ntypes, const double* bocut, long natoms,
                       long ntimestep_now /* <0: the resident run's own counter */, int* reneighbor_reset);
int rxb_species_step(rxb_handle* h, long ntimestep, int* found);
int rxb_species_result(rxb_handle* h, int* nmole, int* composition, long cap);
int rxb_species_cluster(rxb_handle* h, int* cluster_of_local);
/* `position` keyword (WritePos, fix_reaxc_species_sunway.cpp:814-925): the hidden fix ave/atom's result for the q, x, y, z
 * columns of compute SPEC/ATOM over the window that just ended, [nlocal][4]; valid right after rxb_species_step found an
 * output step */
int rxb_species_avg_qxyz(rxb_handle* h, double* qxyz4);
int rxb_species_log_size(rxb_handle* h);
int rxb_species_log_get(rxb_handle* h, int k, long* step, int* nmole, int* composition, long cap);

/* CUDA-event timers on the launch stream (no sync inside a step).  out28 = 14 accumulated ms then 14 call counts for:
 * neigh, qeq far+H, qeq CG (whole solve), bond list, BO, bonded (all), nonbonded, dBond, SpMV (per launch), hbond items,
 * angle+torsion items, multi-body, enumeration, SpMV boundary-row half (multi-GPU split).  enable: 1 reset+start, 0 stop,
 * -1 read only. */
int rxb_profile(rxb_handle* h, int enable, double* out28);
/* counters since rxb_create: out4 = SpMV launches that really multiplied (launches past convergence are gated off on the
 * device and not counted), QEq solves that had to be continued after the end-of-step check (force phase replayed),
 * dual-RHS CG iterations, kernel launches */
int rxb_get_counters(rxb_handle* h, long long* out4);
/* host-side waits on a CUDA stream or event issued by the library in this process so far (every one of them goes through
 * one counting macro): the difference around a run / the number of steps is what "no host in the loop" means in numbers */
long long rxb_host_sync_count(void);
/* Tests only: shrink the capacities of the growable lists (directed bonds, angle / torsion / hydrogen-bond work lists) and
 * of the per-atom shared-memory staging (bonds per atom, strong bonds per centre) so that every grow-and-replay branch
 * of the force phase can be driven on an ordinary cell (values <= 0 are left alone); read the current values back. */
int rxb_debug_set_caps(rxb_handle* h, int row_cap, int strong_cap, int cap_bonds, int cap_ang, int cap_tor, int cap_hb);
int rxb_debug_get_caps(rxb_handle* h, int* out6);
/* Storage format of one off-diagonal H entry (what the SpMV streams per non-zero): bytes per entry and a short name. */
int rxb_get_h_format(rxb_handle* h, int* bytes_per_entry, char* name, int cap);
/* Measured fp64 FMA throughput of the device in TFLOP/s (a pure DFMA kernel, best of 5): the roofline denominator of the
 * fp64-issue-bound kernels (bond orders, angles, torsions, nonbonded).  < 0 on error. */
double rxb_measure_fp64_tflops(int cuda_device);
/* cudaProfilerStart/Stop, so that `ncu --profile-from-start off` captures only the steady-state steps */
int rxb_profiler_range(int start);

#ifdef __cplusplus
}
#endif
#endif
