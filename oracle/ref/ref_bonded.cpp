// ORACLE/_REF — TEST INFRASTRUCTURE ONLY.
// Harness (written for this repo) that drives the reference's own LIVE serial energy routines, compiled UNMODIFIED from
// where they lie in /root/reference, on a state supplied by the caller:
//   * BOp_single                          reaxc_forces_sunway.cpp:677-825       uncorrected bond orders (a4)
//   * Torsion_Angles, control->virial = 1 reaxc_torsion_angles_sunway.cpp:562-1312   valence angle + torsion terms (a10)
//     (+ Calculate_Theta / Calculate_dCos_Theta / Calculate_Omega)
//   * Hydrogen_Bonds, control->virial = 1 reaxc_hydrogen_bonds_sunway.cpp:262-436    hydrogen bonds (a8)
//   * Add_dBond_to_Forces                 reaxc_bond_orders_sunway.cpp:186-331       bond-order chain rule (a11)
//   * Merge_Bonds_Atom_Energy_C_New       reaxc_multi_body_sw64.c:21-333 (MPE-side serial C)  bond, lone-pair, over- and
//                                         under-coordination energies (a7)
//   * vdW_Coulomb_Energy_Full_C_test_err  reaxc_nonbonded_sw64.c:40-258 (MPE-side serial C)   tapered vdW + Coulomb (a9)
//   * Init_Taper                          reaxc_init_md_sunway.cpp:100-136           Taper polynomial (a16)
//   * BO(), serial body                   reaxc_bond_orders_sunway.cpp:460-774       bond-order corrections (a6).  That body sits
//                                         behind an early `return;` in the live build; the Makefile compiles the same
//                                         unmodified file a second time with stubs/prelude_bo_serial.h, which defines the
//                                         single `return` of the file away, so BO() falls through into it after the
//                                         (no-op here) slave-core call.
// so that the CPU oracle (oracle/orc_forces.cpp) is pinned against the reference itself for these terms.  The Sunway
// slave-core entry points the same files reference (…_C) are never reached on these paths; they are defined below as
// traps so the library links.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pair_reaxc_sunway.h"
#include "reaxc_bond_orders_sunway.h"
#include "reaxc_control_sunway.h"
#include "reaxc_ffield_sunway.h"
#include "reaxc_forces_sunway.h"
#include "reaxc_hydrogen_bonds_sunway.h"
#include "reaxc_list_sunway.h"
#include "reaxc_multi_body_sw64.h"
#include "reaxc_torsion_angles_sunway.h"
#include "reaxc_types_sunway.h"

using namespace REAXC_SUNWAY_NS;

#define TRAP(name) { fprintf(stderr, "oracle/_ref: Sunway-only entry point %s reached\n", name); abort(); }
static bool g_bo_serial_run = false;   // set only while ref_bond_orders drives BO_serial_body
#define NOOP_IN_BO_SERIAL(name) { if (g_bo_serial_run) return; TRAP(name) }
extern "C" {
void Add_All_dBond_to_Forces_C(void*) TRAP("Add_All_dBond_to_Forces_C")
void Add_All_dBond_to_Forces_C_org(void*) TRAP("Add_All_dBond_to_Forces_C_org")
void BO_C(void*) NOOP_IN_BO_SERIAL("BO_C")
void Hydrogen_Bonds_C(void*) TRAP("Hydrogen_Bonds_C")
void Init_Forces_noQEq_Full_C(void*) TRAP("Init_Forces_noQEq_Full_C")
void Init_Forces_noQEq_HB_Full_C(void*) TRAP("Init_Forces_noQEq_HB_Full_C")
void Merge_Torsion_Valence_Angles(void*) TRAP("Merge_Torsion_Valence_Angles")
void Validate_Lists_C(void*) TRAP("Validate_Lists_C")
void swcache_init_locks(void*, int) TRAP("swcache_init_locks")
// MPE-side serial C routines (compiled with -DMPE and an athread stub): they take one pointer to a param pack
void Merge_Bonds_Atom_Energy_C_New(void* merge_bonds_eng_t_param);
void vdW_Coulomb_Energy_Full_C_test_err(void* vdw_coulomb_pack_t_param);
}
namespace REAXC_SUNWAY_NS {
// defined (non-static, undeclared in the header) at reaxc_forces_sunway.cpp:677
int BOp_single(storage* workspace, reax_list* bonds, double bo_cut, int i, int btop_i, far_neighbor_data_full* nbr_pj,
               single_body_parameters* sbp_i, single_body_parameters* sbp_j, two_body_parameters* twbp);
void Atom_Energy(reax_system*, control_params*, simulation_data*, storage*, reax_list**, output_controls*) TRAP("Atom_Energy")
void vdW_Coulomb_Energy_Full(reax_system*, control_params*, simulation_data*, storage*, reax_list**, output_controls*) TRAP("vdW_Coulomb_Energy_Full")
void Bonds(reax_system*, control_params*, simulation_data*, storage*, reax_list**, output_controls*) TRAP("Bonds")
void Init_Output_Files(reax_system*, control_params*, output_controls*, mpi_datatypes*, char*) TRAP("Init_Output_Files")
int Allocate_Workspace(reax_system*, control_params*, storage*, int, int, int, char*) TRAP("Allocate_Workspace")
void Init_Taper(control_params* control, storage* workspace, MPI_Comm comm);   // reaxc_init_md_sunway.cpp:100
void _reax_system::to_c_sys(reax_system_c*) NOOP_IN_BO_SERIAL("to_c_sys")
void _reax_system::from_c_sys(reax_system_c*) NOOP_IN_BO_SERIAL("from_c_sys")
// second compile of reaxc_bond_orders_sunway.cpp (see header): BO() with its serial body reachable
void BO_serial_body(reax_system*, control_params*, simulation_data*, storage*, reax_list**, output_controls*);
}

namespace {
struct Params {
  reax_system* sys;
  control_params control;
  output_controls oc;
};
// parse once per (control, ffield) pair with the reference's own readers
Params* load(const char* control_file, const char* ffield_file) {
  Params* P = new Params();
  P->sys = (reax_system*)calloc(1, sizeof(reax_system));
  memset(&P->control, 0, sizeof(P->control));
  memset(&P->oc, 0, sizeof(P->oc));
  Read_Control_File((char*)control_file, &P->control, &P->oc);
  FILE* fp = fopen(ffield_file, "r");
  if (!fp) return nullptr;
  Read_Force_Field(fp, &P->sys->reax_param, &P->control);
  return P;
}
}  // namespace

extern "C" {

void* ref_load(const char* control_file, const char* ffield_file) { return load(control_file, ffield_file); }

// BOp_single on a list of candidate pairs (i, j, d, dvec, element types).  out[p] = accepted flag followed by
// BO, BO_s, BO_pi, BO_pi2, dBOp[3], dln_BOp_pi[3], dln_BOp_pi2[3]  (14 doubles, after the bo_cut subtraction)
void ref_bop_pairs(void* hp, int npairs, const int* ti, const int* tj, const double* d, const double* dvec, double* out15) {
  Params* P = (Params*)hp;
  reax_interaction& rp = P->sys->reax_param;
  storage ws;
  memset(&ws, 0, sizeof(ws));
  rvec2 bo_dboc[1];
  rvec dself[1];
  ws.bo_dboc = bo_dboc;
  ws.dDeltap_self = dself;
  reax_list bonds;
  memset(&bonds, 0, sizeof(bonds));
  bond_data bd[1];
  bond_order_data bod[1];
  double Cdbo[1], Cdbopi[1], Cdbopi2[1], BO[1];
  rvec2 BOpi[1];
  bonds.select.bond_list = bd; bonds.bo_data_list = bod;
  bonds.Cdbo_list = Cdbo; bonds.Cdbopi_list = Cdbopi; bonds.Cdbopi2_list = Cdbopi2; bonds.BO_list = BO; bonds.BOpi_list = BOpi;
  for (int p = 0; p < npairs; p++) {
    far_neighbor_data_full nbr;
    memset(&nbr, 0, sizeof(nbr));
    nbr.nbr = 1; nbr.d = d[p];
    for (int t = 0; t < 3; t++) nbr.dvec[t] = dvec[3 * p + t];
    memset(bd, 0, sizeof(bd)); memset(bod, 0, sizeof(bod));
    bo_dboc[0][0] = bo_dboc[0][1] = 0; dself[0][0] = dself[0][1] = dself[0][2] = 0;
    BO[0] = 0; BOpi[0][0] = BOpi[0][1] = 0;
    const int acc = BOp_single(&ws, &bonds, P->control.bo_cut, 0, 0, &nbr, &rp.sbp[ti[p]], &rp.sbp[tj[p]], &rp.tbp[ti[p]][tj[p]]);
    double* o = out15 + 15 * p;
    o[0] = acc;
    if (acc) {
      o[1] = BO[0]; o[2] = bod[0].BO_s; o[3] = BOpi[0][0]; o[4] = BOpi[0][1];
      for (int t = 0; t < 3; t++) { o[5 + t] = bod[0].dBOp[t]; o[8 + t] = bod[0].dln_BOp_pi[t]; o[11 + t] = bod[0].dln_BOp_pi2[t]; }
      o[14] = bo_dboc[0][0];
    }
  }
}

// data->my_ext_press after the last ref_bonded call: what the control->virial = 1 (NPT) branches accumulate.  The LAMMPS
// interface writes rel_box = 0 for every neighbour (pair_reaxc_sunway.cpp:927), and so does this harness (memset), so
// every rvec_iMultiply(ext_press, rel_box, force) contribution is zero.
static double g_ext_press[3] = {0, 0, 0};
void ref_last_ext_press(double* out3) { for (int t = 0; t < 3; t++) out3[t] = g_ext_press[t]; }

// which: bit 0 = Torsion_Angles (valence + torsion), bit 1 = Hydrogen_Bonds, bit 2 = Add_dBond_to_Forces over i < j,
// bit 3 = Add_dBond_to_Forces_NPT over i < j (reaxc_bond_orders_sunway.cpp:41-180, the control->virial = 1 form).
// fields31 / w16 are the oracle's dumps (orc_get_bonds / orc_get_workspace); Cd_in[3][nb], CdDelta_in[N] seed the
// accumulators (zeros for an isolated term).  Outputs: en6 = e_ang,e_pen,e_coa,e_tor,e_con,e_hb; fCd[N][4] (-force,
// CdDelta); Cd_out[3][nb].
int ref_bonded(void* hp, int which, int n, int N, const double* x, const int* type, const int* tag, const int* b_start,
               const int* b_end, int nb, const int* nbr, const int* sym, const double* fields31, const double* w16,
               const double* dDeltap_self, const int* Hindex, int numH, const int* hb_start, const int* hb_end, int nhb,
               const int* hb_nbr, const double* hb_d, const double* hb_dvec, const double* Cd_in, const double* CdDelta_in,
               double* en6, double* fCd, double* Cd_out) {
  Params* P = (Params*)hp;
  reax_system* sys = P->sys;
  control_params* control = &P->control;
  control->virial = 1;   // selects the serial code paths
  LAMMPS_NS::Pair pair(nullptr);
  sys->pair_ptr = &pair;
  sys->n = n; sys->N = N; sys->numH = numH;
  std::vector<atom_pack_t> atoms(N);
  for (int i = 0; i < N; i++) {
    atoms[i].orig_id = tag[i]; atoms[i].type = type[i]; atoms[i].q = 0;
    for (int t = 0; t < 3; t++) atoms[i].x[t] = x[3 * i + t];
  }
  sys->packed_atoms = atoms.data();
  std::vector<int> hidx(Hindex, Hindex + N);
  sys->Hindex = hidx.data();

  simulation_data data;
  memset(&data, 0, sizeof(data));
  storage ws;
  memset(&ws, 0, sizeof(ws));
  std::vector<rvec2> bo_dboc(N);
  std::vector<double> Deltap(N), Deltap_boc(N), Delta(N), Delta_lp(N), Delta_lp_temp(N), Delta_e(N), Delta_val(N), dDelta_lp(N),
      dDelta_lp_temp(N), nlp(N), nlp_temp(N), Clp(N), vlpex(N);
  std::vector<rvec> dself(N);
  std::vector<rvec4> fcd(N);
  for (int i = 0; i < N; i++) {
    const double* w = w16 + 16 * i;
    bo_dboc[i][0] = w[0]; bo_dboc[i][1] = w[1]; Deltap[i] = w[2]; Deltap_boc[i] = w[3]; Delta[i] = w[4]; Delta_e[i] = w[5];
    Delta_val[i] = w[6]; vlpex[i] = w[7]; nlp[i] = w[8]; Delta_lp[i] = w[9]; Clp[i] = w[10]; dDelta_lp[i] = w[11];
    nlp_temp[i] = w[12]; Delta_lp_temp[i] = w[13]; dDelta_lp_temp[i] = w[14];
    for (int t = 0; t < 3; t++) { dself[i][t] = dDeltap_self[3 * i + t]; fcd[i][t] = 0.0; }
    fcd[i][3] = CdDelta_in ? CdDelta_in[i] : 0.0;
  }
  ws.bo_dboc = bo_dboc.data(); ws.Deltap = Deltap.data(); ws.Deltap_boc = Deltap_boc.data(); ws.Delta = Delta.data();
  ws.Delta_lp = Delta_lp.data(); ws.Delta_lp_temp = Delta_lp_temp.data(); ws.Delta_e = Delta_e.data(); ws.Delta_val = Delta_val.data();
  ws.dDelta_lp = dDelta_lp.data(); ws.dDelta_lp_temp = dDelta_lp_temp.data(); ws.nlp = nlp.data(); ws.nlp_temp = nlp_temp.data();
  ws.Clp = Clp.data(); ws.vlpex = vlpex.data(); ws.dDeltap_self = dself.data(); ws.fCdDelta = fcd.data();

  std::vector<reax_list> lists(LIST_N);
  memset(lists.data(), 0, sizeof(reax_list) * LIST_N);
  reax_list* bonds = &lists[BONDS];
  std::vector<int> bidx(b_start, b_start + N), bend(b_end, b_end + N);
  std::vector<bond_data> bd(nb > 0 ? nb : 1);
  std::vector<bond_order_data> bod(nb > 0 ? nb : 1);
  std::vector<double> Cdbo(nb > 0 ? nb : 1), Cdbopi(nb > 0 ? nb : 1), Cdbopi2(nb > 0 ? nb : 1), BO(nb > 0 ? nb : 1);
  std::vector<rvec2> BOpi(nb > 0 ? nb : 1);
  memset(bd.data(), 0, sizeof(bond_data) * bd.size());
  memset(bod.data(), 0, sizeof(bond_order_data) * bod.size());
  for (int p = 0; p < nb; p++) {
    const double* o = fields31 + 31 * p;
    bd[p].nbr = nbr[p]; bd[p].sym_index = sym[p]; bd[p].dbond_index = p; bd[p].d = o[0];
    for (int t = 0; t < 3; t++) {
      bd[p].dvec[t] = o[1 + t]; bod[p].dBOp[t] = o[8 + t]; bod[p].dln_BOp_pi[t] = o[11 + t]; bod[p].dln_BOp_pi2[t] = o[14 + t];
    }
    BO[p] = o[4]; bod[p].BO_s = o[5]; BOpi[p][0] = o[6]; BOpi[p][1] = o[7];
    bod[p].C1dbo = o[17]; bod[p].C2dbo = o[18]; bod[p].C3dbo = o[19];
    bod[p].C1dbopi = o[20]; bod[p].C2dbopi = o[21]; bod[p].C3dbopi = o[22]; bod[p].C4dbopi = o[23];
    bod[p].C1dbopi2 = o[24]; bod[p].C2dbopi2 = o[25]; bod[p].C3dbopi2 = o[26]; bod[p].C4dbopi2 = o[27];
    Cdbo[p] = Cd_in ? Cd_in[p] : 0.0; Cdbopi[p] = Cd_in ? Cd_in[nb + p] : 0.0; Cdbopi2[p] = Cd_in ? Cd_in[2 * nb + p] : 0.0;
  }
  bonds->n = N; bonds->num_intrs = nb; bonds->index = bidx.data(); bonds->end_index = bend.data(); bonds->type = TYP_BOND;
  bonds->select.bond_list = bd.data(); bonds->bo_data_list = bod.data();
  bonds->Cdbo_list = Cdbo.data(); bonds->Cdbopi_list = Cdbopi.data(); bonds->Cdbopi2_list = Cdbopi2.data();
  bonds->BO_list = BO.data(); bonds->BOpi_list = BOpi.data();

  reax_list* hbonds = &lists[HBONDS];
  std::vector<int> hs(hb_start, hb_start + (numH > 0 ? numH : 0)), he(hb_end, hb_end + (numH > 0 ? numH : 0));
  hs.push_back(0); he.push_back(0);
  std::vector<hbond_data> hb(nhb > 0 ? nhb : 1);
  std::vector<far_neighbor_data_full> far(nhb > 0 ? nhb : 1);
  memset(far.data(), 0, sizeof(far_neighbor_data_full) * far.size());
  for (int p = 0; p < nhb; p++) {
    far[p].nbr = hb_nbr[p]; far[p].d = hb_d[p];
    for (int t = 0; t < 3; t++) far[p].dvec[t] = hb_dvec[3 * p + t];
    hb[p].nbr = hb_nbr[p]; hb[p].scl = 1; hb[p].ptr = &far[p];
  }
  hbonds->n = numH; hbonds->num_intrs = nhb; hbonds->index = hs.data(); hbonds->end_index = he.data(); hbonds->type = TYP_HBOND;
  hbonds->select.hbond_list = hb.data();

  reax_list* lp = lists.data();
  if (which & 1) Torsion_Angles(sys, control, &data, &ws, &lp, &P->oc);
  if (which & 2) Hydrogen_Bonds(sys, control, &data, &ws, &lp, &P->oc);
  if (which & 4)
    for (int i = 0; i < N; i++)   // the stock driver loop: every bond once, from its lower-index end
      for (int pj = bidx[i]; pj < bend[i]; ++pj)
        if (i < bd[pj].nbr) Add_dBond_to_Forces(sys, i, pj, &ws, &lp);
  if (which & 8)
    for (int i = 0; i < N; i++)
      for (int pj = bidx[i]; pj < bend[i]; ++pj)
        if (i < bd[pj].nbr) Add_dBond_to_Forces_NPT(i, pj, &data, &ws, &lp);
  for (int t = 0; t < 3; t++) g_ext_press[t] = data.my_ext_press[t];

  en6[0] = data.my_en.e_ang; en6[1] = data.my_en.e_pen; en6[2] = data.my_en.e_coa;
  en6[3] = data.my_en.e_tor; en6[4] = data.my_en.e_con; en6[5] = data.my_en.e_hb;
  for (int i = 0; i < N; i++) for (int t = 0; t < 4; t++) fCd[4 * i + t] = fcd[i][t];
  for (int p = 0; p < nb; p++) { Cd_out[p] = Cdbo[p]; Cd_out[nb + p] = Cdbopi[p]; Cd_out[2 * nb + p] = Cdbopi2[p]; }
  return 0;
}

// Taper coefficients Tap[0..7] from the reference's Init_Taper on the parsed control parameters
void ref_taper(void* hp, double* tap8) {
  Params* P = (Params*)hp;
  storage ws;
  memset(&ws, 0, sizeof(ws));
  Init_Taper(&P->control, &ws, MPI_COMM_WORLD);
  for (int k = 0; k < 8; k++) tap8[k] = ws.Tap[k];
}

// a7: bonds + atom energies on the oracle's post-bond-order state.  en5 = e_bond, e_lp, e_ov, e_un, e_pol (the reference's
// running-prefix-sum e_pol, see SURVEY.md §8 "Semantics"); fCd[N][4] only receives CdDelta; Cd_out[3][nb].
int ref_atom_energy(void* hp, int enobondsflag, int n, int N, const double* q, const int* type, const int* tag, const int* b_start,
                    const int* b_end, int nb, const int* nbr, const int* sym, const double* fields31, const double* w16,
                    double* en5, double* eng_vdwl, double* fCd, double* Cd_out) {
  Params* P = (Params*)hp;
  control_params* control = &P->control;
  control->virial = 0;
  control->enobondsflag = enobondsflag;
  reax_system_c csys;
  memset(&csys, 0, sizeof(csys));
  csys.reax_param = P->sys->reax_param;
  csys.n = n; csys.N = N;
  std::vector<atom_pack_t> atoms(N);
  for (int i = 0; i < N; i++) { atoms[i].orig_id = tag[i]; atoms[i].type = type[i]; atoms[i].q = q[i]; atoms[i].x[0] = atoms[i].x[1] = atoms[i].x[2] = 0; }
  csys.packed_atoms = atoms.data();
  simulation_data data;
  memset(&data, 0, sizeof(data));
  storage ws;
  memset(&ws, 0, sizeof(ws));
  std::vector<rvec2> bo_dboc(N);
  std::vector<double> Delta(N), Delta_lp(N), Delta_lp_temp(N), Delta_e(N), Delta_val(N), dDelta_lp(N), dDelta_lp_temp(N), nlp(N),
      nlp_temp(N), Clp(N), vlpex(N);
  std::vector<rvec4> fcd(N);
  for (int i = 0; i < N; i++) {
    const double* w = w16 + 16 * i;
    bo_dboc[i][0] = w[0]; bo_dboc[i][1] = w[1]; Delta[i] = w[4]; Delta_e[i] = w[5]; Delta_val[i] = w[6]; vlpex[i] = w[7];
    nlp[i] = w[8]; Delta_lp[i] = w[9]; Clp[i] = w[10]; dDelta_lp[i] = w[11]; nlp_temp[i] = w[12]; Delta_lp_temp[i] = w[13];
    dDelta_lp_temp[i] = w[14];
    fcd[i][0] = fcd[i][1] = fcd[i][2] = fcd[i][3] = 0.0;
  }
  ws.bo_dboc = bo_dboc.data(); ws.Delta = Delta.data(); ws.Delta_lp = Delta_lp.data(); ws.Delta_lp_temp = Delta_lp_temp.data();
  ws.Delta_e = Delta_e.data(); ws.Delta_val = Delta_val.data(); ws.dDelta_lp = dDelta_lp.data();
  ws.dDelta_lp_temp = dDelta_lp_temp.data(); ws.nlp = nlp.data(); ws.nlp_temp = nlp_temp.data(); ws.Clp = Clp.data();
  ws.vlpex = vlpex.data(); ws.fCdDelta = fcd.data();
  std::vector<reax_list> lists(LIST_N);
  memset(lists.data(), 0, sizeof(reax_list) * LIST_N);
  reax_list* bonds = &lists[BONDS];
  std::vector<int> bidx(b_start, b_start + N), bend(b_end, b_end + N);
  std::vector<bond_data> bd(nb > 0 ? nb : 1);
  std::vector<bond_order_data> bod(nb > 0 ? nb : 1);
  std::vector<double> Cdbo(nb > 0 ? nb : 1, 0.0), Cdbopi(nb > 0 ? nb : 1, 0.0), Cdbopi2(nb > 0 ? nb : 1, 0.0), BO(nb > 0 ? nb : 1);
  std::vector<rvec2> BOpi(nb > 0 ? nb : 1);
  memset(bd.data(), 0, sizeof(bond_data) * bd.size());
  memset(bod.data(), 0, sizeof(bond_order_data) * bod.size());
  for (int p = 0; p < nb; p++) {
    const double* o = fields31 + 31 * p;
    bd[p].nbr = nbr[p]; bd[p].sym_index = sym[p]; bd[p].dbond_index = p; bd[p].d = o[0];
    for (int t = 0; t < 3; t++) bd[p].dvec[t] = o[1 + t];
    BO[p] = o[4]; bod[p].BO_s = o[5]; BOpi[p][0] = o[6]; BOpi[p][1] = o[7];
  }
  bonds->n = N; bonds->num_intrs = nb; bonds->index = bidx.data(); bonds->end_index = bend.data(); bonds->type = TYP_BOND;
  bonds->select.bond_list = bd.data(); bonds->bo_data_list = bod.data();
  bonds->Cdbo_list = Cdbo.data(); bonds->Cdbopi_list = Cdbopi.data(); bonds->Cdbopi2_list = Cdbopi2.data();
  bonds->BO_list = BO.data(); bonds->BOpi_list = BOpi.data();
  reax_list* lp = lists.data();
  merge_bonds_eng_t param;
  memset(&param, 0, sizeof(param));
  param.system = &csys; param.control = control; param.data = &data; param.workspace = &ws; param.lists = &lp;
  param.out_control = &P->oc;
  Merge_Bonds_Atom_Energy_C_New(&param);
  en5[0] = data.my_en.e_bond; en5[1] = data.my_en.e_lp; en5[2] = data.my_en.e_ov; en5[3] = data.my_en.e_un; en5[4] = data.my_en.e_pol;
  if (eng_vdwl) { eng_vdwl[0] = csys.eng_vdwl; eng_vdwl[1] = csys.eng_coul; }
  for (int i = 0; i < N; i++) for (int t = 0; t < 4; t++) fCd[4 * i + t] = fcd[i][t];
  for (int p = 0; p < nb; p++) { Cd_out[p] = Cdbo[p]; Cd_out[nb + p] = Cdbopi[p]; Cd_out[2 * nb + p] = Cdbopi2[p]; }
  return 0;
}

// a9: tapered van der Waals + Coulomb over the far list of the local atoms (full list, owner-computes).
// en2 = my_en.e_vdW, my_en.e_ele as the reference accumulates them (FULL value per directed pair, i.e. 2x the stock sum).
int ref_nonbonded(void* hp, int n, int N, const double* x, const double* q, const int* type, const int* tag, const long* far_off,
                  const int* far_nbr, const double* far_d, const double* far_dvec, const double* tap8, double* en2, double* fCd) {
  Params* P = (Params*)hp;
  control_params* control = &P->control;
  control->virial = 0;
  reax_system_c csys;
  memset(&csys, 0, sizeof(csys));
  csys.reax_param = P->sys->reax_param;
  csys.n = n; csys.N = N;
  std::vector<atom_pack_t> atoms(N);
  for (int i = 0; i < N; i++) {
    atoms[i].orig_id = tag[i]; atoms[i].type = type[i]; atoms[i].q = q[i];
    for (int t = 0; t < 3; t++) atoms[i].x[t] = x[3 * i + t];
  }
  csys.packed_atoms = atoms.data();
  simulation_data data;
  memset(&data, 0, sizeof(data));
  storage ws;
  memset(&ws, 0, sizeof(ws));
  for (int k = 0; k < 8; k++) ws.Tap[k] = tap8[k];
  std::vector<rvec4> fcd(N);
  for (int i = 0; i < N; i++) fcd[i][0] = fcd[i][1] = fcd[i][2] = fcd[i][3] = 0.0;
  ws.fCdDelta = fcd.data();
  std::vector<reax_list> lists(LIST_N);
  memset(lists.data(), 0, sizeof(reax_list) * LIST_N);
  reax_list* far = &lists[FAR_NBRS_FULL];
  const long nnz = far_off[n];
  std::vector<int> idx(N, 0), end(N, 0);
  for (int i = 0; i < n; i++) { idx[i] = (int)far_off[i]; end[i] = (int)far_off[i + 1]; }
  std::vector<far_neighbor_data_full> fl(nnz > 0 ? nnz : 1);
  memset(fl.data(), 0, sizeof(far_neighbor_data_full) * fl.size());
  for (long p = 0; p < nnz; p++) {
    fl[p].nbr = far_nbr[p]; fl[p].d = far_d[p]; fl[p].type = type[far_nbr[p]]; fl[p].orig_id = tag[far_nbr[p]]; fl[p].q = q[far_nbr[p]];
    for (int t = 0; t < 3; t++) fl[p].dvec[t] = far_dvec[3 * p + t];
  }
  far->n = N; far->num_intrs = (int)nnz; far->index = idx.data(); far->end_index = end.data();
  far->select.far_nbr_list_full = fl.data();
  reax_list* lp = lists.data();
  vdw_coulomb_pack_t param;
  memset(&param, 0, sizeof(param));
  param.system = &csys; param.control = control; param.data = &data; param.workspace = &ws; param.lists = &lp;
  param.out_control = &P->oc;
  vdW_Coulomb_Energy_Full_C_test_err(&param);
  en2[0] = data.my_en.e_vdW; en2[1] = data.my_en.e_ele;
  for (int i = 0; i < N; i++) for (int t = 0; t < 4; t++) fCd[4 * i + t] = fcd[i][t];
  return 0;
}

// a6: bond-order corrections on the oracle's post-bond-list state.  fields31 in: uncorrected BO', BO_s, BO_pi, BO_pi2 (and
// the geometric fields); total_bop[N] = sum of BO' per atom (bo_dboc[i][0] as Init_Forces leaves it).
// Out: fields31 with BO, BO_s, BO_pi, BO_pi2 and C1dbo..C4dbopi2 replaced; w16 as orc_get_workspace lays it out.
int ref_bond_orders(void* hp, int n, int N, const int* type, const int* tag, const int* b_start, const int* b_end, int nb,
                    const int* nbr, const int* sym, const double* total_bop, double* fields31, double* w16) {
  Params* P = (Params*)hp;
  reax_system* sys = P->sys;
  LAMMPS_NS::Pair pair(nullptr);
  sys->pair_ptr = &pair;
  sys->n = n; sys->N = N;
  std::vector<atom_pack_t> atoms(N);
  for (int i = 0; i < N; i++) { atoms[i].orig_id = tag[i]; atoms[i].type = type[i]; atoms[i].q = 0; atoms[i].x[0] = atoms[i].x[1] = atoms[i].x[2] = 0; }
  sys->packed_atoms = atoms.data();
  simulation_data data;
  memset(&data, 0, sizeof(data));
  storage ws;
  memset(&ws, 0, sizeof(ws));
  std::vector<rvec2> bo_dboc(N);
  std::vector<double> Deltap(N, 0.0), Deltap_boc(N, 0.0), Delta(N, 0.0), Delta_lp(N, 0.0), Delta_lp_temp(N, 0.0), Delta_e(N, 0.0),
      Delta_val(N, 0.0), dDelta_lp(N, 0.0), dDelta_lp_temp(N, 0.0), nlp(N, 0.0), nlp_temp(N, 0.0), Clp(N, 0.0), vlpex(N, 0.0);
  for (int i = 0; i < N; i++) { bo_dboc[i][0] = total_bop[i]; bo_dboc[i][1] = 0.0; }
  ws.bo_dboc = bo_dboc.data(); ws.Deltap = Deltap.data(); ws.Deltap_boc = Deltap_boc.data(); ws.Delta = Delta.data();
  ws.Delta_lp = Delta_lp.data(); ws.Delta_lp_temp = Delta_lp_temp.data(); ws.Delta_e = Delta_e.data(); ws.Delta_val = Delta_val.data();
  ws.dDelta_lp = dDelta_lp.data(); ws.dDelta_lp_temp = dDelta_lp_temp.data(); ws.nlp = nlp.data(); ws.nlp_temp = nlp_temp.data();
  ws.Clp = Clp.data(); ws.vlpex = vlpex.data();
  std::vector<reax_list> lists(LIST_N);
  memset(lists.data(), 0, sizeof(reax_list) * LIST_N);
  reax_list* bonds = &lists[BONDS];
  std::vector<int> bidx(b_start, b_start + N), bend(b_end, b_end + N);
  std::vector<bond_data> bd(nb > 0 ? nb : 1);
  std::vector<bond_order_data> bod(nb > 0 ? nb : 1);
  std::vector<double> Cdbo(nb > 0 ? nb : 1, 0.0), Cdbopi(nb > 0 ? nb : 1, 0.0), Cdbopi2(nb > 0 ? nb : 1, 0.0), BO(nb > 0 ? nb : 1);
  std::vector<rvec2> BOpi(nb > 0 ? nb : 1);
  memset(bd.data(), 0, sizeof(bond_data) * bd.size());
  memset(bod.data(), 0, sizeof(bond_order_data) * bod.size());
  for (int p = 0; p < nb; p++) {
    const double* o = fields31 + 31 * p;
    bd[p].nbr = nbr[p]; bd[p].sym_index = sym[p]; bd[p].dbond_index = p; bd[p].d = o[0];
    for (int t = 0; t < 3; t++) bd[p].dvec[t] = o[1 + t];
    BO[p] = o[4]; bod[p].BO_s = o[5]; BOpi[p][0] = o[6]; BOpi[p][1] = o[7];
  }
  bonds->n = N; bonds->num_intrs = nb; bonds->index = bidx.data(); bonds->end_index = bend.data(); bonds->type = TYP_BOND;
  bonds->select.bond_list = bd.data(); bonds->bo_data_list = bod.data();
  bonds->Cdbo_list = Cdbo.data(); bonds->Cdbopi_list = Cdbopi.data(); bonds->Cdbopi2_list = Cdbopi2.data();
  bonds->BO_list = BO.data(); bonds->BOpi_list = BOpi.data();
  reax_list* lp = lists.data();
  g_bo_serial_run = true;
  BO_serial_body(sys, &P->control, &data, &ws, &lp, &P->oc);
  g_bo_serial_run = false;
  for (int p = 0; p < nb; p++) {
    double* o = fields31 + 31 * p;
    o[4] = BO[p]; o[5] = bod[p].BO_s; o[6] = BOpi[p][0]; o[7] = BOpi[p][1];
    o[17] = bod[p].C1dbo; o[18] = bod[p].C2dbo; o[19] = bod[p].C3dbo;
    o[20] = bod[p].C1dbopi; o[21] = bod[p].C2dbopi; o[22] = bod[p].C3dbopi; o[23] = bod[p].C4dbopi;
    o[24] = bod[p].C1dbopi2; o[25] = bod[p].C2dbopi2; o[26] = bod[p].C3dbopi2; o[27] = bod[p].C4dbopi2;
  }
  for (int i = 0; i < N; i++) {
    double* w = w16 + 16 * i;
    w[0] = bo_dboc[i][0]; w[1] = bo_dboc[i][1]; w[2] = Deltap[i]; w[3] = Deltap_boc[i]; w[4] = Delta[i]; w[5] = Delta_e[i];
    w[6] = Delta_val[i]; w[7] = vlpex[i]; w[8] = nlp[i]; w[9] = Delta_lp[i]; w[10] = Clp[i]; w[11] = dDelta_lp[i];
    w[12] = nlp_temp[i]; w[13] = Delta_lp_temp[i]; w[14] = dDelta_lp_temp[i]; w[15] = 0.0;
  }
  return 0;
}

}  // extern "C"
