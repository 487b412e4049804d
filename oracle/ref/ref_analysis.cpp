// ORACLE/_REF — TEST INFRASTRUCTURE ONLY.
// Harness (written for this repo) that runs the reference's own output fixes — fix_reaxc_bonds_sunway.cpp and
// fix_reaxc_species_sunway.cpp, compiled UNMODIFIED from /root/reference against the LAMMPS-core stand-in in
// stubs/lammps_stub.h — on a bond list / averaged bond orders supplied by the caller, and lets them write their files:
//   * FixReaxCBondsSunway:   constructor, init, end_of_step -> Output_ReaxC_Bonds = FindBond + PassBuffer + RecvBuffer
//   * FixReaxCSpeciesSunway: constructor (incl. `cutoff` keywords and the reneighbouring reset), init, setup,
//                            post_integrate -> FindMolecule (+ pack/unpack forward comm) + SortMolecule + FindSpecies +
//                            WriteFormulas.  The hidden `fix ave/atom` (LAMMPS core, absent) is a stand-in object whose
//                            averaged array the caller fills.
#include <math.h>

#include <vector>

#include "fix_ave_atom.h"
#include "fix_reaxc_bonds_sunway.h"
#include "fix_reaxc_species_sunway.h"
#include "pair_reaxc_sunway.h"
#include "reaxc_defs_sunway.h"
#include "reaxc_list_sunway.h"
#include "reaxc_types_sunway.h"

using namespace LAMMPS_NS;
using namespace REAXC_SUNWAY_NS;

namespace {
struct World {
  Atom atom; Comm comm; Memory memory; Error error; Force force; Neighbor neighbor; Update update; Group group; CiteMe cite;
  Domain domain; Modify modify;
  LAMMPS lmp;
  std::vector<int> type, tag, mask, ilist;
  std::vector<double> q;
  NeighList list;
  void* fake_pair = nullptr;
  World(int nlocal, int nghost, int ntypes, const int* ty, const int* tg, const double* qq, const int* ghost_owner, long natoms) {
    const int nall = nlocal + nghost;
    lmp.atom = &atom; lmp.comm = &comm; lmp.memory = &memory; lmp.error = &error; lmp.force = &force; lmp.neighbor = &neighbor;
    lmp.update = &update; lmp.group = &group; lmp.citeme = &cite; lmp.domain = &domain; lmp.modify = &modify;
    type.assign(ty, ty + nall); tag.assign(tg, tg + nall); mask.assign(nall, 1);
    q.assign(nall, 0.0);
    if (qq) q.assign(qq, qq + nall);
    atom.type = type.data(); atom.tag = tag.data(); atom.mask = mask.data(); atom.q = q.data();
    atom.nlocal = nlocal; atom.nghost = nghost; atom.nmax = nall + 16; atom.ntypes = ntypes; atom.natoms = natoms;
    comm.atom = &atom;
    if (ghost_owner) comm.ghost_owner.assign(ghost_owner, ghost_owner + nghost);
    group.natoms = nlocal;
    ilist.resize(nall);
    for (int i = 0; i < nall; i++) ilist[i] = i;
    list.inum = nlocal; list.gnum = nghost; list.ilist = ilist.data();
    fake_pair = calloc(1, sizeof(PairReaxCSunway) + 64);   // only public data members are ever read through it
    pair()->list = &list;
    pair()->listfull = &list;
    force.pair = pair();
  }
  ~World() { free(fake_pair); }
  PairReaxCSunway* pair() { return reinterpret_cast<PairReaxCSunway*>(fake_pair); }
};
}  // namespace

extern "C" {

// fix reax/c/bonds: one end_of_step() at `ntimestep`, appended to `path`.  BO[nb] = corrected bond orders in bond-row order.
int ref_bonds_write(const char* path, long ntimestep, double bg_cut, int nlocal, int nghost, int ntypes, const int* type,
                    const int* tag, const double* q, const int* b_start, const int* b_end, int nb, const int* nbr,
                    const double* BO, const double* total_bo, const double* nlp) {
  const int nall = nlocal + nghost;
  World W(nlocal, nghost, ntypes, type, tag, q, nullptr, nlocal);
  control_params control;
  memset(&control, 0, sizeof(control));
  control.bg_cut = bg_cut;
  storage ws;
  memset(&ws, 0, sizeof(ws));
  std::vector<rvec2> bo_dboc(nall);
  std::vector<double> nlpv(nlp, nlp + nall);
  for (int i = 0; i < nall; i++) { bo_dboc[i][0] = total_bo[i]; bo_dboc[i][1] = 0; }
  ws.bo_dboc = bo_dboc.data(); ws.nlp = nlpv.data();
  std::vector<reax_list> lists(LIST_N);
  memset(lists.data(), 0, sizeof(reax_list) * LIST_N);
  reax_list* bonds = &lists[BONDS];
  std::vector<int> bidx(b_start, b_start + nall), bend(b_end, b_end + nall);
  std::vector<bond_data> bd(nb > 0 ? nb : 1);
  memset(bd.data(), 0, sizeof(bond_data) * bd.size());
  for (int p = 0; p < nb; p++) bd[p].nbr = nbr[p];
  std::vector<double> bo(BO, BO + nb);
  bonds->n = nall; bonds->num_intrs = nb; bonds->index = bidx.data(); bonds->end_index = bend.data();
  bonds->select.bond_list = bd.data(); bonds->BO_list = bo.data();
  W.pair()->control = &control; W.pair()->workspace = &ws; W.pair()->lists = lists.data();
  W.update.ntimestep = ntimestep;
  char a0[] = "b", a1[] = "all", a2[] = "reax/c/bonds", a3[] = "1";
  std::vector<char> a4(path, path + strlen(path) + 1);
  char* args[5] = {a0, a1, a2, a3, a4.data()};
  try {
    FixReaxCBondsSunway fix(&W.lmp, 5, args);
    fix.init();
    fix.end_of_step();
  } catch (const std::exception& e) {
    fprintf(stderr, "oracle/_ref bonds: %s\n", e.what());
    return -2;
  }
  return 0;
}

// fix reax/c/species nevery nrepeat nfreq: the output step `nfreq` with tmpid (current bond partners, [nlocal][12] local
// indices, 0 = none) and avg_bo (the fix ave/atom result for the abo columns, [nlocal][12]).  cutoffs = ncut x (i, j, value).
// ... plus, when pos_path is given, the `position posfreq pos_path` keyword: avg_qxyz = the fix ave/atom result for the
// q, x, y, z columns ([nlocal][4]) and box6 = boxlo[3], boxhi[3] (WritePos, fix_reaxc_species_sunway.cpp:814-925).
int ref_species_write_pos(const char* path, int nevery, int nrepeat, int nfreq, int nlocal, int nghost, int ntypes, const int* type,
                          const int* tag, const int* ghost_owner, const int* tmpid, const double* avg_bo, int ncut,
                          const double* cutoffs, int* nmole_out, int* cluster_out, int* neighbor_every_out,
                          const char* pos_path, int posfreq, const double* avg_qxyz, const double* box6) {
  World W(nlocal, nghost, ntypes, type, tag, nullptr, ghost_owner, nlocal);
  if (box6) for (int t = 0; t < 3; t++) { W.domain.boxlo[t] = box6[t]; W.domain.boxhi[t] = box6[3 + t]; }
  const int nall = nlocal + nghost;
  // pair style arrays the fix reads: tmpid[i][jj]
  std::vector<int*> idrow(nall + 16);
  std::vector<int> idflat((size_t)(nall + 16) * MAXSPECBOND, 0);
  for (int i = 0; i < nall + 16; i++) idrow[i] = &idflat[(size_t)i * MAXSPECBOND];
  for (int i = 0; i < nlocal; i++) for (int k = 0; k < MAXSPECBOND; k++) idrow[i][k] = tmpid[i * MAXSPECBOND + k];
  W.pair()->tmpid = idrow.data();
  // the hidden fix ave/atom: columns 0..6 = q, x, y, z, vx, vy, vz (zero here), 7.. = abo01..
  char f0[] = "SPECBOND", f1[] = "all", f2[] = "ave/atom";
  char* fargs[3] = {f0, f1, f2};
  FixAveAtom ave(&W.lmp, 3, fargs);
  std::vector<double*> arow(nall + 16);
  std::vector<double> aflat((size_t)(nall + 16) * 31, 0.0);
  for (int i = 0; i < nall + 16; i++) arow[i] = &aflat[(size_t)i * 31];
  for (int i = 0; i < nlocal; i++) for (int k = 0; k < MAXSPECBOND; k++) arow[i][7 + k] = avg_bo[i * MAXSPECBOND + k];
  if (avg_qxyz) for (int i = 0; i < nlocal; i++) for (int k = 0; k < 4; k++) arow[i][k] = avg_qxyz[4 * i + k];
  ave.array_atom = arow.data();
  Fix* fixes[1] = {&ave};
  W.modify.nfix = 1; W.modify.fix = fixes;
  W.neighbor.every = 5; W.neighbor.delay = 0; W.neighbor.dist_check = 0;
  std::vector<std::string> sargs = {"s", "all", "reax/c/species", std::to_string(nevery), std::to_string(nrepeat),
                                    std::to_string(nfreq), path};
  for (int c = 0; c < ncut; c++) {
    char buf[64];
    sargs.push_back("cutoff");
    sargs.push_back(std::to_string((int)cutoffs[3 * c]));
    sargs.push_back(std::to_string((int)cutoffs[3 * c + 1]));
    snprintf(buf, sizeof buf, "%.17g", cutoffs[3 * c + 2]);
    sargs.push_back(buf);
  }
  if (pos_path) { sargs.push_back("position"); sargs.push_back(std::to_string(posfreq)); sargs.push_back(pos_path); }
  std::vector<std::vector<char>> store;
  std::vector<char*> args;
  for (auto& a : sargs) { store.emplace_back(a.begin(), a.end()); store.back().push_back(0); }
  for (auto& v : store) args.push_back(v.data());
  try {
    W.update.ntimestep = 0;
    FixReaxCSpeciesSunway fix(&W.lmp, (int)args.size(), args.data());
    if (neighbor_every_out) *neighbor_every_out = W.neighbor.every;
    fix.init();                       // nvalid = 0 + nfreq; creates (stubbed) compute + picks up the ave/atom fix
    fix.setup(0);                     // post_integrate at step 0: nothing to write yet
    W.update.ntimestep = nfreq;
    fix.post_integrate();             // FindMolecule, SortMolecule, FindSpecies, WriteFormulas
    if (nmole_out) *nmole_out = (int)fix.compute_vector(0);
    if (cluster_out) for (int i = 0; i < nlocal; i++) cluster_out[i] = (int)fix.vector_atom[i];
  } catch (const std::exception& e) {
    fprintf(stderr, "oracle/_ref species: %s\n", e.what());
    return -2;
  }
  return 0;
}

int ref_species_write(const char* path, int nevery, int nrepeat, int nfreq, int nlocal, int nghost, int ntypes, const int* type,
                      const int* tag, const int* ghost_owner, const int* tmpid, const double* avg_bo, int ncut,
                      const double* cutoffs, int* nmole_out, int* cluster_out, int* neighbor_every_out) {
  return ref_species_write_pos(path, nevery, nrepeat, nfreq, nlocal, nghost, ntypes, type, tag, ghost_owner, tmpid, avg_bo, ncut,
                               cutoffs, nmole_out, cluster_out, neighbor_every_out, nullptr, 0, nullptr, nullptr);
}

}  // extern "C"
