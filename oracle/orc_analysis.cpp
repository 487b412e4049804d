// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under sw_reaxff_b200/ may include, link or call this.
//
// CPU restatement of the two analysis fixes that read the ReaxFF bond list, single rank:
//   * fix reax/c/bonds   — FindBond, PassBuffer, RecvBuffer  (fix_reaxc_bonds_sunway.cpp:187-330)
//   * fix reax/c/species — PairReaxCSunway::FindBond (pair_reaxc_sunway.cpp:1170-1198), the hidden
//     `fix ave/atom nevery nrepeat nfreq` over compute SPEC/ATOM's abo columns (fix_reaxc_species_sunway.cpp:377-425;
//     FixAveAtom itself is LAMMPS core, absent from /root/reference: stock semantics restated), FindMolecule (:498-566)
//     with its iterated min-label sweeps and ghost forward_comm, SortMolecule (:570-648), FindSpecies (:652-717),
//     WriteFormulas (:745-780).
// Pinned (tests/test_oracle_vs_ref.py): the reference's own FixReaxCBondsSunway / FixReaxCSpeciesSunway, compiled unmodified
// against a LAMMPS-core stand-in (oracle/ref/ref_analysis.cpp), write byte-identical files from the same state.  Only the
// sampling schedule of the hidden fix ave/atom (LAMMPS core, absent from /root/reference) is restated from stock semantics.
#include "orc_analysis.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>

namespace orc {

namespace {
void appendf(std::string& s, const char* fmt, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  s += buf;
}
int nint(double r) {  // fix_reaxc_bonds_sunway.cpp:334-341
  int i = 0;
  if (r > 0.0) i = static_cast<int>(r + 0.5);
  else if (r < 0.0) i = static_cast<int>(r - 0.5);
  return i;
}
}  // namespace

std::string bonds_text(const MD& md, long ntimestep) {
  const System& s = md.sys;
  const int nlocal = md.nlocal;
  const double bo_cut = s.prm.bg_cut;
  // FindBond (:187-222): neighbours in bond-row order with BO > bg_cut
  std::vector<std::vector<int>> neighid(nlocal);
  std::vector<std::vector<double>> abo(nlocal);
  int numbonds = 0;
  for (int i = 0; i < nlocal; i++) {
    for (int pj = s.b_start[i]; pj < s.b_end[i]; ++pj) {
      const Bond& b = s.bonds[pj];
      if (b.BO > bo_cut) { neighid[i].push_back(s.tag[b.nbr]); abo[i].push_back(b.BO); }
    }
    numbonds = std::max(numbonds, (int)neighid[i].size());
  }
  // RecvBuffer (:264-330), one rank
  std::string out;
  appendf(out, "# Timestep %ld \n", ntimestep);
  out += "# \n";
  appendf(out, "# Number of particles %d \n", nlocal);
  out += "# \n";
  appendf(out, "# Max number of bonds per atom %d with coarse bond order cutoff %5.3f \n", numbonds, bo_cut);
  out += "# Particle connection table and bond orders \n";
  out += "# id type nb id_1...id_nb mol bo_1...bo_nb abo nlp q \n";
  for (int i = 0; i < nlocal; i++) {
    const int nb = (int)neighid[i].size();
    appendf(out, " %d %d %d", s.tag[i], md.ltype[i], nb);
    for (int k = 0; k < nb; k++) appendf(out, " %d", neighid[i][k]);
    appendf(out, " %d", 0);  // atom->molecule == NULL (atom_style charge)
    for (int k = 0; k < nb; k++) appendf(out, "%14.3f", abo[i][k]);
    appendf(out, "%14.3f%14.3f%14.3f\n", s.total_bo[i], s.nlp[i], s.q[i]);
  }
  out += "# \n";
  return out;
}

// ---------------------------------------------------------------------------------------------------------------------
void SpeciesFix::init(const MD& md, int nevery_, int nrepeat_, int nfreq_, const std::vector<double>& bocut_) {
  nevery = nevery_; nrepeat = nrepeat_; nfreq = nfreq_;
  ntypes = (int)md.mass.size() - 1;
  BOCut = bocut_;
  nvalid = md.ntimestep + nfreq;  // FixReaxCSpeciesSunway::init :318-319
  // FixAveAtom constructor: nvalid = nextvalid()
  long nv = (md.ntimestep / nfreq) * nfreq + nfreq;
  if (nv - nfreq == md.ntimestep && nrepeat == 1) nv = md.ntimestep;
  else nv -= (long)(nrepeat - 1) * nevery;
  if (nv < md.ntimestep) nv += nfreq;
  ave_nvalid = nv;
  irepeat = 0;
}

void SpeciesFix::pair_find_bond(const MD& md) {  // pair_reaxc_sunway.cpp:771-790, 1170-1198
  const System& s = md.sys;
  tmpid.assign((size_t)s.N * MAXSPECBOND, 0);
  tmpbo.assign((size_t)s.N * MAXSPECBOND, 0.0);
  for (int i = 0; i < md.nlocal; i++) {   // system->n
    int nj = 0;
    for (int pj = s.b_start[i]; pj < s.b_end[i]; ++pj) {
      const Bond& b = s.bonds[pj];
      const int j = b.nbr;
      if (j < i) continue;
      if (b.BO >= 0.10) {
        if (nj >= MAXSPECBOND) { error = "Increase MAXSPECBOND in reaxc_defs_sunway.h"; return; }
        tmpid[(size_t)i * MAXSPECBOND + nj] = j;
        tmpbo[(size_t)i * MAXSPECBOND + nj] = b.BO;
        nj++;
      }
    }
  }
}

// post_integrate of timestep `step` (:419-423): f_SPECBOND->end_of_step(), then output when step == nvalid
bool SpeciesFix::post_integrate(const MD& md, long step) {
  if (step == ave_nvalid) {            // FixAveAtom::end_of_step
    pair_find_bond(md);                // the pair style refreshed tmpid/tmpbo in its last compute()
    const size_t m = (size_t)md.nlocal * MAXSPECBOND;
    if (irepeat == 0) { array.assign(m, 0.0); qxyz.assign((size_t)4 * md.nlocal, 0.0); }
    for (size_t k = 0; k < m; k++) array[k] += tmpbo[k];
    for (int i = 0; i < md.nlocal; i++) {            // columns 0..3 of compute SPEC/ATOM: q, x, y, z (:142-170)
      qxyz[4 * (size_t)i] += md.sys.q[i];
      for (int t = 0; t < 3; t++) qxyz[4 * (size_t)i + 1 + t] += md.sys.x[3 * (size_t)i + t];
    }
    irepeat++;
    if (irepeat < nrepeat) ave_nvalid += nevery;
    else {
      irepeat = 0;
      ave_nvalid = step + nfreq - (long)(nrepeat - 1) * nevery;
      for (size_t k = 0; k < m; k++) array[k] /= nrepeat;
      for (double& a : qxyz) a /= nrepeat;
    }
  }
  if (step != nvalid) return false;
  find_molecule(md);
  sort_molecule(md);
  find_species(md);
  nvalid += nfreq;
  return true;
}

void SpeciesFix::find_molecule(const MD& md) {  // :498-566 (x0 anchors omitted: they never change the labelling)
  const System& s = md.sys;
  const int nlocal = md.nlocal;
  clusterID.assign(s.N, 0.0);
  for (int i = 0; i < nlocal; i++) clusterID[i] = s.tag[i];
  int loop = 0;
  while (true) {
    for (int g = nlocal; g < s.N; g++) clusterID[g] = clusterID[md.ghost_owner[g - nlocal]];  // comm->forward_comm_fix
    loop++;
    int change = 0;
    while (true) {
      int done = 1;
      for (int i = 0; i < nlocal; i++) {
        const int itype = md.ltype[i];
        for (int jj = 0; jj < MAXSPECBOND; jj++) {
          const int j = tmpid[(size_t)i * MAXSPECBOND + jj];
          if (j == 0 || j < i) continue;
          if (clusterID[i] == clusterID[j]) continue;
          const int jtype = j < nlocal ? md.ltype[j] : md.ltype[md.ghost_owner[j - nlocal]];
          const double bo_cut = BOCut[(size_t)itype * (ntypes + 1) + jtype];
          const double bo_tmp = array[(size_t)i * MAXSPECBOND + jj];
          if (bo_tmp > bo_cut) {
            clusterID[i] = clusterID[j] = std::min(clusterID[i], clusterID[j]);
            done = 0;
          }
        }
      }
      if (!done) change = 1;
      if (done) break;
    }
    // a lowered ghost label reaches its owner in the next pass through the owner's own mirrored bond to the image of i
    if (!change) break;
    if (loop >= 400) break;
  }
}

void SpeciesFix::sort_molecule(const MD& md) {  // :570-648
  const int nlocal = md.nlocal;
  int lo = 1 << 30, hi = -(1 << 30);
  for (int n = 0; n < nlocal; n++) { lo = std::min(lo, nint(clusterID[n])); hi = std::max(hi, nint(clusterID[n])); }
  const int nlen = hi - lo + 1;
  std::vector<int> molmap(nlen, 0);
  for (int n = 0; n < nlocal; n++) molmap[nint(clusterID[n]) - lo] = 1;
  Nmole = 0;
  for (int n = 0; n < nlen; n++) molmap[n] = molmap[n] ? Nmole++ : -1;
  for (int n = 0; n < nlocal; n++) clusterID[n] = molmap[nint(clusterID[n]) - lo] + 1;
}

void SpeciesFix::find_species(const MD& md) {  // :652-717
  const int nlocal = md.nlocal;
  std::vector<int> comp((size_t)Nmole * ntypes, 0);
  for (int n = 0; n < nlocal; n++) comp[(size_t)(nint(clusterID[n]) - 1) * ntypes + md.ltype[n] - 1]++;
  composition = comp;
  MolName.clear(); NMol.clear();
  Nspec = 0;
  for (int m = 0; m < Nmole; m++) {
    const int* Name = &comp[(size_t)m * ntypes];
    int flag_identity = 1;
    for (int k = 0; k < Nspec; k++) {
      int flag_spec = 0;
      for (int l = 0; l < ntypes; l++) if (MolName[(size_t)ntypes * k + l] != Name[l]) flag_spec = 1;
      if (flag_spec == 0) NMol[k]++;
      flag_identity *= flag_spec;
    }
    if (Nspec == 0 || flag_identity == 1) {
      for (int l = 0; l < ntypes; l++) MolName.push_back(Name[l]);
      NMol.push_back(1);
      Nspec++;
    }
  }
}

std::string SpeciesFix::formulas_text(long ntimestep) const {  // WriteFormulas :745-780
  static const char ele[4] = {'C', 'H', 'O', 'N'};   // default element letters by LAMMPS type (:236-245)
  std::string out = "# Timestep     No_Moles     No_Specs     ";
  for (int i = 0; i < Nspec; i++) {
    for (int j = 0; j < ntypes; j++) {
      const int itemp = MolName[(size_t)ntypes * i + j];
      if (itemp != 0) {
        appendf(out, "%c", ele[j]);
        if (itemp != 1) appendf(out, "%d", itemp);
      }
    }
    out += "\t";
  }
  out += "\n";
  appendf(out, "%ld", ntimestep);
  appendf(out, "%11d%11d\t", Nmole, Nspec);
  for (int i = 0; i < Nspec; i++) appendf(out, " %d\t", NMol[i]);
  out += "\n";
  return out;
}

// WritePos :814-925.  x0 of a molecule is the fixed point of FindMolecule's anchor propagation (:530-532, :559-569 with
// chAnchor :497-510): every atom starts from its own averaged position, bonded atoms take the lexicographically smaller
// (x, then y, then z) of their two anchors until nothing changes, ghosts carry their owner's anchor unshifted
// (pack/unpack_forward_comm :955-982) - i.e. the lexicographic minimum over the molecule.
std::string SpeciesFix::pos_text(const MD& md, long ntimestep, const double* box6) {
  static const char ele[4] = {'C', 'H', 'O', 'N'};
  const int nlocal = md.nlocal;
  const double* lo = box6;
  const double* hi = box6 + 3;
  const double box[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
  const double halfbox[3] = {box[0] / 2, box[1] / 2, box[2] / 2};
  // (the reference shifts the averaged columns in place; they are re-zeroed by the next averaging cycle before anything
  // reads them again, so a working copy is equivalent and keeps this routine repeatable)
  std::vector<double> col = qxyz;
  std::vector<double> anchor((size_t)3 * Nmole, 0.0);
  std::vector<char> have(Nmole, 0);
  for (int i = 0; i < nlocal; i++) {
    const int m = nint(clusterID[i]) - 1;
    const double* xi = &col[4 * (size_t)i + 1];
    double* a = &anchor[3 * (size_t)m];
    const bool less = !have[m] || xi[0] < a[0] || (xi[0] == a[0] && (xi[1] < a[1] || (xi[1] == a[1] && xi[2] < a[2])));
    if (less) { a[0] = xi[0]; a[1] = xi[1]; a[2] = xi[2]; have[m] = 1; }
  }
  std::string out;
  appendf(out, "Timestep %ld NMole %d  NSpec %d  xlo %f  xhi %f  ylo %f  yhi %f  zlo %f  zhi %f\n", ntimestep, Nmole, Nspec, lo[0],
          hi[0], lo[1], hi[1], lo[2], hi[2]);
  out += "ID\tAtom_Count\tType\tAve_q\t\tCoM_x\t\tCoM_y\t\tCoM_z\n";
  std::vector<int> Name(ntypes);
  for (int m = 1; m <= Nmole; m++) {
    int count = 0;
    double avq = 0.0, avx[3] = {0, 0, 0};
    std::fill(Name.begin(), Name.end(), 0);
    const double* x0 = &anchor[3 * (size_t)(m - 1)];
    for (int i = 0; i < nlocal; i++) {
      if (nint(clusterID[i]) != m) continue;
      Name[md.ltype[i] - 1]++;
      count++;
      double* sa = &col[4 * (size_t)i];
      avq += sa[0];
      for (int t = 0; t < 3; t++) {
        if ((x0[t] - sa[1 + t]) > halfbox[t]) sa[1 + t] += box[t];
        if ((sa[1 + t] - x0[t]) > halfbox[t]) sa[1 + t] -= box[t];
      }
      for (int t = 0; t < 3; t++) avx[t] += sa[1 + t];
    }
    appendf(out, "%d\t%d\t", m, count);
    for (int n = 0; n < ntypes; n++)
      if (Name[n] != 0) {
        appendf(out, "%c", ele[n]);
        if (Name[n] != 1) appendf(out, "%d", Name[n]);
      }
    if (count > 0) {
      avq /= count;
      for (int k = 0; k < 3; k++) {
        avx[k] /= count;
        if (avx[k] >= hi[k]) avx[k] -= box[k];
        if (avx[k] < lo[k]) avx[k] += box[k];
        avx[k] -= lo[k];
        avx[k] /= box[k];
      }
      appendf(out, "\t%.8f \t%.8f \t%.8f \t%.8f", avq, avx[0], avx[1], avx[2]);
    }
    out += "\n";
  }
  out += "#\n";
  return out;
}

}  // namespace orc
