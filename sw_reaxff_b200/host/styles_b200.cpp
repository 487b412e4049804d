// See styles_b200.h.  Error texts are the reference's (pair_reaxc_sunway.cpp, fix_qeq_reax_sunway.cpp, fix_nve_sunway.cpp).
#include "styles_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace LAMMPS_MINI {

// ---------------------------------------------------------------------------------------------------------------
PairReaxCB200::PairReaxCB200(LAMMPS* l) : Pair(l) {}

PairReaxCB200::~PairReaxCB200() {
  if (rxb) rxb_destroy(rxb);
}

void PairReaxCB200::settings(int narg, char** arg) {
  Error* error = lmp->error;
  if (narg < 1) error->all(FLERR, "Illegal pair_style command");
  control_file_ = arg[0];
  qeqflag = 1; lgflag = 0; enobondsflag = 1; mincap = 50; safezone = 1.2; saferzone = 1.4; maxfar = 1024;
  int iarg = 1;
  while (iarg < narg) {
    auto yesno = [&](int& flag) {
      if (iarg + 2 > narg) error->all(FLERR, "Illegal pair_style reax/c command");
      if (!strcmp(arg[iarg + 1], "yes")) flag = 1;
      else if (!strcmp(arg[iarg + 1], "no")) flag = 0;
      else error->all(FLERR, "Illegal pair_style reax/c command");
      iarg += 2;
    };
    if (!strcmp(arg[iarg], "checkqeq")) yesno(qeqflag);
    else if (!strcmp(arg[iarg], "enobonds")) yesno(enobondsflag);
    else if (!strcmp(arg[iarg], "lgvdw")) yesno(lgflag);
    else if (!strcmp(arg[iarg], "safezone")) {
      if (iarg + 2 > narg) error->all(FLERR, "Illegal pair_style reax/c command");
      safezone = atof(arg[iarg + 1]);
      if (safezone < 0.0) error->all(FLERR, "Illegal pair_style reax/c safezone command");
      saferzone = safezone * 1.2 + 0.2;
      iarg += 2;
    } else if (!strcmp(arg[iarg], "mincap")) {
      if (iarg + 2 > narg) error->all(FLERR, "Illegal pair_style reax/c command");
      mincap = atoi(arg[iarg + 1]);
      if (mincap < 0) error->all(FLERR, "Illegal pair_style reax/c mincap command");
      iarg += 2;
    } else if (!strcmp(arg[iarg], "maxfar")) {
      if (iarg + 2 > narg) error->all(FLERR, "Illegal pair_style reax/c command");
      maxfar = atoi(arg[iarg + 1]);   // accepted for script compatibility; rows are carved dynamically on the device
      if (maxfar < 0) error->all(FLERR, "Illegal pair_style reax/c maxfar command");
      iarg += 2;
    } else error->all(FLERR, "Illegal pair_style reax/c command");
  }
  if (!rxb && rxb_create(lmp->cuda_device, &rxb)) error->all(FLERR, rxb_last_error());
  if (rxb_pair_settings(rxb, control_file_.c_str(), lgflag, enobondsflag)) error->all(FLERR, rxb_last_error());
}

void PairReaxCB200::coeff(int nargs, char** args) {
  Error* error = lmp->error;
  Atom* atom = lmp->atom;
  if (nargs != 3 + atom->ntypes) error->all(FLERR, "Incorrect args for pair coefficients");
  if (strcmp(args[0], "*") != 0 || strcmp(args[1], "*") != 0) error->all(FLERR, "Incorrect args for pair coefficients");
  if (rxb_pair_coeff(rxb, args[2], atom->ntypes, args + 3)) {
    std::string e = rxb_last_error();
    error->all(FLERR, e);
  }
  chi.assign(atom->ntypes + 1, 0); eta.assign(atom->ntypes + 1, 0); gamma.assign(atom->ntypes + 1, 0);
  rxb_pair_extract(rxb, "chi", chi.data(), atom->ntypes);
  rxb_pair_extract(rxb, "eta", eta.data(), atom->ntypes);
  rxb_pair_extract(rxb, "gamma", gamma.data(), atom->ntypes);
  coeff_done_ = true;
}

void PairReaxCB200::init_style() {
  Error* error = lmp->error;
  Atom* atom = lmp->atom;
  if (!atom->q_flag) error->all(FLERR, "Pair style reax/c requires atom attribute q");
  bool have_qeq = false;
  for (auto& f : lmp->fixes) have_qeq = have_qeq || f->style.find("qeq/reax") != std::string::npos;
  if (!have_qeq && qeqflag == 1) error->all(FLERR, "Pair reax/c requires use of fix qeq/reax");
  if (atom->tag_enable == 0) error->all(FLERR, "Pair style reax/c requires atom IDs");
  if (lmp->force->newton_pair == 0) error->all(FLERR, "Pair style reax/c requires newton pair on");
  if (!coeff_done_) error->all(FLERR, "All pair coeffs are not set");
  // cutmax = MAX3(nonb_cut, hbond_cut, 2*bond_cut): read back from the parsed parameters
  long n = rxb_params_dump(rxb, nullptr, 0);
  std::vector<double> d(n);
  rxb_params_dump(rxb, d.data(), n);
  const int ngp = (int)d[2];
  const double* ctl = &d[3 + ngp];
  cutmax = std::max(ctl[2], std::max(ctl[4], 2 * ctl[3]));
  bg_cut = ctl[5];
  rxb_neighbor_skin(rxb, lmp->neighbor->skin);
}

void PairReaxCB200::upload_if_needed() {
  Atom* atom = lmp->atom;
  // Keyed on the timestep AND the build generation: LAMMPS::setup() of a second `run` re-does remap + borders without
  // advancing ntimestep, which gives a new ghost set / index space that must reach the device.
  if (uploaded_step == lmp->update->ntimestep && uploaded_build == lmp->neighbor->ncalls) return;
  pin_x_.ensure(atom->x.data(), atom->x.capacity() * sizeof(double));   // direct PCIe copies of atom->x
  if (uploaded_build != lmp->neighbor->ncalls) {
    // reneighbouring step: new index space (what write_reax_atoms + NPair::build + write_reax_lists do in the reference).
    // atom->q goes up with it, so it must be the device's latest solution (fix qeq/reax with nevery > 1 skips steps, and
    // the host copy is otherwise refreshed on thermo steps only): the old index space is still the host's local order.
    if (device_q_newer && uploaded_build >= 0) {
      std::vector<double> qd((size_t)device_nall_);
      if (rxb_get_charges(rxb, qd.data())) lmp->error->all(FLERR, rxb_last_error());
      // local atoms keep their order across exchange in this stand-in core; ghosts are rebuilt from their owners
      const int nl = std::min(device_nlocal_, atom->nlocal);
      for (int i = 0; i < nl; i++) atom->q[i] = qd[i];
      for (int g = 0; g < atom->nghost; g++) {
        const int o = lmp->comm->ghost_owner[g];
        if (o >= 0) atom->q[atom->nlocal + g] = atom->q[o];
      }
      device_q_newer = false;
    }
    if (rxb_set_atoms(rxb, atom->nlocal, atom->nghost, atom->x.data(), atom->type.data(), atom->tag.data(), atom->q.data(),
                      lmp->comm->ghost_owner.data()))
      lmp->error->all(FLERR, rxb_last_error());
    if (rxb_neigh_build(rxb)) lmp->error->all(FLERR, rxb_last_error());
    device_nlocal_ = atom->nlocal; device_nall_ = atom->nall();
  } else {
    if (rxb_set_positions(rxb, atom->nall(), atom->x.data())) lmp->error->all(FLERR, rxb_last_error());
  }
  uploaded_step = lmp->update->ntimestep;
  uploaded_build = lmp->neighbor->ncalls;
}

void PairReaxCB200::compute(int eflag, int vflag) {
  Atom* atom = lmp->atom;
  upload_if_needed();
  const int nall = atom->nall();
  if (fbuf_.capacity() < (size_t)3 * nall) fbuf_.reserve((size_t)3 * nall + (size_t)3 * nall / 4);
  fbuf_.resize((size_t)3 * nall);
  pin_f_.ensure(fbuf_.data(), fbuf_.capacity() * sizeof(double));
  double eng[2], vir[6];
  if (rxb_pair_compute(rxb, nall, eflag, vflag, fbuf_.data(), pvector, eng, vir)) lmp->error->all(FLERR, rxb_last_error());
  double* f = atom->f.data();
  const double* fb = fbuf_.data();
  const long n3 = 3L * nall;
#pragma omp parallel for schedule(static)
  for (long k = 0; k < n3; k++) f[k] += fb[k];
  if (eflag) { eng_vdwl = eng[0]; eng_coul = eng[1]; }
  if (vflag) for (int k = 0; k < 6; k++) virial[k] = vir[k];
}

void* PairReaxCB200::extract(const char* str, int& dim) {
  dim = 1;
  if (!strcmp(str, "chi")) return chi.data();
  if (!strcmp(str, "eta")) return eta.data();
  if (!strcmp(str, "gamma")) return gamma.data();
  if (!strcmp(str, "bg_cut")) { dim = 0; return &bg_cut; }
  return nullptr;
}

// ---------------------------------------------------------------------------------------------------------------
ComputeSpecAtomB200::ComputeSpecAtomB200(LAMMPS* l, int narg, char** arg) : Compute(l) {
  Error* error = lmp->error;
  if (narg < 4) error->all(FLERR, "Illegal compute reax/c/atom command");
  id = arg[0]; style = arg[2];
  const int nvalues = narg - 3;
  size_peratom_cols = nvalues == 1 ? 0 : nvalues;
  static const char* plain[7] = {"q", "x", "y", "z", "vx", "vy", "vz"};
  for (int iarg = 3; iarg < narg; iarg++) {
    int code = -1;
    for (int k = 0; k < 7; k++) if (!strcmp(arg[iarg], plain[k])) code = k;
    if (code < 0 && !strncmp(arg[iarg], "abo", 3) && strlen(arg[iarg]) == 5) {
      const int k = atoi(arg[iarg] + 3);
      if (k >= 1 && k <= 24 && arg[iarg][3] >= '0' && arg[iarg][3] <= '2') code = 10 + (k - 1);
    }
    if (code < 0) error->all(FLERR, "Invalid keyword in compute reax/c/atom command");
    codes.push_back(code);
  }
}

void ComputeSpecAtomB200::init() {
  reaxc = dynamic_cast<PairReaxCB200*>(lmp->pair.get());
  if (!reaxc) lmp->error->all(FLERR, "Cannot use compute SPEC/ATOM without pair_style reax/c");
}

void ComputeSpecAtomB200::compute_peratom() {
  Atom* atom = lmp->atom;
  invoked_peratom = lmp->update->ntimestep;
  const int n = atom->nlocal, nv = (int)codes.size();
  array.assign((size_t)n * nv, 0.0);
  bool want_abo = false, want_q = false;
  for (int c : codes) { want_abo = want_abo || c >= 10; want_q = want_q || c == 0; }
  std::vector<double> abo, q;
  if (want_abo) {          // tmpbo of the pair style's last force evaluation (FindBond), MAXSPECBOND = 12 columns
    abo.assign((size_t)n * 12, 0.0);
    if (rxb_spec_atom_abo(reaxc->rxb, abo.data())) lmp->error->all(FLERR, rxb_last_error());
  }
  if (want_q) {            // the device holds the current charges
    q.assign((size_t)atom->nall(), 0.0);
    if (rxb_get_charges(reaxc->rxb, q.data())) lmp->error->all(FLERR, rxb_last_error());
  }
  for (int i = 0; i < n; i++)
    for (int v = 0; v < nv; v++) {
      const int c = codes[v];
      double val = 0.0;
      if (c == 0) val = q[i];
      else if (c <= 3) val = atom->x[3 * i + (c - 1)];
      else if (c <= 6) val = atom->v[3 * i + (c - 4)];
      else if (c - 10 < 12) val = abo[(size_t)i * 12 + (c - 10)];
      // (abo13..abo24 index past the reference's own tmpbo rows - MAXSPECBOND is 12 - and are reported as 0 here)
      array[(size_t)i * nv + v] = val;
    }
}

// ---------------------------------------------------------------------------------------------------------------
FixQEqReaxB200::FixQEqReaxB200(LAMMPS* l, int narg, char** arg) : Fix(l) {
  Error* error = lmp->error;
  if (narg < 8 || narg > 9) error->all(FLERR, "Illegal fix qeq/reax command");
  id = arg[0]; style = arg[2];
  nevery = atoi(arg[3]);
  if (nevery <= 0) error->all(FLERR, "Illegal fix qeq/reax command");
  swa = atof(arg[4]); swb = atof(arg[5]); tolerance = atof(arg[6]);
  pertype_option = arg[7];          // "reax/c" or a parameter file (pertype_parameters, fix_qeq_reax_sunway.cpp:198-245)
  if (narg == 9 && strcmp(arg[8], "dual") != 0) error->all(FLERR, "Illegal fix qeq/reax command");
  // ("dual" is accepted: both solves always run fused here)
}

void FixQEqReaxB200::init() {
  Error* error = lmp->error;
  if (!lmp->atom->q_flag) error->all(FLERR, "Fix qeq/reax requires atom attribute q");
  reaxc = dynamic_cast<PairReaxCB200*>(lmp->pair.get());
  if (!reaxc) error->all(FLERR, "Fix qeq/reax: could not extract params from pair reax/c");
  if (swb < 0) error->all(FLERR, "Fix qeq/reax has negative upper Taper radius cutoff");
  if (fabs(swa) > 0.01) error->warning(FLERR, "Fix qeq/reax has non-zero lower Taper radius cutoff");
  else if (swb < 5) error->warning(FLERR, "Fix qeq/reax has very low Taper radius cutoff");
  if (rxb_fix_qeq(reaxc->rxb, swa, swb, tolerance, 200)) error->all(FLERR, rxb_last_error());
  // pertype_parameters(): "reax/c" takes chi/eta/gamma from the pair style, anything else names a file with one
  // `itype chi eta gamma` line per atom type
  if (pertype_option == "reax/c") {
    if (rxb_fix_qeq_params(reaxc->rxb, 0, nullptr, nullptr, nullptr)) error->all(FLERR, rxb_last_error());
  } else {
    const int ntypes = lmp->atom->ntypes;
    std::vector<double> chi(ntypes + 1, 0.0), eta(ntypes + 1, 0.0), gamma(ntypes + 1, 0.0);
    FILE* pf = fopen(pertype_option.c_str(), "r");
    if (!pf) error->one(FLERR, "Fix qeq/reax parameter file could not be found");
    int i;
    for (i = 1; i <= ntypes && !feof(pf); i++) {
      int itype = 0;
      double v1 = 0, v2 = 0, v3 = 0;
      if (fscanf(pf, "%d %lg %lg %lg", &itype, &v1, &v2, &v3) != 4) break;
      if (itype < 1 || itype > ntypes) { fclose(pf); error->one(FLERR, "Fix qeq/reax invalid atom type in param file"); }
      chi[itype] = v1; eta[itype] = v2; gamma[itype] = v3;
    }
    fclose(pf);
    if (i <= ntypes) error->one(FLERR, "Invalid param file for fix qeq/reax");
    if (rxb_fix_qeq_params(reaxc->rxb, ntypes, chi.data(), eta.data(), gamma.data())) error->all(FLERR, rxb_last_error());
  }
}

void FixQEqReaxB200::setup_pre_force(int vflag) { pre_force(vflag); }

void FixQEqReaxB200::pre_force(int) {
  if (lmp->update->ntimestep % nevery) return;
  reaxc->upload_if_needed();
  // The iteration counts (and a host copy of q) are only consumed on thermo steps: there the call waits for convergence
  // and reports matvecs like the reference; on the other steps the solve is enqueued without a host round trip and is
  // settled with the end-of-step status inside rxb_pair_compute (include/rxb200.h).
  const bool thermo_step = !lmp->thermo_every || lmp->update->ntimestep % lmp->thermo_every == 0;
  int mv[2] = {matvecs_s, matvecs_t};
  if (rxb_qeq_pre_force(reaxc->rxb, thermo_step ? mv : nullptr)) lmp->error->all(FLERR, rxb_last_error());
  matvecs_s = mv[0]; matvecs_t = mv[1]; matvecs = mv[0] + mv[1];
  if (mv[0] >= 200 || mv[1] >= 200)
    lmp->error->warning(FLERR, "Fix qeq/reax CG convergence failed after 200 iterations at " + std::to_string(lmp->update->ntimestep) + " step");
  // atom->q on the host is only needed by host-side consumers (thermo/dump): refresh it on thermo steps
  reaxc->device_q_newer = true;
  if (lmp->thermo_every && lmp->update->ntimestep % lmp->thermo_every == 0) {
    rxb_get_charges(reaxc->rxb, lmp->atom->q.data());
    reaxc->device_q_newer = false;
  }
}

// ---------------------------------------------------------------------------------------------------------------
FixReaxCBondsB200::FixReaxCBondsB200(LAMMPS* l, int narg, char** arg) : Fix(l) {
  Error* error = lmp->error;
  if (narg != 5) error->all(FLERR, "Illegal fix reax/c/bonds command");
  id = arg[0]; style = arg[2];
  nevery = atoi(arg[3]);
  if (nevery <= 0) error->all(FLERR, "Illegal fix reax/c/bonds command");
  const char* suffix = strrchr(arg[4], '.');
  if (suffix && strcmp(suffix, ".gz") == 0) error->all(FLERR, "Cannot open gzipped file");
  fp = fopen(arg[4], "w");
  if (!fp) error->all(FLERR, std::string("Cannot open fix reax/c/bonds file ") + arg[4]);
  for (int i = 0; i < lmp->atom->nlocal; i++)   // tag_consecutive()
    if (lmp->atom->tag[i] < 1 || lmp->atom->tag[i] > lmp->atom->natoms)
      error->all(FLERR, "Atom IDs must be consecutive for fix reax/c bonds");
}

FixReaxCBondsB200::~FixReaxCBondsB200() { if (fp) fclose(fp); }

void FixReaxCBondsB200::init() {
  reaxc = dynamic_cast<PairReaxCB200*>(lmp->pair.get());
  if (!reaxc) lmp->error->all(FLERR, "Cannot use fix reax/c/bonds without pair_style reax/c, reax/c/kk, or reax/c/omp");
}

void FixReaxCBondsB200::end_of_step() {
  Error* error = lmp->error;
  int n = 0, m = 0, maxnum = 0;
  if (rxb_bond_table(reaxc->rxb, -1.0, &n, &m, &maxnum)) error->all(FLERR, rxb_last_error());
  tag_.resize(n); type_.resize(n); off_.resize(n + 1); nbr_.resize(m > 0 ? m : 1); bo_.resize(m > 0 ? m : 1);
  abo_.resize(n); nlp_.resize(n); q_.resize(n);
  if (rxb_bond_table_get(reaxc->rxb, tag_.data(), type_.data(), off_.data(), nbr_.data(), bo_.data(), abo_.data(),
                         nlp_.data(), q_.data()))
    error->all(FLERR, rxb_last_error());
  int dim;
  double bg_cut = 0.3;
  if (double* c = (double*)reaxc->extract("bg_cut", dim)) bg_cut = *c;
  // RecvBuffer, fix_reaxc_bonds_sunway.cpp:264-330 (one rank)
  fprintf(fp, "# Timestep %ld \n", lmp->update->ntimestep);
  fprintf(fp, "# \n");
  fprintf(fp, "# Number of particles %d \n", (int)lmp->atom->natoms);
  fprintf(fp, "# \n");
  fprintf(fp, "# Max number of bonds per atom %d with coarse bond order cutoff %5.3f \n", maxnum, bg_cut);
  fprintf(fp, "# Particle connection table and bond orders \n");
  fprintf(fp, "# id type nb id_1...id_nb mol bo_1...bo_nb abo nlp q \n");
  for (int i = 0; i < n; i++) {
    const int a = off_[i], b = off_[i + 1];
    fprintf(fp, " %d %d %d", tag_[i], type_[i], b - a);
    for (int k = a; k < b; k++) fprintf(fp, " %d", nbr_[k]);
    fprintf(fp, " %d", 0);
    for (int k = a; k < b; k++) fprintf(fp, "%14.3f", bo_[k]);
    fprintf(fp, "%14.3f%14.3f%14.3f\n", abo_[i], nlp_[i], q_[i]);
  }
  fprintf(fp, "# \n");
  fflush(fp);
}

// ---------------------------------------------------------------------------------------------------------------
FixReaxCSpeciesB200::FixReaxCSpeciesB200(LAMMPS* l, int narg, char** arg) : Fix(l) {
  Error* error = lmp->error;
  if (narg < 7) error->all(FLERR, "Illegal fix reax/c/species command");
  id = arg[0]; style = arg[2];
  ntypes = lmp->atom->ntypes;
  const int nev = atoi(arg[3]);
  nrepeat = atoi(arg[4]);
  nfreq = atoi(arg[5]);
  if (nev <= 0 || nrepeat <= 0 || nfreq <= 0) error->all(FLERR, "Illegal fix reax/c/species command");
  if (nfreq % nev || nrepeat * nev > nfreq) error->all(FLERR, "Illegal fix reax/c/species command");
  nevery = 0;            // POST_INTEGRATE only; the sampling period lives in nev_
  // neighbour lists must stay unchanged while bond orders are averaged (fix_reaxc_species_sunway.cpp:81-108)
  Neighbor* neighbor = lmp->neighbor;
  int rene_flag = 0;
  if (nev * nrepeat != 1 && (nfreq % neighbor->every != 0 || neighbor->every < nev * nrepeat)) {
    int newevery = nev * nrepeat;
    while (nfreq % newevery != 0 && newevery <= nfreq / 2) newevery++;
    if (nfreq % newevery != 0) newevery = nfreq;
    neighbor->every = newevery;
    rene_flag = 1;
  }
  if (nev * nrepeat != 1 && (neighbor->delay != 0 || neighbor->dist_check != 0)) {
    neighbor->delay = 0;
    neighbor->dist_check = 0;
    rene_flag = 1;
  }
  if (rene_flag) error->warning(FLERR, "Resetting reneighboring criteria for fix reax/c/species");
  const char* suffix = strrchr(arg[6], '.');
  if (suffix && strcmp(suffix, ".gz") == 0) error->all(FLERR, "Cannot open gzipped file");
  fp = fopen(arg[6], "w");
  if (!fp) error->all(FLERR, std::string("Cannot open fix reax/c/species file ") + arg[6]);
  const int n = ntypes + 1;
  BOCut.assign((size_t)n * n, 0.30);
  int iarg = 7;
  while (iarg < narg) {
    if (strcmp(arg[iarg], "cutoff") == 0) {
      if (iarg + 4 > narg) error->all(FLERR, "Illegal fix reax/c/species command");
      const int itype = atoi(arg[iarg + 1]), jtype = atoi(arg[iarg + 2]);
      const double bo_cut = atof(arg[iarg + 3]);
      if (itype > ntypes || jtype > ntypes) error->all(FLERR, "Illegal fix reax/c/species command");
      if (itype <= 0 || jtype <= 0) error->all(FLERR, "Illegal fix reax/c/species command");
      if (bo_cut > 1.0 || bo_cut < 0.0) error->all(FLERR, "Illegal fix reax/c/species command");
      BOCut[(size_t)itype * n + jtype] = bo_cut;
      BOCut[(size_t)jtype * n + itype] = bo_cut;
      iarg += 4;
    } else if (strcmp(arg[iarg], "element") == 0) {
      if (iarg + ntypes + 1 > narg) error->all(FLERR, "Illegal fix reax/c/species command");
      for (int i = 0; i < ntypes; i++) eletype.push_back(arg[iarg + 1 + i]);
      iarg += ntypes + 1;
    } else if (strcmp(arg[iarg], "position") == 0) {
      if (iarg + 3 > narg) error->all(FLERR, "Illegal fix reax/c/species command");
      posflag_ = 1;
      posfreq_ = atoi(arg[iarg + 1]);
      if (posfreq_ < nfreq || (posfreq_ % nfreq != 0)) error->all(FLERR, "Illegal fix reax/c/species command");
      filepos_ = arg[iarg + 2];
      if (filepos_.find('*') != std::string::npos) multipos_ = 1;
      else {
        pos_ = fopen(filepos_.c_str(), "w");
        if (!pos_) error->one(FLERR, "Cannot open fix reax/c/species position file");
      }
      iarg += 3;
    } else error->all(FLERR, "Illegal fix reax/c/species command");
  }
  if (eletype.empty()) { const char* d[4] = {"C", "H", "O", "N"}; for (int i = 0; i < ntypes && i < 4; i++) eletype.push_back(d[i]); }
  nev_ = nev;
}

FixReaxCSpeciesB200::~FixReaxCSpeciesB200() {
  if (fp) fclose(fp);
  if (pos_) fclose(pos_);
}

void FixReaxCSpeciesB200::init() {
  Error* error = lmp->error;
  if (lmp->atom->tag_enable == 0) error->all(FLERR, "Cannot use fix reax/c/species unless atoms have IDs");
  reaxc = dynamic_cast<PairReaxCB200*>(lmp->pair.get());
  if (!reaxc) error->all(FLERR, "Cannot use fix reax/c/species without pair_style reax/c, reax/c/kk, or reax/c/omp");
  reaxc->fixspecies_flag = 1;
  int count = 0;
  for (auto& f : lmp->fixes) if (f->style == "reax/c/species" || f->style == "reax/c/species/b200") count++;
  if (count > 1) error->warning(FLERR, "More than one fix reax/c/species");
  configured_ = false;   // the handle exists only after pair init_style: configure at the first hook
}

void FixReaxCSpeciesB200::post_integrate() {
  Error* error = lmp->error;
  if (!configured_) {
    int reset = 0;
    if (rxb_species_config(reaxc->rxb, nev_, nrepeat, nfreq, ntypes, BOCut.data(), lmp->atom->natoms,
                           lmp->update->ntimestep, &reset))
      error->all(FLERR, rxb_last_error());
    configured_ = true;
  }
  int found = 0;
  if (rxb_species_step(reaxc->rxb, lmp->update->ntimestep, &found)) error->all(FLERR, rxb_last_error());
  if (!found) return;
  if (rxb_species_result(reaxc->rxb, &Nmole, nullptr, 0)) error->all(FLERR, rxb_last_error());
  std::vector<int> comp((size_t)Nmole * ntypes);
  if (rxb_species_result(reaxc->rxb, &Nmole, comp.data(), (long)comp.size())) error->all(FLERR, rxb_last_error());
  clusterID.resize(lmp->atom->nlocal);
  if (rxb_species_cluster(reaxc->rxb, clusterID.data())) error->all(FLERR, rxb_last_error());
  write_formulas(comp);
  fflush(fp);
  if (posflag_ && lmp->update->ntimestep % posfreq_ == 0) write_pos(comp);
}

// WritePos, fix_reaxc_species_sunway.cpp:814-925 (one rank).  The anchor x0 of a molecule - the fixed point of FindMolecule's
// chAnchor propagation (:497-510, :530-532, :559-569) - is the lexicographically smallest averaged position of its atoms.
void FixReaxCSpeciesB200::write_pos(const std::vector<int>& comp) {
  Error* error = lmp->error;
  const int nlocal = lmp->atom->nlocal;
  std::vector<double> col((size_t)4 * std::max(nlocal, 1));
  if (rxb_species_avg_qxyz(reaxc->rxb, col.data())) error->all(FLERR, rxb_last_error());
  const long ntimestep = lmp->update->ntimestep;
  if (multipos_) {                        // OpenPos: '*' -> timestep
    if (pos_) fclose(pos_);
    const size_t star = filepos_.find('*');
    const std::string name = filepos_.substr(0, star) + std::to_string(ntimestep) + filepos_.substr(star + 1);
    pos_ = fopen(name.c_str(), "w");
    if (!pos_) error->one(FLERR, "Cannot open fix reax/c/species position file");
  }
  const Domain* d = lmp->domain;
  const double lo[3] = {d->boxlo[0], d->boxlo[1], d->boxlo[2]};
  const double hi[3] = {lo[0] + d->h[0], lo[1] + d->h[1], lo[2] + d->h[2]};
  const double box[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
  const double halfbox[3] = {box[0] / 2, box[1] / 2, box[2] / 2};
  std::vector<double> anchor((size_t)3 * std::max(Nmole, 1), 0.0);
  std::vector<char> have(std::max(Nmole, 1), 0);
  for (int i = 0; i < nlocal; i++) {
    const int m = clusterID[i] - 1;
    const double* xi = &col[4 * (size_t)i + 1];
    double* a = &anchor[3 * (size_t)m];
    const bool less = !have[m] || xi[0] < a[0] || (xi[0] == a[0] && (xi[1] < a[1] || (xi[1] == a[1] && xi[2] < a[2])));
    if (less) { a[0] = xi[0]; a[1] = xi[1]; a[2] = xi[2]; have[m] = 1; }
  }
  fprintf(pos_, "Timestep %ld NMole %d  NSpec %d  xlo %f  xhi %f  ylo %f  yhi %f  zlo %f  zhi %f\n", ntimestep, Nmole, Nspec, lo[0],
          hi[0], lo[1], hi[1], lo[2], hi[2]);
  fprintf(pos_, "ID\tAtom_Count\tType\tAve_q\t\tCoM_x\t\tCoM_y\t\tCoM_z\n");
  // members of every molecule in ascending local index (the reference scans all local atoms once per molecule)
  std::vector<int> mstart(Nmole + 1, 0), member(nlocal);
  for (int i = 0; i < nlocal; i++) mstart[clusterID[i]]++;
  for (int m = 0; m < Nmole; m++) mstart[m + 1] += mstart[m];
  { std::vector<int> cur(mstart.begin(), mstart.end() - 1); for (int i = 0; i < nlocal; i++) member[cur[clusterID[i] - 1]++] = i; }
  for (int m = 1; m <= Nmole; m++) {
    const int* Name = &comp[(size_t)(m - 1) * ntypes];
    int count = 0;
    double avq = 0.0, avx[3] = {0, 0, 0};
    const double* x0 = &anchor[3 * (size_t)(m - 1)];
    for (int k = mstart[m - 1]; k < mstart[m]; k++) {
      double* sa = &col[4 * (size_t)member[k]];
      count++;
      avq += sa[0];
      for (int t = 0; t < 3; t++) {
        if ((x0[t] - sa[1 + t]) > halfbox[t]) sa[1 + t] += box[t];
        if ((sa[1 + t] - x0[t]) > halfbox[t]) sa[1 + t] -= box[t];
      }
      for (int t = 0; t < 3; t++) avx[t] += sa[1 + t];
    }
    fprintf(pos_, "%d\t%d\t", m, count);
    for (int n = 0; n < ntypes; n++)
      if (Name[n] != 0) {
        fprintf(pos_, "%s", eletype[n].c_str());
        if (Name[n] != 1) fprintf(pos_, "%d", Name[n]);
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
      fprintf(pos_, "\t%.8f \t%.8f \t%.8f \t%.8f", avq, avx[0], avx[1], avx[2]);
    }
    fprintf(pos_, "\n");
  }
  if (!multipos_) fprintf(pos_, "#\n");
  fflush(pos_);
}

// FindSpecies + WriteFormulas, fix_reaxc_species_sunway.cpp:652-717, 745-780
void FixReaxCSpeciesB200::write_formulas(const std::vector<int>& comp) {
  std::vector<int> MolName, NMol;
  Nspec = 0;
  for (int m = 0; m < Nmole; m++) {
    const int* Name = &comp[(size_t)m * ntypes];
    int k = 0;
    for (; k < Nspec; k++) if (std::equal(Name, Name + ntypes, &MolName[(size_t)k * ntypes])) break;
    if (k < Nspec) NMol[k]++;
    else { MolName.insert(MolName.end(), Name, Name + ntypes); NMol.push_back(1); Nspec++; }
  }
  fprintf(fp, "# Timestep     No_Moles     No_Specs     ");
  for (int i = 0; i < Nspec; i++) {
    for (int j = 0; j < ntypes; j++) {
      const int itemp = MolName[(size_t)ntypes * i + j];
      if (itemp != 0) {
        fprintf(fp, "%s", eletype[j].c_str());
        if (itemp != 1) fprintf(fp, "%d", itemp);
      }
    }
    fprintf(fp, "\t");
  }
  fprintf(fp, "\n");
  fprintf(fp, "%ld", lmp->update->ntimestep);
  fprintf(fp, "%11d%11d\t", Nmole, Nspec);
  for (int i = 0; i < Nspec; i++) fprintf(fp, " %d\t", NMol[i]);
  fprintf(fp, "\n");
}

// ---------------------------------------------------------------------------------------------------------------
FixNVEB200::FixNVEB200(LAMMPS* l, int narg, char** arg) : Fix(l) {
  if (narg < 3) lmp->error->all(FLERR, "Illegal fix nve command");
  id = arg[0]; style = arg[2];
}

void FixNVEB200::init() {
  dtv = lmp->update->dt;
  dtf = 0.5 * lmp->update->dt * lmp->force->ftm2v;
}

void FixNVEB200::initial_integrate(int) {
  Atom* a = lmp->atom;
  double* x = a->x.data(); double* v = a->v.data(); const double* f = a->f.data();
#pragma omp parallel for schedule(static)
  for (int i = 0; i < a->nlocal; i++) {
    const double dtfm = dtf / a->mass[a->type[i]];
    for (int t = 0; t < 3; t++) {
      v[3 * i + t] += dtfm * f[3 * i + t];
      x[3 * i + t] += dtv * v[3 * i + t];
    }
  }
}

void FixNVEB200::final_integrate() {
  Atom* a = lmp->atom;
  double* v = a->v.data(); const double* f = a->f.data();
#pragma omp parallel for schedule(static)
  for (int i = 0; i < a->nlocal; i++) {
    const double dtfm = dtf / a->mass[a->type[i]];
    for (int t = 0; t < 3; t++) v[3 * i + t] += dtfm * f[3 * i + t];
  }
}

}  // namespace LAMMPS_MINI
