// B200 styles behind the reference's plugin surface: same class roles, member names and argument meaning as
//   PairReaxCSunway   (/root/reference/pair_reaxc_sunway.{h,cpp})      -> PairReaxCB200      pair_style reax/c
//   FixQEqReaxSunway  (/root/reference/fix_qeq_reax_sunway.{h,cpp})    -> FixQEqReaxB200     fix qeq/reax
//   FixNVESunway      (/root/reference/fix_nve_sunway.{h,cpp})         -> FixNVEB200         fix nve
// Each hook validates its arguments exactly like the reference and then makes ONE call into the C ABI (include/rxb200.h).
#pragma once
#include "../../include/rxb200.h"
#include "mini_lammps.h"

namespace LAMMPS_MINI {

// keeps one host array page-locked (rxb_host_register) across reallocations
struct PinnedRegion {
  void* p = nullptr;
  size_t bytes = 0;
  void ensure(void* np, size_t nbytes) {
    if (np == p && nbytes <= bytes) return;
    release();
    if (np && nbytes && rxb_host_register(np, nbytes) == 0) { p = np; bytes = nbytes; }
  }
  void release() { if (p) rxb_host_unregister(p); p = nullptr; bytes = 0; }
  ~PinnedRegion() { release(); }
};

class PairReaxCB200 : public Pair {
 public:
  explicit PairReaxCB200(LAMMPS* lmp);
  ~PairReaxCB200() override;
  void settings(int narg, char** arg) override;      // pair_reaxc_sunway.cpp:202-290
  void coeff(int narg, char** arg) override;         // :294-362
  void init_style() override;                        // :366-424
  void compute(int eflag, int vflag) override;       // :541-793
  void* extract(const char* str, int& dim) override; // :1106-1128
  double cutghost_request() const override { return cutmax; }
  void upload_if_needed();                            // shared with fix qeq/reax: whoever runs first this step uploads x

  rxb_handle* rxb = nullptr;
  int qeqflag = 1, lgflag = 0, enobondsflag = 1;
  double safezone = 1.2, saferzone = 1.4;
  int mincap = 50, maxfar = 1024;
  double cutmax = 0.0, bg_cut = 0.3;
  std::vector<double> chi, eta, gamma;
  std::vector<int> map;
  int device_nlocal_ = 0, device_nall_ = 0;
  long uploaded_step = -1, uploaded_build = -1;   // what the device currently holds: (ntimestep, neighbor->ncalls)
  bool device_q_newer = false;                    // the device ran QEq since atom->q was last refreshed on the host
  int fixspecies_flag = 0, fixbond_flag = 0;

 private:
  std::vector<double> fbuf_;
  PinnedRegion pin_x_, pin_f_;
  std::string control_file_;
  bool coeff_done_ = false;
};

class FixQEqReaxB200 : public Fix {
 public:
  FixQEqReaxB200(LAMMPS* lmp, int narg, char** arg); // fix_qeq_reax_sunway.cpp:76-140
  void init() override;                              // :402-436
  void setup_pre_force(int vflag) override;          // :488-499
  void pre_force(int vflag) override;                // :539-600
  int nevery = 1, matvecs = 0, matvecs_s = 0, matvecs_t = 0;
  double swa = 0.0, swb = 10.0, tolerance = 1e-6;
  std::string pertype_option = "reax/c";
  PairReaxCB200* reaxc = nullptr;
};

class FixNVEB200 : public Fix {
 public:
  FixNVEB200(LAMMPS* lmp, int narg, char** arg);     // fix_nve_sunway.cpp:31-40
  void init() override;                              // :57-65
  void initial_integrate(int vflag) override;        // fix_nve_sw64.c:43-99
  void final_integrate() override;                   // :101-170
  double dtv = 0, dtf = 0;
};

class FixReaxCBondsB200 : public Fix {               // fix ID all reax/c/bonds Nevery file
 public:
  FixReaxCBondsB200(LAMMPS* lmp, int narg, char** arg);  // fix_reaxc_bonds_sunway.cpp:45-91
  ~FixReaxCBondsB200() override;
  void init() override;                              // :121-127
  void setup(int vflag) override { end_of_step(); }  // :114-117
  void end_of_step() override;                       // :131-135  Output_ReaxC_Bonds
  PairReaxCB200* reaxc = nullptr;
  FILE* fp = nullptr;
 private:
  std::vector<int> tag_, type_, off_, nbr_;
  std::vector<double> bo_, abo_, nlp_, q_;
};

class ComputeSpecAtomB200 : public Compute {         // compute ID all SPEC/ATOM q x y z vx vy vz abo01 ... abo24
 public:
  ComputeSpecAtomB200(LAMMPS* lmp, int narg, char** arg);   // compute_spec_atom_sunway.cpp:35-128
  void init() override;
  void compute_peratom() override;                   // :142-170 (+ the pack_* members :178-500)
  std::vector<int> codes;                            // 0 q, 1-3 x y z, 4-6 vx vy vz, 10 + k: abo(k+1)
  PairReaxCB200* reaxc = nullptr;
};

class FixReaxCSpeciesB200 : public Fix {             // fix ID all reax/c/species Nevery Nrepeat Nfreq file [cutoff i j v] [element ...] [position f file]
 public:
  FixReaxCSpeciesB200(LAMMPS* lmp, int narg, char** arg);  // fix_reaxc_species_sunway.cpp:50-252
  ~FixReaxCSpeciesB200() override;
  void init() override;                              // :301-329
  void setup(int vflag) override { post_integrate(); }   // :291-297
  void post_integrate() override;                    // :419-423  Output_ReaxC_Bonds
  PairReaxCB200* reaxc = nullptr;
  FILE* fp = nullptr;
  int nrepeat = 1, nfreq = 1, ntypes = 0;
  int Nmole = 0, Nspec = 0;                          // vector_nmole, vector_nspec
  std::vector<double> BOCut;
  std::vector<std::string> eletype;
  std::vector<int> clusterID;                        // vector_atom
 private:
  bool configured_ = false;
  int nev_ = 1;
  void write_formulas(const std::vector<int>& comp);
  // `position posfreq filepos` (fix_reaxc_species_sunway.cpp:210-231, OpenPos :784-810, WritePos :814-925)
  int posflag_ = 0, posfreq_ = 0, multipos_ = 0;
  std::string filepos_;
  FILE* pos_ = nullptr;
  void write_pos(const std::vector<int>& comp);
};

}  // namespace LAMMPS_MINI
