// mini_lammps.h — single-rank stand-in for the LAMMPS core that the reference's styles plug into.
//
// The LAMMPS core is ABSENT from /root/reference and from this container, so the drop-in claim is
// demonstrated against this stand-in: the style classes in styles_b200.{h,cpp} are written against the same member
// names and call order as the reference's PairReaxCSunway / FixQEqReaxSunway / FixNVESunway (atom->x, atom->nlocal,
// neighbor->ago, comm->forward_comm(), force->ftm2v, update->dt, error->all ...), and the input parser accepts the
// reference's own input script (in.reaxc.lattice: variable / units / atom_style / lattice custom / region prism /
// create_box / create_atoms ... basis / mass / pair_style / pair_coeff / neighbor / neigh_modify / fix / thermo /
// timestep / run) as well as read_data + replicate.  Host arrays are LAMMPS-layout (x[nall][3] doubles), the GPU
// sits behind the C ABI of include/rxb200.h.
#pragma once
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace LAMMPS_MINI {

struct LAMMPS;

struct Error {
  [[noreturn]] void all(const char* file, int line, const std::string& msg) const {
    throw std::runtime_error("ERROR: " + msg + " (" + file + ":" + std::to_string(line) + ")");
  }
  [[noreturn]] void one(const char* file, int line, const std::string& msg) const { all(file, line, msg); }   // single rank: same thing
  void warning(const char*, int, const std::string& msg) const { fprintf(stderr, "WARNING: %s\n", msg.c_str()); }
};
#define FLERR __FILE__, __LINE__

struct Atom {
  int nlocal = 0, nghost = 0, ntypes = 0;
  long natoms = 0;
  std::vector<double> x, v, f, q;   // [nall][3], [nlocal][3], [nall][3], [nall]
  std::vector<int> type, tag;       // [nall]
  std::vector<double> mass;         // [ntypes+1]
  int q_flag = 1, tag_enable = 1;
  int nall() const { return nlocal + nghost; }
};

struct Domain {
  double boxlo[3] = {0, 0, 0};
  double h[6] = {1, 1, 1, 0, 0, 0}, h_inv[6];   // xprd yprd zprd yz xz xy
  int triclinic = 1;
  void set_box(double xprd, double yprd, double zprd, double xy, double xz, double yz);
  void x2lamda(const double* x, double* l) const;
  void image_shift(int sx, int sy, int sz, double* d) const;
  void remap(Atom& a) const;                     // pbc()
};

struct Comm {   // single rank: ghosts are periodic images
  double cutghostuser = 0.0;
  std::vector<int> ghost_owner;
  std::vector<double> ghost_shift;               // [nghost][3] cartesian
  std::vector<size_t> block_off;                 // ghost range of every image shift (distinct owners inside a block)
  void borders(const Domain& d, Atom& a, double cutghost);   // (re)creates the ghost atoms
  void forward_comm(Atom& a) const;              // x of ghosts
  void reverse_comm(Atom& a) const;              // f of ghosts -> owners
  void forward_comm_q(Atom& a) const;
};

struct Neighbor {
  double skin = 2.0;
  int every = 1, delay = 10, dist_check = 1, ago = 0;
  long ncalls = 0;   // number of borders/neighbour builds so far (LAMMPS Neighbor::ncalls): a new index space each time
  int decide() { ago++; return (ago >= delay && ago % every == 0) ? 1 : 0; }   // 'check no' semantics
};

struct Update { long ntimestep = 0; double dt = 1.0; };
struct Force { double ftm2v = 1.0 / 48.88821291 / 48.88821291, mvv2e = 48.88821291 * 48.88821291, boltz = 0.0019872067; int newton_pair = 1; };

struct Fix {
  LAMMPS* lmp;
  std::string id, style;
  explicit Fix(LAMMPS* l) : lmp(l) {}
  virtual ~Fix() {}
  virtual void init() {}
  virtual void setup_pre_force(int) {}
  virtual void setup(int) {}
  virtual void initial_integrate(int) {}
  virtual void post_integrate() {}
  virtual void end_of_step() {}
  int nevery = 0;                                 // end_of_step() runs when ntimestep % nevery == 0 (0: never)
  virtual void pre_force(int) {}
  virtual void final_integrate() {}
};

struct Compute {                                  // per-atom computes only (what compute SPEC/ATOM is)
  LAMMPS* lmp;
  std::string id, style;
  int size_peratom_cols = 0;                      // 0: vector_atom, else array_atom with this many columns
  std::vector<double> array;                      // [nlocal][max(size_peratom_cols, 1)]
  long invoked_peratom = -1;
  explicit Compute(LAMMPS* l) : lmp(l) {}
  virtual ~Compute() {}
  virtual void init() {}
  virtual void compute_peratom() = 0;
};

struct Pair {
  LAMMPS* lmp;
  double eng_vdwl = 0, eng_coul = 0, virial[6] = {0, 0, 0, 0, 0, 0};
  double pvector[14] = {0};
  int nextra = 14;
  explicit Pair(LAMMPS* l) : lmp(l) {}
  virtual ~Pair() {}
  virtual void settings(int, char**) = 0;
  virtual void coeff(int, char**) = 0;
  virtual void init_style() = 0;
  virtual void compute(int eflag, int vflag) = 0;
  virtual void* extract(const char*, int&) { return nullptr; }
  virtual double cutghost_request() const = 0;
};

struct Thermo { long step; double temp, pe, ke, etotal; double pvector[14]; };

struct LAMMPS {
  Atom atom_, *atom = &atom_;
  Domain domain_, *domain = &domain_;
  Comm comm_, *comm = &comm_;
  Neighbor neighbor_, *neighbor = &neighbor_;
  Update update_, *update = &update_;
  Force force_, *force = &force_;
  Error error_, *error = &error_;
  std::unique_ptr<Pair> pair;
  std::vector<std::unique_ptr<Fix>> fixes;
  std::vector<std::unique_ptr<Compute>> computes;  // evaluated on thermo output steps (stand-in for dump / fix ave/atom consumers)
  int thermo_every = 0;
  int cuda_device = 0;
  bool echo_thermo = true;
  std::vector<Thermo> thermo_log;
  // wall seconds per phase of iterate(): integrate, borders, forward_comm, zero f, pre_force (upload + QEq), pair compute
  // (+ force download), reverse_comm, final integrate + output
  double phase_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::map<std::string, std::string> vars;

  void file(const std::string& path);            // execute an input script
  void one(const std::string& line);             // execute one command
  void run(long nsteps);                         // Verlet::setup + Verlet::run
  void iterate(long nsteps);                     // Verlet::run only (after a setup)
  void setup();
  double kinetic() const;

 private:
  std::string substitute(const std::string& s) const;
  void thermo_line(int eflag_done);
  bool setup_done_ = false;
  // lattice state
  double a1[3] = {1, 0, 0}, a2[3] = {0, 1, 0}, a3[3] = {0, 0, 1};
  std::vector<double> basis;                     // fractional, [nbasis][3]
  int box_ntypes = 0;
};

double evaluate(const std::string& expr);        // + - * / and parentheses on numbers

}  // namespace LAMMPS_MINI
