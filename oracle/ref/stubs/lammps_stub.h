// Stand-in for the parts of the LAMMPS core that fix_qeq_reax_sunway.cpp touches (written for this repo; the core is
// absent from /root/reference).  Single process, ghosts are periodic images whose owner is a local atom.  Only what the
// reference source needs to COMPILE UNMODIFIED and run its own CG / init_matvec / calculate_Q code is provided.
#pragma once
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "lmptype.h"

#define FLERR __FILE__, __LINE__

namespace LAMMPS_NS {

class Fix;
class Pair;
class LAMMPS;

class NeighList {
 public:
  int inum = 0, gnum = 0;
  int *ilist = nullptr, *numneigh = nullptr;
  int** firstneigh = nullptr;
};

class NeighRequest {
 public:
  int pair, fix, half, full, newton, ghost;
};

class Neighbor {
 public:
  NeighRequest** requests = nullptr;
  int every = 1, delay = 10, dist_check = 1;
  int request(void*, int = 0) { return 0; }
};

class Error {
 public:
  void all(const char* file, int line, const char* msg) { throw std::runtime_error(std::string("ERROR: ") + msg); }
  void one(const char* file, int line, const char* msg) { throw std::runtime_error(std::string("ERROR on proc 0: ") + msg); }
  void warning(const char*, int, const char* msg, int = 1) { warnings.push_back(msg); }
  std::vector<std::string> warnings;
};

class Memory {
 public:
  template <typename T> T* create(T*& a, int n, const char*) { a = (T*)calloc(n > 0 ? n : 1, sizeof(T)); return a; }
  template <typename T> T** create(T**& a, int n1, int n2, const char*) {
    T* data = (T*)calloc((size_t)(n1 > 0 ? n1 : 1) * n2, sizeof(T));
    a = (T**)calloc(n1 > 0 ? n1 : 1, sizeof(T*));
    for (int i = 0; i < n1; i++) a[i] = data + (size_t)i * n2;
    return a;
  }
  template <typename T> T** grow(T**& a, int n1, int n2, const char* name) {
    if (a == nullptr) return create(a, n1, n2, name);
    T* data = (T*)realloc(a[0], (size_t)n1 * n2 * sizeof(T));
    a = (T**)realloc(a, n1 * sizeof(T*));
    for (int i = 0; i < n1; i++) a[i] = data + (size_t)i * n2;
    return a;
  }
  template <typename T> void destroy(T*& a) { free(a); a = nullptr; }
  template <typename T> void destroy(T**& a) { if (a) { free(a[0]); free(a); } a = nullptr; }
};

class Atom {
 public:
  double** x = nullptr;
  int *type = nullptr, *mask = nullptr;
  tagint* tag = nullptr;
  double* q = nullptr;
  int nlocal = 0, nghost = 0, nmax = 0, ntypes = 0, q_flag = 1, tag_enable = 1;
  bigint natoms = 0;
  tagint* molecule = nullptr;          // atom_style charge: no molecule IDs
  double** v = nullptr;
  int tag_consecutive() { return 1; }
  void add_callback(int) {}
  void delete_callback(const char*, int) {}
};

class Comm {
 public:
  int me = 0, nprocs = 1;
  Atom* atom = nullptr;
  std::vector<int> ghost_owner;              // periodic images: owner index of every ghost
  int forward_calls = 0;
  void forward_comm_fix(Fix* fix, int size = 0);
  void reverse_comm_fix(Fix*, int = 0) {}
};

class Force {
 public:
  Pair* pair = nullptr;
  double numeric(const char*, int, char* s) { return atof(s); }
  int inumeric(const char*, int, char* s) { return atoi(s); }
  Pair* pair_match(const char*, int) { return pair; }
};

class Respa {
 public:
  int nlevels = 1;
};

class Update {
 public:
  bigint ntimestep = 0;
  char integrate_style_[16];
  char* integrate_style = integrate_style_;
  void* integrate = nullptr;
  int whichflag = 1;
  Update() { strcpy(integrate_style_, "verlet"); }
};

class Group {
 public:
  bigint natoms = 0;
  bigint count(int) { return natoms; }
};

class CiteMe {
 public:
  void add(const char*) {}
};

class Domain {
 public:
  double boxlo[3] = {0, 0, 0}, boxhi[3] = {0, 0, 0};
};
class Compute;
class Modify {
 public:
  int nfix = 0, ncompute = 0;
  Fix** fix = nullptr;
  Compute** compute = nullptr;
  void add_compute(int, char**, int = 1) {}
  void add_fix(int, char**, int = 1) {}
  void delete_compute(const char*) {}
  void delete_fix(const char*) {}
  int find_compute(const char*) { return -1; }
  int find_fix(const char*) { return -1; }
};

class LAMMPS {
 public:
  Atom* atom; Comm* comm; Memory* memory; Error* error; Force* force; Neighbor* neighbor; Update* update; Group* group;
  CiteMe* citeme; Domain* domain; Modify* modify;
  MPI_Comm world = 0;
};

class Pointers {
 public:
  explicit Pointers(LAMMPS* l)
      : lmp(l), memory(l->memory), error(l->error), atom(l->atom), comm(l->comm), force(l->force), neighbor(l->neighbor),
        update(l->update), group(l->group), domain(l->domain), modify(l->modify), world(l->world) {}
  virtual ~Pointers() {}
 protected:
  LAMMPS* lmp;
  Memory* memory; Error* error; Atom* atom; Comm* comm; Force* force; Neighbor* neighbor; Update* update; Group* group;
  Domain* domain; Modify* modify;
  MPI_Comm world;
};

namespace FixConst {
static const int INITIAL_INTEGRATE = 1 << 0, POST_INTEGRATE = 1 << 1, PRE_EXCHANGE = 1 << 2, PRE_NEIGHBOR = 1 << 3,
                 PRE_FORCE = 1 << 4, POST_FORCE = 1 << 5, FINAL_INTEGRATE = 1 << 6, END_OF_STEP = 1 << 7,
                 THERMO_ENERGY = 1 << 8, INITIAL_INTEGRATE_RESPA = 1 << 9, POST_INTEGRATE_RESPA = 1 << 10,
                 PRE_FORCE_RESPA = 1 << 11, POST_FORCE_RESPA = 1 << 12, FINAL_INTEGRATE_RESPA = 1 << 13,
                 MIN_PRE_EXCHANGE = 1 << 14, MIN_PRE_NEIGHBOR = 1 << 15, MIN_PRE_FORCE = 1 << 16, MIN_POST_FORCE = 1 << 17,
                 MIN_ENERGY = 1 << 18, POST_RUN = 1 << 19;
}

class Fix : protected Pointers {
 public:
  char *id = nullptr, *style = nullptr;
  int igroup = 0, groupbit = 1;
  int comm_forward = 0, comm_reverse = 0;
  int instance_me = 0;
  int copymode = 0;
  Fix(LAMMPS* l, int narg, char** arg) : Pointers(l) {
    if (narg > 0) id = arg[0];
    if (narg > 2) style = arg[2];
  }
  virtual ~Fix() {}
  int nevery = 1, peratom_flag = 0, size_peratom_cols = 0, peratom_freq = 1, global_freq = 1, vector_flag = 0, size_vector = 0,
      extvector = 0, restart_global = 0, time_integrate = 0, box_change = 0;
  int force_reneighbor = 0, next_reneighbor = 0;
  double* vector_atom = nullptr;
  double** array_atom = nullptr;
  virtual int setmask() = 0;
  virtual void setup(int) {}
  virtual void end_of_step() {}
  virtual void post_integrate() {}
  virtual double compute_vector(int) { return 0.0; }
  virtual void post_constructor() {}
  virtual void init() {}
  virtual void init_list(int, NeighList*) {}
  virtual void setup_pre_force(int) {}
  virtual void pre_force(int) {}
  virtual void setup_pre_force_respa(int, int) {}
  virtual void pre_force_respa(int, int, int) {}
  virtual void min_setup_pre_force(int) {}
  virtual void min_pre_force(int) {}
  virtual int pack_forward_comm(int, int*, double*, int, int*) { return 0; }
  virtual void unpack_forward_comm(int, int, double*) {}
  virtual int pack_reverse_comm(int, int, double*) { return 0; }
  virtual void unpack_reverse_comm(int, int*, double*) {}
  virtual double memory_usage() { return 0.0; }
  virtual void grow_arrays(int) {}
  virtual void copy_arrays(int, int, int) {}
  virtual int pack_exchange(int, double*) { return 0; }
  virtual int unpack_exchange(int, double*) { return 0; }
};

// fix ave/atom (LAMMPS core): only what fix reax/c/species reads — the averaged per-atom array and the end_of_step hook
class FixAveAtom : public Fix {
 public:
  FixAveAtom(LAMMPS* l, int narg, char** arg) : Fix(l, narg, arg) {}
  int setmask() { return FixConst::END_OF_STEP; }
};

class Compute : protected Pointers {
 public:
  explicit Compute(LAMMPS* l) : Pointers(l) {}
};

// one swap: every ghost receives its owner's value(s)
inline void Comm::forward_comm_fix(Fix* fix, int) {
  forward_calls++;
  const int ng = (int)ghost_owner.size();
  if (ng == 0) return;
  std::vector<double> buf((size_t)ng * (fix->comm_forward > 0 ? fix->comm_forward : 1));
  fix->pack_forward_comm(ng, ghost_owner.data(), buf.data(), 0, nullptr);
  fix->unpack_forward_comm(ng, atom->nlocal, buf.data());
}

}  // namespace LAMMPS_NS
