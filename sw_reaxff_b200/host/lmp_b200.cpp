// lmp_b200 — minimal LAMMPS-like driver for the B200 ReaxFF styles:  lmp_b200 -in in.reaxc.lattice -var S 2 -var t 20
// (same command line as the reference's run.sh:2 minus the Sunway launcher).  Also exports rxh_run_script for tests.
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "mini_lammps.h"

using namespace LAMMPS_MINI;

extern "C" long rxh_run_script(const char* script, int nvars, const char* const* names, const char* const* values, int device,
                               double* thermo19, long max_rows, char* err, int errlen) {
  try {
    LAMMPS lmp;
    lmp.cuda_device = device;
    lmp.echo_thermo = false;
    for (int i = 0; i < nvars; i++) lmp.vars[names[i]] = values[i];
    lmp.file(script);
    long n = 0;
    for (const Thermo& t : lmp.thermo_log) {
      if (n >= max_rows) break;
      double* o = thermo19 + 19 * n++;
      o[0] = (double)t.step; o[1] = t.temp; o[2] = t.pe; o[3] = t.ke; o[4] = t.etotal;
      memcpy(o + 5, t.pvector, 14 * sizeof(double));
    }
    return n;
  } catch (const std::exception& e) {
    if (err && errlen > 0) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    return -1;
  }
}

// runs a script and returns the per-atom array of compute `cid` as evaluated at the last thermo output step:
// out[nlocal][ncols], returns nlocal * ncols (tests of the compute SPEC/ATOM style)
extern "C" long rxh_run_script_compute(const char* script, int nvars, const char* const* names, const char* const* values,
                                       int device, const char* cid, double* out, long cap, int* ncols, char* err, int errlen) {
  try {
    LAMMPS lmp;
    lmp.cuda_device = device;
    lmp.echo_thermo = false;
    for (int i = 0; i < nvars; i++) lmp.vars[names[i]] = values[i];
    lmp.file(script);
    for (auto& c : lmp.computes) {
      if (c->id != cid) continue;
      const long m = (long)c->array.size();
      if (ncols) *ncols = c->size_peratom_cols ? c->size_peratom_cols : 1;
      for (long k = 0; k < m && k < cap; k++) out[k] = c->array[k];
      return m;
    }
    throw std::runtime_error(std::string("no compute with ID ") + cid);
  } catch (const std::exception& e) {
    if (err && errlen > 0) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    return -1;
  }
}

// velocities that `velocity all create T seed` gives the atoms of a data file (tests feed them to the CPU oracle)
extern "C" long rxh_velocities(const char* datafile, double T, long seed, double* v, long cap_atoms) {
  try {
    LAMMPS lmp;
    lmp.echo_thermo = false;
    lmp.one("units real");
    lmp.one("atom_style charge");
    lmp.one(std::string("read_data ") + datafile);
    lmp.one("velocity all create " + std::to_string(T) + " " + std::to_string(seed));
    const long n = lmp.atom->nlocal;
    if (n > cap_atoms) return -2;
    memcpy(v, lmp.atom->v.data(), (size_t)3 * n * sizeof(double));
    return n;
  } catch (const std::exception&) {
    return -1;
  }
}

// e2e benchmark hook: executes `script` (which must NOT contain a run command), then setup + `warm` untimed steps +
// `steps` timed steps of the Verlet loop with host buffers.  out4 = natoms, nall, seconds, last PotEng.
extern "C" int rxh_bench_script(const char* script, int nvars, const char* const* names, const char* const* values, int device,
                                int warm, int steps, double* out4, char* err, int errlen) {
  try {
    LAMMPS lmp;
    lmp.cuda_device = device;
    lmp.echo_thermo = false;
    for (int i = 0; i < nvars; i++) lmp.vars[names[i]] = values[i];
    lmp.file(script);
    lmp.setup();
    lmp.iterate(warm);
    for (double& p : lmp.phase_s) p = 0.0;
    auto t0 = std::chrono::steady_clock::now();
    lmp.iterate(steps);
    out4[2] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out4[0] = (double)lmp.atom->natoms; out4[1] = (double)lmp.atom->nall();
    out4[3] = lmp.thermo_log.empty() ? 0.0 : lmp.thermo_log.back().pe;
    if (getenv("RXH_PHASES")) {
      static const char* nm[8] = {"integrate", "borders", "forward_comm", "zero_f", "pre_force", "pair_compute", "reverse_comm", "final+output"};
      for (int k = 0; k < 8; k++) fprintf(stderr, "  host phase %-14s %8.3f ms/step\n", nm[k], 1e3 * lmp.phase_s[k] / steps);
    }
    return 0;
  } catch (const std::exception& e) {
    if (err && errlen > 0) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    return -1;
  }
}

#ifndef RXH_NO_MAIN
int main(int argc, char** argv) {
  LAMMPS lmp;
  std::string in;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-in") && i + 1 < argc) in = argv[++i];
    else if (!strcmp(argv[i], "-var") && i + 2 < argc) { lmp.vars[argv[i + 1]] = argv[i + 2]; i += 2; }
    else if (!strcmp(argv[i], "-device") && i + 1 < argc) lmp.cuda_device = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-sf") && i + 1 < argc) ++i;   // suffix accepted (the styles are the b200 ones anyway)
    else { fprintf(stderr, "usage: lmp_b200 -in script [-var NAME VALUE]... [-device N]\n"); return 2; }
  }
  if (in.empty()) { fprintf(stderr, "lmp_b200: no input script (-in)\n"); return 2; }
  try {
    auto t0 = std::chrono::steady_clock::now();
    lmp.file(in);
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("Total wall time: %.3f s, %ld atoms, %ld steps\n", s, lmp.atom->natoms, lmp.update->ntimestep);
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
#endif
