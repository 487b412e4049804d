// Host accuracy check of sw_reaxff_b200/csrc/rxb_math.cuh (same source as the device code; every operation is an explicit
// fma/add/mul, so these results are the device results).  Prints max errors against 80-bit libm; tests/test_fast_math.py
// asserts on them.
#include <cstdio>
#include <random>

#include "rxb_math.cuh"

using namespace rxb::fm;

int main() {
  const double* tab = host_tables();
  std::mt19937_64 g(12345);
  std::uniform_real_distribution<double> wide(-700.0, 700.0), narrow(-60.0, 10.0), lexp(-690.0, 690.0), dec(-3.0, 4.5), unit(-1.0, 1.0);
  double e_exp = 0, e_log = 0, e_log1 = 0, e_pow = 0, e_cbrt = 0, e_cbrt_seed = 0;
  const int n = 2000000;
  for (int i = 0; i < n; i++) {
    const double x = (i & 1) ? wide(g) : narrow(g);
    const long double ex = expl((long double)x);
    e_exp = fmax(e_exp, fabs((double)((exp_b(x, tab) - ex) / ex)));
    const double lx = exp(lexp(g) * ((i & 2) ? 1.0 : 0.01));
    const long double ll = logl((long double)lx);
    const double ea = fabs((double)(log_b(lx, tab) - ll));
    e_log = fmax(e_log, ea / fmax(1.0, fabs((double)ll)));
    if (fabs(lx - 1.0) < 0.5) e_log1 = fmax(e_log1, ea);
    // the chain of the vdW kernel: r^p as exp(p log r), r in [0.3, 12]
    const double r = 0.3 + 11.7 * (i % 100003) / 100003.0, p = 1.5591;
    const long double pr = powl((long double)r, (long double)p);
    e_pow = fmax(e_pow, fabs((double)((exp_b(p * log_b(r, tab), tab) - pr) / pr)));
    const double c = pow(10.0, dec(g));
    const long double cr = powl((long double)c, -1.0L / 3.0L);
    e_cbrt = fmax(e_cbrt, fabs((double)((rcbrt_b(c) - cr) / cr)));
    // any seed within 1e-5 (the device seed is two MUFU approximations, ~1e-6) must still converge
    e_cbrt_seed = fmax(e_cbrt_seed, fabs((double)((rcbrt_seeded(c, (double)cr * (1.0 + 1e-5 * unit(g))) - cr) / cr)));
  }
  printf("exp_rel %.3e\nlog_abs_over_max1 %.3e\nlog_abs_near_1 %.3e\npow_rel %.3e\nrcbrt_rel %.3e\nrcbrt_seed1e-5_rel %.3e\n", e_exp, e_log,
         e_log1, e_pow, e_cbrt, e_cbrt_seed);
  // edges: exponent boundaries of log (2^k and its neighbours), the clamp ends of exp, table-interval boundaries
  double e_edge_log = 0, e_edge_exp = 0;
  for (int k = -1000; k <= 1000; k += 7) {
    const double base = ldexp(1.0, k);
    const double xs[5] = {base, nextafter(base, 0.0), nextafter(base, INFINITY), base * (1.0 + 1.0 / 32.0), nextafter(base * (1.0 + 1.0 / 32.0), 0.0)};
    for (double x : xs) {
      const long double ll = logl((long double)x);
      e_edge_log = fmax(e_edge_log, fabs((double)(log_b(x, tab) - ll)) / fmax(1.0, fabs((double)ll)));
    }
  }
  const double ends[8] = {-700.0, 700.0, -699.999999, 699.999999, -0.0, 5e-324, 0.021660849392498291, -0.021660849392498291};
  for (double x : ends) {
    const long double ex = expl((long double)x);
    e_edge_exp = fmax(e_edge_exp, fabs((double)((exp_b(x, tab) - ex) / ex)));
  }
  // monotonic where the kernels rely on it (a larger distance never gives a larger r^p)
  int mono_bad = 0;
  double prev = 0.0;
  for (int i = 0; i < 200000; i++) {
    const double r = 0.5 + 12.0 * i / 200000.0;
    const double y = exp_b(1.5591 * log_b(r, tab), tab);
    if (y < prev * (1.0 - 4e-16)) mono_bad++;
    prev = y;
  }
  printf("edge_log %.3e\nedge_exp %.3e\nmono_bad %d\n", e_edge_log, e_edge_exp, mono_bad);
  printf("exp0 %.17g\nlog1 %.17g\n", exp_b(0.0, tab), log_b(1.0, tab));
  return 0;
}
