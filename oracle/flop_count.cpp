// flop_count.cpp -- TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py).
//
// Compiles the CPU restatement oracle/fill_port.c with `double` replaced by a counting number type, so that the
// floating-point operations of the Cartesian formulation of SURVEY.md App. A are COUNTED BY EXECUTING IT (BASELINE.md
// §3: "exact values ... from an instrumented CPU restatement") instead of estimated.  bench.py divides the counts
// of one fill by the number of elements to get `roofline.flops_per_element`.
// Counted: + - * / (one flop each), sqrt / cbrt / pow (one each, reported separately); comparisons, negation and
// fabs are free.  The dense element block of the port multiplies by structural zeros nowhere: every counted
// operation belongs to a term the in-scope physics needs.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "../include/goma_gpu_fill.h"

static long long g_cnt[4];  // add/sub, mul, div, functions

struct cdouble {
  double v;
  cdouble() : v(0.0) {}
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  cdouble(T x) : v((double)x) {}
};
inline cdouble operator+(cdouble a, cdouble b) { g_cnt[0]++; return a.v + b.v; }
inline cdouble operator-(cdouble a, cdouble b) { g_cnt[0]++; return a.v - b.v; }
inline cdouble operator*(cdouble a, cdouble b) { g_cnt[1]++; return a.v * b.v; }
inline cdouble operator/(cdouble a, cdouble b) { g_cnt[2]++; return a.v / b.v; }
inline cdouble operator-(cdouble a) { return -a.v; }
inline cdouble &operator+=(cdouble &a, cdouble b) { g_cnt[0]++; a.v += b.v; return a; }
inline cdouble &operator-=(cdouble &a, cdouble b) { g_cnt[0]++; a.v -= b.v; return a; }
inline cdouble &operator*=(cdouble &a, cdouble b) { g_cnt[1]++; a.v *= b.v; return a; }
inline cdouble &operator/=(cdouble &a, cdouble b) { g_cnt[2]++; a.v /= b.v; return a; }
inline bool operator<(cdouble a, cdouble b) { return a.v < b.v; }
inline bool operator>(cdouble a, cdouble b) { return a.v > b.v; }
inline bool operator<=(cdouble a, cdouble b) { return a.v <= b.v; }
inline bool operator>=(cdouble a, cdouble b) { return a.v >= b.v; }
inline bool operator==(cdouble a, cdouble b) { return a.v == b.v; }
inline bool operator!=(cdouble a, cdouble b) { return a.v != b.v; }
inline cdouble fabs(cdouble a) { return std::fabs(a.v); }
inline cdouble sqrt(cdouble a) { g_cnt[3]++; return std::sqrt(a.v); }
inline cdouble cbrt(cdouble a) { g_cnt[3]++; return std::cbrt(a.v); }
inline cdouble pow(cdouble a, cdouble b) { g_cnt[3]++; return std::pow(a.v, b.v); }

#define double cdouble
#define goma_port_fill goma_port_fill_counted_impl
#include "fill_port.c"
#undef goma_port_fill
#undef double

static_assert(sizeof(cdouble) == sizeof(double), "cdouble must alias double arrays");

// same signature as goma_port_fill; counts[0..3] = add/sub, mul, div, functions executed by this fill
extern "C" int goma_port_fill_counted(const struct goma_gpu_problem *p, const int *ija, const double *x, const double *x_old,
                                      const double *xdot, double delta_t, double theta, double time_value, double h_elem_avg,
                                      double U_norm, int assemble_residual, int assemble_jacobian, double *a, double *resid,
                                      long long counts[4]) {
  memset(g_cnt, 0, sizeof(g_cnt));
  int rc = goma_port_fill_counted_impl(p, ija, (const cdouble *)x, (const cdouble *)x_old, (const cdouble *)xdot, delta_t, theta,
                                       time_value, h_elem_avg, U_norm, assemble_residual, assemble_jacobian, (cdouble *)a,
                                       (cdouble *)resid);
  for (int k = 0; k < 4; k++) counts[k] = g_cnt[k];
  return rc;
}
