// Host-side build of the product's deterministic math (include/slgpu_math.h) for tests/test_math.py.
#include <cmath>
#include "slgpu_math.h"
extern "C" {
void v_exp(const double* x, int n, double* o) { for (int i = 0; i < n; i++) o[i] = slm_exp(x[i]); }
void v_log(const double* x, int n, double* o) { for (int i = 0; i < n; i++) o[i] = slm_log(x[i]); }
void v_sincos2pi(const double* x, int n, double* s, double* c) { for (int i = 0; i < n; i++) slm_sincos2pi(x[i], s + i, c + i); }
// the step kernels' division by a shared divisor (slgpu_step_fast.cuh div_by): must equal n / c bit for bit
void v_div_by(const double* n, const double* c, int cnt, double* o) {
  for (int i = 0; i < cnt; i++) {
    const double rc = 1.0 / c[i];
    const double q0 = n[i] * rc;
    const double r = std::fma(-q0, c[i], n[i]);
    o[i] = std::fma(r, rc, q0);
  }
}
}
