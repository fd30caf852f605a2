// gsl akima_init + akima_calc (interpolation/akima.c) for one interval: {y_i, b_i, c_i, d_i} of
//   y(x) = y_i + dx (b_i + dx (c_i + dx d_i)),  dx = x - x_i,
// from the five slopes around interval i; the two slopes beyond either end repeat the first difference of
// slopes (non-periodic boundary rule).  n >= 5 nodes.
#pragma once
#include <cuda_runtime.h>

namespace cclb {

__device__ __forceinline__ double4 akima_interval_coef(const double* x, const double* y, int n, int i) {
  auto slope = [&](int j) -> double {  // m[j], j in [-2, n]
    auto raw = [&](int q) { return (y[q + 1] - y[q]) / (x[q + 1] - x[q]); };
    if (j >= 0 && j <= n - 2) return raw(j);
    if (j == -1) return 2.0 * raw(0) - raw(1);
    if (j == -2) return 3.0 * raw(0) - 2.0 * raw(1);
    if (j == n - 1) return 2.0 * raw(n - 2) - raw(n - 3);
    return 3.0 * raw(n - 2) - 2.0 * raw(n - 3);
  };
  const double mm2 = slope(i - 2), mm1 = slope(i - 1), m0 = slope(i), mp1 = slope(i + 1), mp2 = slope(i + 2);
  const double NE = fabs(mp1 - m0) + fabs(mm1 - mm2);
  double b, c, d;
  if (NE == 0.0) {
    b = m0; c = 0.0; d = 0.0;
  } else {
    const double h_i = x[i + 1] - x[i];
    const double NE_next = fabs(mp2 - mp1) + fabs(m0 - mm1);
    const double alpha_i = fabs(mm1 - mm2) / NE;
    double tL_ip1;
    if (NE_next == 0.0) tL_ip1 = m0;
    else {
      const double alpha_ip1 = fabs(m0 - mm1) / NE_next;
      tL_ip1 = (1.0 - alpha_ip1) * m0 + alpha_ip1 * mp1;
    }
    b = (1.0 - alpha_i) * mm1 + alpha_i * m0;
    c = (3.0 * m0 - 2.0 * b - tL_ip1) / h_i;
    d = (b + tL_ip1 - 2.0 * m0) / (h_i * h_i);
  }
  return make_double4(y[i], b, c, d);
}

}  // namespace cclb
