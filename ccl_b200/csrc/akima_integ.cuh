// Akima integral of samples on a uniform grid, one warp per integral.
//
// ccl_integ_spline (src/ccl_utils.c:274-345) -> gsl_spline_init(akima) + gsl_spline_eval_integ over the full range,
// for x_s = xmin + dx*s with the end nodes overwritten (ccl_linear_spacing, src/ccl_utils.c:17-37).  With m_i the
// slopes, b_i the left-node derivative and tL_{i+1} the right-node derivative GSL assigns to interval i, its cubic
// integrates to h (y_i + y_{i+1})/2 + h^2 (b_i - tL_{i+1}) / 12 (c = d = 0 when NE_i == 0), and for NE != 0 both b_i
// and tL_i equal the same Akima node derivative t_i, so the correction telescopes to t_0 - t_{n-1} unless some NE
// vanishes (exact ties, e.g. stretches of zeros), in which case the per-interval form is evaluated.  Same polynomial
// as akima_eval_integ; only the association of the sum differs (<= 1e-15 relative).
#pragma once
#include <cuda_runtime.h>

namespace cclb {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sf[0..nk): the samples (shared or global memory, complete and visible to the warp); nk >= 5.  Every lane
// returns the integral.
__device__ __forceinline__ double akima_integral_uniform(const double* sf, int nk, double lkmin, double lkmax,
                                                         double dx, int lane) {
  // node spacing: interior nodes are lkmin + dx*s; the two end intervals use the overwritten ends
  const double inv_dx = 1.0 / dx;
  auto slope = [&](int j) -> double {  // m_j for j in [-2, nk]
    auto raw = [&](int q) -> double {
      if (q == nk - 2) return (sf[q + 1] - sf[q]) / (lkmax - (lkmin + dx * q));
      return (sf[q + 1] - sf[q]) * inv_dx;
    };
    if (j >= 0 && j <= nk - 2) return raw(j);
    if (j == -1) return 2.0 * raw(0) - raw(1);
    if (j == -2) return 3.0 * raw(0) - 2.0 * raw(1);
    if (j == nk - 1) return 2.0 * raw(nk - 2) - raw(nk - 3);
    return 3.0 * raw(nk - 2) - 2.0 * raw(nk - 3);
  };
  // pass 1: trapezoid sum and tie detection (NE_j == 0 at any node j in [0, nk-1])
  double trap = 0.0;
  bool tie = false;
  for (int j = lane; j < nk; j += 32) {
    const double mm2 = slope(j - 2), mm1 = slope(j - 1), m0 = slope(j), mp1 = slope(j + 1);
    const double NE = fabs(mp1 - m0) + fabs(mm1 - mm2);
    tie = tie || (NE == 0.0);
    if (j < nk - 1) {
      const double h = (j == nk - 2) ? (lkmax - (lkmin + dx * j)) : dx;
      trap += 0.5 * h * (sf[j] + sf[j + 1]);
    }
  }
  const bool any_tie = __any_sync(0xffffffffu, tie);
  auto node_deriv = [&](int j, bool from_left) -> double {
    // Akima derivative at node j as used by interval j (b_j, from_left = false) or by interval
    // j-1 (tL_j, from_left = true); they differ only in GSL's NE == 0 rule.
    const double mm2 = slope(j - 2), mm1 = slope(j - 1), m0 = slope(j), mp1 = slope(j + 1);
    const double NE = fabs(mp1 - m0) + fabs(mm1 - mm2);
    if (NE == 0.0) return from_left ? mm1 : m0;
    const double alpha = fabs(mm1 - mm2) / NE;
    return (1.0 - alpha) * mm1 + alpha * m0;
  };
  double corr = 0.0;
  if (!any_tie) {
    if (lane == 0) {
      const double h0 = dx, hl = lkmax - (lkmin + dx * (nk - 2));
      // sum_i h_i^2 (t_i - t_{i+1}) with h uniform except the last interval
      corr = h0 * h0 * (node_deriv(0, false) - node_deriv(nk - 2, true)) +
             hl * hl * (node_deriv(nk - 2, false) - node_deriv(nk - 1, true));
    }
  } else {
    for (int i = lane; i < nk - 1; i += 32) {
      const double mm2 = slope(i - 2), mm1 = slope(i - 1), m0 = slope(i), mp1 = slope(i + 1);
      const double NE = fabs(mp1 - m0) + fabs(mm1 - mm2);
      if (NE != 0.0) {
        const double h = (i == nk - 2) ? (lkmax - (lkmin + dx * i)) : dx;
        corr += h * h * (node_deriv(i, false) - node_deriv(i + 1, true));
      }
    }
  }
  return warp_sum(trap + corr * (1.0 / 12.0));
}

}  // namespace cclb
