// sm_100a kernels of the Limber C_ell path:
//   * table preparation (GSL akima / natural cspline / bicubic derivative planes, index LUTs)
//   * the Limber integrators: 'spline' (src/ccl_cls.c:189-225 + src/ccl_utils.c:274-345)
//     and 'qag_quad' (src/ccl_cls.c:227-263, gsl_integration_qag with GK41)
// One warp owns one (cosmology, pair, ell) C_ell: lanes walk the ln k sample axis, the
// integrand samples live in shared memory, the Akima / Gauss-Kronrod sums are reduced with
// warp shuffles.  Bound: fp64 DFMA pipe.
#include <math_constants.h>
#include <cfloat>
#include <cstdio>

#include "akima_coef.cuh"
#include "akima_integ.cuh"
#include "device_eval.cuh"
#include "gk41_table.h"

namespace cclb {

// ------------------------------------------------------------------------------------------
// Table preparation
// ------------------------------------------------------------------------------------------

// gsl akima_init + akima_calc (interpolation/akima.c), one CTA per spline.
__global__ void akima_coef_kernel(const AkimaJob* jobs) {
  const AkimaJob jb = jobs[blockIdx.x];
  const int n = jb.n;
  const double* x = jb.x;
  const double* y = jb.y;
  for (int i = threadIdx.x; i < n - 1; i += blockDim.x) {
    const double4 c = akima_interval_coef(x, y, n, i);
    jb.xr[i] = make_double2(x[i], x[i + 1]);
    jb.cr[i] = c;
  }
}

// Index LUT: lut[j] = gsl_interp_bsearch(x, x0 + j*bin).
__global__ void lut_kernel(const LutJob* jobs) {
  const LutJob jb = jobs[blockIdx.x];
  const double x0 = jb.x[0];
  const double bin = (jb.x[jb.n - 1] - x0) / jb.nlut;
  for (int j = threadIdx.x; j < jb.nlut; j += blockDim.x) {
    // evaluated just left of the bin edge, so that a lookup whose bin index rounds up by one still
    // starts at or before its interval (find_rec only walks right on this path)
    const double v = x0 + (j - 1e-6) * bin;
    int lo = 0, hi = jb.n - 1;
    while (hi > lo + 1) {
      const int m = (hi + lo) >> 1;
      if (jb.x[m] > v) hi = m; else lo = m;
    }
    // guard against rounding of x0 + j*bin: never start past the true interval
    while (lo > 0 && jb.x[lo] > v) --lo;
    jb.lut[j] = lo;
  }
}

// Bicubic axis records {x_i, x_{i+1}, 1/(x_{i+1}-x_i)}.
__global__ void axis_kernel(const AxisJob* jobs) {
  const AxisJob jb = jobs[blockIdx.x];
  for (int i = threadIdx.x; i < jb.n - 1; i += blockDim.x) {
    const double x0 = jb.x[i], x1 = jb.x[i + 1];
    jb.rec[i] = AxisRec{x0, x1, 1.0 / (x1 - x0), 0.0};
  }
}

// LDL^T factors of the natural-spline tridiagonal system (gsl linalg/tridiag.c solve_tridiag):
// alpha/gamma depend on the node spacing only.  sys = n-2 unknowns c_1..c_{n-2}.
__device__ void tridiag_factor(const double* x, int n, double* alpha, double* gamma) {
  const int N = n - 2;
  if (N < 1) return;
  auto diag = [&](int i) { return 2.0 * ((x[i + 2] - x[i + 1]) + (x[i + 1] - x[i])); };
  auto off = [&](int i) { return x[i + 2] - x[i + 1]; };
  alpha[0] = diag(0);
  gamma[0] = off(0) / alpha[0];
  for (int i = 1; i < N - 1; ++i) {
    alpha[i] = diag(i) - off(i - 1) * gamma[i - 1];
    gamma[i] = off(i) / alpha[i];
  }
  if (N > 1) alpha[N - 1] = diag(N - 1) - off(N - 2) * gamma[N - 2];
}

// Solve for the natural-spline c_i of NP data planes y_p(stride) on the same nodes x and emit, per
// node, the first derivative gsl_spline_eval_deriv(spline, x_i) into out_p(stride); plane p has its
// data at base + yoff[p] and its output at base + ooff[p].  out is also the scratch of the forward
// sweep.  Mirrors cspline_init + cspline_eval_deriv operation by operation, except that the grid-only
// reciprocals are precomputed once per grid (AxisFactors: 1/h_i as in GSL, 1/alpha_i and 1/3 replacing
// three divisions by multiplications: a last-bit difference).  One thread runs the two serial sweeps;
// their only true dependence is one FMA per node and plane, so the strided table reads of PF nodes are
// issued together ahead of the chain (out aliases y's buffer: the compiler cannot hoist them), and the
// planes of one node share its cache line.
constexpr int PF = 8;

struct AxisFactors {
  const double* alpha;   // LDL^T diagonal
  const double* gamma;   // LDL^T off-diagonal factor
  const double* inv_h;   // 1 / (x_{i+1} - x_i), 0 for a zero interval
  const double* inv_al;  // 1 / alpha_i
};

__device__ __forceinline__ AxisFactors axis_factors(double* scratch, int n) {
  return AxisFactors{scratch, scratch + n, scratch + 2 * n, scratch + 3 * n};
}

template <int NP>
__device__ __forceinline__ void natural_spline_node_derivs(const double* x, int n, double* base, size_t stride,
                                                           const int (&yoff)[NP], const int (&ooff)[NP],
                                                           const AxisFactors F) {
  const int N = n - 2;
  const double third = 1.0 / 3.0;
  auto Y = [&](int p, int i) { return base[(size_t)i * stride + yoff[p]]; };
  auto O = [&](int p, int i) -> double& { return base[(size_t)i * stride + ooff[p]]; };
  // forward sweep: z_i = G_i - gamma_{i-1} z_{i-1}, c~_i = z_i / alpha_i stored at node i+1, with
  // G_i = 3 ((y_{i+2} - y_{i+1}) / h_{i+1} - (y_{i+1} - y_i) / h_i)
  if (N == 1) {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const double G0 = 3.0 * ((Y(p, 2) - Y(p, 1)) * F.inv_h[1] - (Y(p, 1) - Y(p, 0)) * F.inv_h[0]);
      O(p, 1) = G0 / (2.0 * ((x[2] - x[1]) + (x[1] - x[0])));
    }
  } else {
    const double g0 = F.inv_h[0];
    double ya[NP], slope[NP], z[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      ya[p] = Y(p, 1);
      slope[p] = (ya[p] - Y(p, 0)) * g0;  // (y_{i+1} - y_i) / h_i of the current i
      z[p] = 0.0;
    }
    for (int i0 = 0; i0 < N; i0 += PF) {
      double yb[NP][PF];
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int i = min(i0 + u, N - 1);
#pragma unroll
        for (int p = 0; p < NP; ++p) yb[p][u] = Y(p, i + 2);
      }
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int i = i0 + u;
        if (i < N) {
          const double g = F.inv_h[i + 1];
          const double ial = F.inv_al[i], gm = F.gamma[max(i - 1, 0)];
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const double snext = (yb[p][u] - ya[p]) * g;
            const double G = 3.0 * (snext - slope[p]);
            z[p] = (i == 0) ? G : G - gm * z[p];
            O(p, i + 1) = z[p] * ial;
            slope[p] = snext;
            ya[p] = yb[p][u];
          }
        }
      }
    }
  }
  // back substitution fused with the derivative evaluation.  c_node[n-1] = 0.
  double c_next[NP], yr[NP];
  double xr = x[n - 1];
  {
    const double xl = x[n - 2];
    const double h = xr - xl, g = F.inv_h[n - 2];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const double c_cur = O(p, n - 2);  // node n-2 <-> system index N-1: c = c~
      const double yl = Y(p, n - 2);
      const double dyv = Y(p, n - 1) - yl;
      const double b = (dyv * g) - h * (0.0 + 2.0 * c_cur) * third;
      const double d = (0.0 - c_cur) * (third * g);
      O(p, n - 1) = b + h * (2.0 * c_cur + 3.0 * d * h);
      O(p, n - 2) = b;
      c_next[p] = c_cur;
      yr[p] = yl;
    }
    xr = xl;
  }
  for (int node0 = n - 3; node0 >= 0; node0 -= PF) {
    double yb[NP][PF], ct[NP][PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int node = max(node0 - u, 0);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        yb[p][u] = Y(p, node);
        ct[p][u] = O(p, node);
      }
    }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int node = node0 - u;
      if (node >= 0) {
        const double xl = x[node];
        const double gm = F.gamma[max(node - 1, 0)];
        const double h = xr - xl, g = F.inv_h[node];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const double c_cur = (node >= 1) ? ct[p][u] - gm * c_next[p] : 0.0;
          const double dyv = yr[p] - yb[p][u];
          O(p, node) = (dyv * g) - h * (c_next[p] + 2.0 * c_cur) * third;
          c_next[p] = c_cur;
          yr[p] = yb[p][u];
        }
        xr = xl;
      }
    }
  }
}

// Factorised-transfer splines: natural cspline coefficients {y, b, c, d} per interval
// (cspline_init + coeff_calc of interpolation/cspline.c).  One warp per job: the nodes are staged in
// shared memory with coalesced loads, lane 0 runs the three serial sweeps there (same operations in
// the same order as GSL), all lanes write the interval records.  Layout of the dynamic shared
// memory: x[n] y[n] alpha[n] gamma[n] c[n] 1/h[n].
__global__ void cspline_coef_smem_kernel(const CsplineJob* jobs, int max_n) {
  extern __shared__ double csm[];
  const CsplineJob jb = jobs[blockIdx.x];
  const int n = jb.n, lane = threadIdx.x;
  double* x = csm;
  double* y = x + max_n;
  double* alpha = y + max_n;
  double* gamma = alpha + max_n;
  double* c = gamma + max_n;
  double* ih = c + max_n;
  const int N = n - 2;
  for (int i = lane; i < n; i += 32) { x[i] = jb.x[i]; y[i] = jb.y[i]; }
  __syncwarp();
  // everything without a serial dependence runs on all lanes: 1/h_i, the right-hand side G_i (kept in
  // c[i+1] until the forward sweep overwrites it), the division by alpha_i and the interval records
  for (int i = lane; i < n - 1; i += 32) {
    const double h = x[i + 1] - x[i];
    ih[i] = (h != 0.0) ? 1.0 / h : 0.0;
  }
  __syncwarp();
  for (int i = lane; i < N; i += 32) c[i + 1] = 3.0 * ((y[i + 2] - y[i + 1]) * ih[i + 1] - (y[i + 1] - y[i]) * ih[i]);
  if (lane == 0) { c[0] = 0.0; c[n - 1] = 0.0; }
  __syncwarp();
  if (N == 1) {
    if (lane == 0) c[1] = c[1] / (2.0 * ((x[2] - x[1]) + (x[1] - x[0])));
  } else {
    if (lane == 0) {
      tridiag_factor(x, n, alpha, gamma);
      double z = c[1];
      for (int i = 1; i < N; ++i) {  // z_i = G_i - gamma_{i-1} z_{i-1}
        z = c[i + 1] - gamma[i - 1] * z;
        c[i + 1] = z;
      }
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) c[i + 1] = c[i + 1] / alpha[i];
    __syncwarp();
    if (lane == 0)
      for (int i = N - 2; i >= 0; --i) c[i + 1] = c[i + 1] - gamma[i] * c[i + 2];
  }
  __syncwarp();
  for (int i = lane; i < n - 1; i += 32) {
    const double dx = x[i + 1] - x[i];
    const double dy = y[i + 1] - y[i];
    const double b = (dy / dx) - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
    const double d = (c[i + 1] - c[i]) / (3.0 * dx);
    jb.xr[i] = make_double2(x[i], x[i + 1]);
    jb.cr[i] = make_double4(y[i], b, c[i], d);
  }
}

// Same for splines too long for shared memory: one thread per job on global scratch.
__global__ void cspline_coef_kernel(const CsplineJob* jobs, int njobs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  const CsplineJob jb = jobs[j];
  const int n = jb.n;
  const double* x = jb.x;
  const double* y = jb.y;
  double* alpha = jb.scratch;
  double* gamma = jb.scratch + n;
  double* c = jb.scratch + 2 * n;  // c at the nodes
  const int N = n - 2;
  c[0] = 0.0;
  c[n - 1] = 0.0;
  auto G = [&](int i) {
    const double h_i = x[i + 1] - x[i];
    const double h_ip1 = x[i + 2] - x[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0;
    const double g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    return 3.0 * ((y[i + 2] - y[i + 1]) * g_ip1 - (y[i + 1] - y[i]) * g_i);
  };
  if (N == 1) {
    c[1] = G(0) / (2.0 * ((x[2] - x[1]) + (x[1] - x[0])));
  } else {
    tridiag_factor(x, n, alpha, gamma);
    double z = G(0);
    c[1] = z / alpha[0];
    for (int i = 1; i < N; ++i) {
      z = G(i) - gamma[i - 1] * z;
      c[i + 1] = z / alpha[i];
    }
    for (int i = N - 2; i >= 0; --i) c[i + 1] = c[i + 1] - gamma[i] * c[i + 2];
  }
  for (int i = 0; i < n - 1; ++i) {
    const double dx = x[i + 1] - x[i];
    const double dy = y[i + 1] - y[i];
    const double b = (dy / dx) - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
    const double d = (c[i + 1] - c[i]) / (3.0 * dx);
    jb.xr[i] = make_double2(x[i], x[i + 1]);
    jb.cr[i] = make_double4(y[i], b, c[i], d);
  }
}

// bicubic_init (interp2d/bicubic.c).  Pass 0: axis factors (once per distinct pair of grids) + copy z.
__global__ void bicubic_factor_kernel(const BicubicJob* jobs) {
  const BicubicJob jb = jobs[blockIdx.x];
  const int nk = jb.nk, na = jb.na;
  const int nn = nk * na;
  for (int i = threadIdx.x; i < nn; i += blockDim.x) jb.tab[i].x = jb.z[i];
  if (!jb.factor) return;
  double* sk = jb.scratch;
  double* sa = jb.scratch + 4 * nk;
  if (threadIdx.x == 0) tridiag_factor(jb.xk, nk, sk, sk + nk);
  if (threadIdx.x == 32) tridiag_factor(jb.ya, na, sa, sa + na);
  auto invh = [](const double* x, int i) {
    const double h = x[i + 1] - x[i];
    return (h != 0.0) ? 1.0 / h : 0.0;
  };
  for (int i = threadIdx.x; i < nk - 1; i += blockDim.x) sk[2 * nk + i] = invh(jb.xk, i);
  for (int i = threadIdx.x; i < na - 1; i += blockDim.x) sa[2 * na + i] = invh(jb.ya, i);
  __syncthreads();
  for (int i = threadIdx.x; i < nk - 2; i += blockDim.x) sk[3 * nk + i] = 1.0 / sk[i];
  for (int i = threadIdx.x; i < na - 2; i += blockDim.x) sa[3 * na + i] = 1.0 / sa[i];
}

// Pass 1: zy (columns, spline along a).  Pass 2: zx and zxy = d/dx of the z and zy planes (rows, spline
// along ln k; one thread carries both planes of its row through the same sweeps).
__global__ void bicubic_deriv_kernel(const BicubicJob* jobs, int pass) {
  const BicubicJob jb = jobs[blockIdx.y];
  const int nk = jb.nk, na = jb.na;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double* tab = reinterpret_cast<double*>(jb.tab);
  if (pass == 1) {
    if (t < nk) {
      const int yoff[1] = {0}, ooff[1] = {2};
      natural_spline_node_derivs<1>(jb.ya, na, tab + (size_t)t * 4, (size_t)nk * 4, yoff, ooff,
                                    axis_factors(jb.scratch + 4 * nk, na));
    }
  }
}

// Pass 2 over the rows of ALL jobs as one index space (row r of job r / max_na): with one CTA row per job a table of
// 50 scale factors leaves 14 of every 64 lanes idle through the whole serial sweep; here only the last warp is partial.
__global__ void bicubic_rows_kernel(const BicubicJob* jobs, int njobs, int max_na) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (int)(r / max_na), t = (int)(r % max_na);
  if (j >= njobs) return;
  const BicubicJob jb = jobs[j];
  if (t >= jb.na) return;
  const int yoff[2] = {0, 2}, ooff[2] = {1, 3};
  natural_spline_node_derivs<2>(jb.xk, jb.nk, reinterpret_cast<double*>(jb.tab) + (size_t)t * jb.nk * 4, 4, yoff, ooff,
                                axis_factors(jb.scratch, jb.nk));
}

void launch_prep(const AkimaJob* d_ak, int nak, const CsplineJob* d_cs, int ncs, const LutJob* d_lut,
                 int nlut, const AxisJob* d_ax, int nax, const BicubicJob* d_bc, int nbc, int max_nk,
                 int max_na, int max_ncs, cudaStream_t stream) {
  if (nak > 0) akima_coef_kernel<<<nak, 128, 0, stream>>>(d_ak);
  if (nlut > 0) lut_kernel<<<nlut, 128, 0, stream>>>(d_lut);
  if (nax > 0) axis_kernel<<<nax, 128, 0, stream>>>(d_ax);
  if (ncs > 0) {
    const size_t csm = sizeof(double) * 6 * (size_t)max_ncs;
    if (csm <= 200 * 1024) {
      if (csm > 48 * 1024)
        cudaFuncSetAttribute(cspline_coef_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm);
      cspline_coef_smem_kernel<<<ncs, 32, csm, stream>>>(d_cs, max_ncs);
    } else
      cspline_coef_kernel<<<(ncs + 63) / 64, 64, 0, stream>>>(d_cs, ncs);
  }
  if (nbc > 0) {
    bicubic_factor_kernel<<<nbc, 256, 0, stream>>>(d_bc);
    dim3 g1((max_nk + 63) / 64, nbc);
    bicubic_deriv_kernel<<<g1, 64, 0, stream>>>(d_bc, 1);
    const long long nrows = (long long)nbc * max_na;
    bicubic_rows_kernel<<<(unsigned)((nrows + 31) / 32), 32, 0, stream>>>(d_bc, nbc, max_na);
  }
}

// ------------------------------------------------------------------------------------------
// Vectorised accessors (the pyccl-facing getters; not on the hot path)
// ------------------------------------------------------------------------------------------
__global__ void eval_achi_kernel(const Cosmo* cs, int n, const double* chi, double* out, int* status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int st = 0;
  out[i] = scale_factor_of_chi(*cs, chi[i], st);
  if (st) atomicOr(status, st);  // ccl_scale_factor_of_chis: *status |= _status
}

__global__ void eval_f2d_kernel(const F2D* f, const Cosmo* cs, int n, const double* lk, const double* a,
                                double* out, int* status, int deriv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int st = 0;
  out[i] = deriv ? f2d_dlogf_dlk_eval(*f, lk[i], a[i], cs, st) : f2d_eval(*f, lk[i], a[i], cs, st);
  if (st) atomicCAS(status, 0, st);
}

__global__ void eval_tracer_kernel_kernel(const Tracer* tr, int n, const double* chi, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = tracer_kernel(*tr, chi[i]);
}

__global__ void eval_tracer_transfer_kernel(const Tracer* tr, int n, const double* lk, const double* a,
                                            double* out, int* status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int st = 0;
  out[i] = tracer_transfer(*tr, lk[i], a[i], st);
  if (st) atomicCAS(status, 0, st);
}

void launch_eval_achi(const Cosmo* d_cs, int n, const double* d_chi, double* d_out, int* d_status,
                      cudaStream_t stream) {
  if (n > 0) eval_achi_kernel<<<(n + 127) / 128, 128, 0, stream>>>(d_cs, n, d_chi, d_out, d_status);
}
void launch_eval_f2d(const F2D* d_f, const Cosmo* d_cs, int n, const double* d_lk, const double* d_a,
                     double* d_out, int* d_status, cudaStream_t stream, int deriv) {
  if (n > 0) eval_f2d_kernel<<<(n + 127) / 128, 128, 0, stream>>>(d_f, d_cs, n, d_lk, d_a, d_out, d_status, deriv);
}
void launch_eval_tracer_kernel(const Tracer* d_tr, int n, const double* d_chi, double* d_out,
                               cudaStream_t stream) {
  if (n > 0) eval_tracer_kernel_kernel<<<(n + 127) / 128, 128, 0, stream>>>(d_tr, n, d_chi, d_out);
}
void launch_eval_tracer_transfer(const Tracer* d_tr, int n, const double* d_lk, const double* d_a,
                                 double* d_out, int* d_status, cudaStream_t stream) {
  if (n > 0)
    eval_tracer_transfer_kernel<<<(n + 127) / 128, 128, 0, stream>>>(d_tr, n, d_lk, d_a, d_out, d_status);
}

// ------------------------------------------------------------------------------------------
// Limber integrators
// ------------------------------------------------------------------------------------------
constexpr int WARPS_PER_CTA = 4;

__device__ __forceinline__ int warp_first_nonzero(int st) {
  const unsigned bad = __ballot_sync(0xffffffffu, st != 0);
  return bad ? __shfl_sync(0xffffffffu, st, __ffs(bad) - 1) : 0;
}

__device__ __forceinline__ bool setup_task(const LimberProblem& P, long long task, TaskCtx& c, int& ic,
                                           int& ip, int& il) {
  il = (int)(task % P.nl);
  const long long r = task / P.nl;
  ip = (int)(r % P.npair);
  ic = (int)(r / P.npair);
  const Cosmo* cs = P.cosmos + ic;
  const int2 pr = P.pairs[ip];
  const int2 c1 = P.colls[pr.x], c2 = P.colls[pr.y];
  c.cs = cs;
  c.t1 = cs->tracers + c1.x;
  c.n1 = c1.y;
  c.t2 = cs->tracers + c2.x;
  c.n2 = c2.y;
  c.same = (pr.x == pr.y);
  c.l = P.ells[il];
  return true;
}

// per-warp prefactor table: f_ell (/ (l+1/2)^2) of the first MAX_SUB sub-tracers of each side
__device__ __forceinline__ void setup_prefactors(const TaskCtx& c, double* pref, int lane) {
  if (lane < MAX_SUB) {
    if (lane < c.n1) pref[lane] = tracer_prefactor(c.t1[lane], c.l);
    if (lane < c.n2) pref[MAX_SUB + lane] = tracer_prefactor(c.t2[lane], c.l);
  }
  __syncwarp();
}

__device__ __forceinline__ void finish_task(const LimberProblem& P, int ic, int ip, int il, double result,
                                            int st, int lane) {
  if (lane != 0) return;
  const size_t o = ((size_t)ic * P.npair + ip) * P.nl + il;
  if (st == 0) {
    P.cl[o] = result / (P.ells[il] + 0.5);  // src/ccl_cls.c:340
  } else {
    P.cl[o] = nan("");                       // src/ccl_cls.c:343-345
    const size_t s = (size_t)ic * P.npair + ip;
    P.status[s] = ST_ERR_INTEG;
    atomicCAS(P.status_detail + s, 0, st);
  }
}

// integ_cls_limber_spline (src/ccl_cls.c:189-225) + ccl_integ_spline / akima_eval_integ
// (src/ccl_utils.c:274-345, gsl interpolation/akima.c).
//
// Node abscissae: lk_s = lkmin + dx*s (ccl_linear_spacing, ends overwritten).  k_s = exp(lk_s) is
// advanced per lane by the constant ratio exp(32 dx) (one exp per lane per C_ell instead of one
// per node; relative drift <= 22 steps * 2 ulp), chi_s = (l+1/2)/k_s by the inverse ratio.
//
// Akima integral on the (uniform) node grid.  With m_i the slopes, b_i the left-node derivative
// and tL_{i+1} the right-node derivative GSL assigns to interval i, its cubic integrates to
//     h (y_i + y_{i+1})/2 + h^2 (b_i - tL_{i+1}) / 12            (c = d = 0 when NE_i == 0),
// and for NE != 0 both b_i and tL_i equal the same Akima node derivative t_i, so the correction
// telescopes to t_0 - t_{n-1} unless some NE vanishes (exact ties, e.g. stretches of zeros), in
// which case the per-interval form is evaluated.  Same polynomial as akima_eval_integ; only the
// association of the sum differs (<= 1e-15 relative).
__device__ __forceinline__ void spline_task(const LimberProblem& P, long long task, double* sf, double* pref,
                                            int lane) {
  TaskCtx c;
  int ic, ip, il;
  setup_task(P, task, c, ic, ip, il);
  if (!c.cs->has_achi) {  // src/ccl_cls.c:274-280
    finish_task(P, ic, ip, il, 0.0, ST_ERR_DISTANCES_INIT, lane);
    return;
  }
  double lkmin, lkmax;
  get_k_interval(c, P.par, lkmin, lkmax);
  const int nk = (int)(fmax((lkmax - lkmin) / P.par.DLOGK_INTEGRATION + 0.5, 1.0)) + 1;
  if (nk < 5 || nk > P.par.nk_cap) {
    // gsl_spline_alloc(akima, n<5) == NULL -> CCL_ERROR_MEMORY.  nk > cap cannot
    // happen: the host sizes nk_cap from (K_MAX, K_MIN, DLOGK); flagged rather than overrun.
    finish_task(P, ic, ip, il, 0.0, ST_ERR_MEMORY, lane);
    return;
  }
  setup_prefactors(c, pref, lane);
  // ccl_linear_spacing (src/ccl_utils.c:17-37)
  const double dx = (lkmax - lkmin) / (nk - 1.);
  const double lp1h = c.l + 0.5;
  double k = exp(lkmin + dx * lane);
  double chi = lp1h / k;
  const double ratio = exp(32.0 * dx);
  const double iratio = 1.0 / ratio;
  int st = 0;
  for (int s = lane; s < nk; s += 32) {
    const double lk = (s == 0) ? lkmin : ((s == nk - 1) ? lkmax : lkmin + dx * s);
    int s1 = 0;
    sf[s] = cl_integrand_kchi(c, pref, pref + MAX_SUB, lk, k, chi, s1);
    if (s1 && !st) st = s1;
    k *= ratio;
    chi *= iratio;
  }
  st = warp_first_nonzero(st);  // any failing node fails the C_ell
  __syncwarp();
  if (st) {
    finish_task(P, ic, ip, il, 0.0, st, lane);
    return;
  }
  const double acc = akima_integral_uniform(sf, nk, lkmin, lkmax, dx, lane);
  finish_task(P, ic, ip, il, acc, 0, lane);
}

// General 'spline' kernel: one warp per (cosmology, pair, ell), every tracer / P(k) form.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4)   // 128 registers: 16 warps per SM
limber_spline_kernel(const LimberProblem P, long long ntasks) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sf = smem + (size_t)warp * (P.par.nk_cap + 2 * MAX_SUB);
  double* pref = sf + P.par.nk_cap;
  const long long task = (long long)blockIdx.x * WARPS_PER_CTA + warp;
  if (task >= ntasks) return;
  spline_task(P, task, sf, pref, lane);
}

// Persistent variant over the redo list the fast kernel leaves behind (tasks it could not take:
// a node outside the a(chi) or P(k,a) tables, ...).  list[0] = count; when the list overflowed
// every task is recomputed.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
limber_spline_redo_kernel(const LimberProblem P, long long ntasks, const unsigned long long* list,
                          unsigned long long cap) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sf = smem + (size_t)warp * (P.par.nk_cap + 2 * MAX_SUB);
  double* pref = sf + P.par.nk_cap;
  const unsigned long long count = list[0];
  if (count == 0) return;
  const bool overflow = count > cap;
  const unsigned long long n = overflow ? (unsigned long long)ntasks : count;
  const unsigned long long stride = (unsigned long long)gridDim.x * WARPS_PER_CTA;
  for (unsigned long long t = (unsigned long long)blockIdx.x * WARPS_PER_CTA + warp; t < n; t += stride) {
    const long long task = overflow ? (long long)t : (long long)list[1 + t];
    __syncwarp();
    spline_task(P, task, sf, pref, lane);
  }
}

// ---- qag_quad ------------------------------------------------------------------------
__constant__ double c_xgk[21];
__constant__ double c_wgk[21];
__constant__ double c_wg[10];
static bool g_gk_uploaded[64] = {false};

// rescale_error (integration/qk.c)
__device__ __forceinline__ double rescale_error(double err, double result_abs, double result_asc) {
  err = fabs(err);
  if (result_asc != 0 && err != 0) {
    const double x = 200 * err / result_asc;
    const double scale = x * sqrt(x);  // pow(x, 1.5): within an ulp, and the value only ranks intervals
    err = (scale < 1) ? result_asc * scale : result_asc;
  }
  if (result_abs > DBL_MIN / (50 * DBL_EPSILON)) {
    const double min_err = 50 * DBL_EPSILON * result_abs;
    if (min_err > err) err = min_err;
  }
  return err;
}

// gsl_integration_qk with the 41-point rule over NP (1 or 2) panels at once: the 41*NP
// abscissae are spread over the warp's lanes, the four sums are shuffle-reduced.
// fv: per-warp shared buffer of 2*41 doubles.
template <int NP>
__device__ __forceinline__ void qk41_warp(const TaskCtx& c, const double* pref, const double* pa,
                                          const double* pb, double* fv, int lane, double* result,
                                          double* abserr, double* resabs, double* resasc, int& st) {
  // evaluation e in [0, 41): e<20 -> center - h*xgk[e]; 20<=e<40 -> center + h*xgk[e-20]; 40 -> center
  for (int e = lane; e < 41 * NP; e += 32) {
    const int p = e / 41, q = e - p * 41;
    const double center = 0.5 * (pa[p] + pb[p]);
    const double half_length = 0.5 * (pb[p] - pa[p]);
    double xq;
    if (q == 40) xq = center;
    else if (q < 20) xq = center - half_length * c_xgk[q];
    else xq = center + half_length * c_xgk[q - 20];
    int s1 = 0;
    fv[e] = cl_integrand(c, pref, pref + MAX_SUB, xq, s1);
    if (s1 && !st) st = s1;
  }
  st = warp_first_nonzero(st);  // every lane leaves with the same status word
  __syncwarp();
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const double* f = fv + 41 * p;
    const double half_length = 0.5 * (pb[p] - pa[p]);
    const double abs_half_length = fabs(half_length);
    double rk = 0, rg = 0, ra = 0;
    double f1 = 0, f2 = 0;
    if (lane < 20) {
      f1 = f[lane];
      f2 = f[lane + 20];
      const double fsum = f1 + f2;
      rk = c_wgk[lane] * fsum;
      ra = c_wgk[lane] * (fabs(f1) + fabs(f2));
      if (lane & 1) rg = c_wg[lane >> 1] * fsum;
    } else if (lane == 20) {
      f1 = f[40];
      rk = f1 * c_wgk[20];
      ra = fabs(rk);
    }
    rk = warp_sum(rk);
    rg = warp_sum(rg);
    ra = warp_sum(ra);
    const double mean = rk * 0.5;
    double rs = 0;
    if (lane < 20) rs = c_wgk[lane] * (fabs(f1 - mean) + fabs(f2 - mean));
    else if (lane == 20) rs = c_wgk[20] * fabs(f1 - mean);
    rs = warp_sum(rs);
    const double err = (rk - rg) * half_length;
    rk *= half_length;
    ra *= abs_half_length;
    rs *= abs_half_length;
    result[p] = rk;
    resabs[p] = ra;
    resasc[p] = rs;
    abserr[p] = rescale_error(err, ra, rs);
  }
  __syncwarp();
}

// subinterval_too_small (integration/util.c)
__device__ __forceinline__ bool subinterval_too_small(double a1, double a2, double b2) {
  const double tmp = (1 + 100 * DBL_EPSILON) * (fabs(a2) + 1000 * DBL_MIN);
  return fabs(a1) <= tmp && fabs(b2) <= tmp;
}

// ------------------------------------------------------------------------------------------
// gsl_integration_cquad (GSL 2.x integration/cquad.c; Gonnet 2010): the reference's fallback after
// gsl_integration_qag reports GSL_EROUND (src/ccl_cls.c:245-259).  Doubly-adaptive Clenshaw-Curtis
// on nested 5/9/17/33 Chebyshev nodes, interpolants kept in an orthonormal Legendre basis, interval
// heap ordered by error.  One warp per C_ell: integrand evaluations and the matrix-vector products
// are spread over the lanes (each coefficient is still ONE sequential sum, as in GSL), the scalar
// bookkeeping is executed redundantly by every lane on data that lane 0 writes to the warp's slice of
// the workspace.  GSL's removal of non-finite samples (`downdate`) is not restated: a non-finite
// integrand fails the C_ell.  Constant tables: tools/gen_cquad.py.
// ------------------------------------------------------------------------------------------
#define CQUAD_TABLE __device__ const
#include "cquad_tables.h"

struct CqIval {
  double a, b, igral, err;
  double c[64];
  double fx[33];
  int depth, rdepth, ndiv, pad;
};

constexpr int CQ_WARPS = 64;          // warps of the fallback kernel (it is a rare path)
constexpr unsigned long long CQ_LIST_CAP = 1ull << 20;

__device__ __forceinline__ void cq_vinvfx(const double* fx, double* c, int d, int lane) {
  const int m = (4 << d) + 1, skip = 8 >> d;
  const double* V = d == 0 ? CQ_V1INV : d == 1 ? CQ_V2INV : d == 2 ? CQ_V3INV : CQ_V4INV;
  for (int i = lane; i < m; i += 32) {
    double acc = 0.0;
    for (int j = 0; j < m; ++j) acc += V[i * m + j] * fx[j * skip];
    c[i] = acc;
  }
  __syncwarp();
}

__device__ __forceinline__ void cq_sift_down(const CqIval* ivals, int* heap, int i, int n) {
  while (2 * i + 1 < n) {
    int j = 2 * i + 1;
    if (j + 1 < n && ivals[heap[j + 1]].err >= ivals[heap[j]].err) j++;
    if (ivals[heap[j]].err <= ivals[heap[i]].err) break;
    const int t = heap[j]; heap[j] = heap[i]; heap[i] = t;
    i = j;
  }
}

// Returns the GSL status; every lane returns the same values.
__device__ int cquad_warp(const TaskCtx& c, const double* pref, double a, double b, double epsabs, double epsrel,
                          int wsize, CqIval* ivals, int* heap, int lane, double& result, int& st,
                          unsigned long long& nev) {
  const int n[4] = {4, 8, 16, 32};
  const int skip[4] = {8, 4, 2, 1};
  const int idx[4] = {0, 5, 14, 31};
  const double w = 0.70710678118654752440;  // M_SQRT2 / 2
  const int ndiv_max = 20;
  const double* xi = CQ_XI;
  if (epsabs < 0.0 || epsrel < 0.0) return ST_GSL_EBADTOL;
  if (epsabs <= 0 && epsrel < DBL_EPSILON) return ST_GSL_EBADTOL;
  if (wsize < 3) return ST_GSL_EDOM;
  bool bad = false;
  auto feval = [&](double x) {
    int s1 = 0;
    const double v = cl_integrand(c, pref, pref + MAX_SUB, x, s1);
    if (s1 && !st) st = s1;
    if (!isfinite(v)) bad = true;
    return v;
  };
  auto settle = [&]() {  // every lane leaves with the same status / bad flag
    st = warp_first_nonzero(st);
    bad = __any_sync(0xffffffffu, bad);
    __syncwarp();
  };
  double nc, ncdiff, temp;
  CqIval* iv = &ivals[0];
  double m = (a + b) / 2, h = (b - a) / 2;
  for (int i = lane; i <= n[3]; i += 32) iv->fx[i] = feval(m + xi[i] * h);
  nev += 33;
  settle();
  cq_vinvfx(iv->fx, &iv->c[idx[0]], 0, lane);
  cq_vinvfx(iv->fx, &iv->c[idx[3]], 3, lane);
  cq_vinvfx(iv->fx, &iv->c[idx[2]], 2, lane);
  nc = 0.0;
  for (int i = n[2] + 1; i <= n[3]; i++) { temp = iv->c[idx[3] + i]; nc += temp * temp; }
  ncdiff = nc;
  for (int i = 0; i <= n[2]; i++) {
    temp = iv->c[idx[2] + i] - iv->c[idx[3] + i];
    ncdiff += temp * temp;
    nc += iv->c[idx[3] + i] * iv->c[idx[3] + i];
  }
  ncdiff = sqrt(ncdiff);
  nc = sqrt(nc);
  {
    const double ig0 = 2 * h * iv->c[idx[3]] * w;
    double er0 = ncdiff * 2 * h;
    if (ncdiff / nc > 0.1 && er0 < 2 * h * nc) er0 = 2 * h * nc;
    __syncwarp();
    if (lane == 0) {
      iv->a = a; iv->b = b; iv->depth = 3; iv->rdepth = 1; iv->ndiv = 0;
      iv->igral = ig0; iv->err = er0;
    }
  }
  for (int i = lane; i < wsize; i += 32) heap[i] = i;
  __syncwarp();
  double igral = iv->igral, err = iv->err, igral_final = 0.0, err_final = 0.0;
  int nivals = 1, status = 0;

  while (!bad && !st && nivals > 0 && err > 0.0 && !(err <= fabs(igral) * epsrel || err <= epsabs) &&
         !(err_final > fabs(igral) * epsrel || err - err_final < fabs(igral) * epsrel)) {
    int split, d;
    iv = &ivals[heap[0]];  // the interval with the largest error
    m = (iv->a + iv->b) / 2;
    h = (iv->b - iv->a) / 2;
    double iv_err = iv->err, iv_igral = iv->igral;
    if (iv->depth < 3) {  // raise the degree first
      d = iv->depth + 1;
      const int nnew = 16 / skip[d];  // samples at i = skip, 3 skip, 5 skip, ...
      if (lane < nnew) {
        const int i = skip[d] * (2 * lane + 1);
        iv->fx[i] = feval(m + xi[i] * h);
      }
      nev += nnew;
      settle();
      cq_vinvfx(iv->fx, &iv->c[idx[d]], d, lane);
      nc = 0.0;
      for (int i = n[d - 1] + 1; i <= n[d]; i++) { temp = iv->c[idx[d] + i]; nc += temp * temp; }
      ncdiff = nc;
      for (int i = 0; i <= n[d - 1]; i++) {
        temp = iv->c[idx[d - 1] + i] - iv->c[idx[d] + i];
        ncdiff += temp * temp;
        nc += iv->c[idx[d] + i] * iv->c[idx[d] + i];
      }
      ncdiff = 2 * h * sqrt(ncdiff);
      nc = sqrt(nc);
      iv_err = ncdiff;
      iv_igral = 2 * h * w * iv->c[idx[d]];
      __syncwarp();
      if (lane == 0) { iv->depth = d; iv->err = iv_err; iv->igral = iv_igral; }
      __syncwarp();
      split = (nc > 0 && ncdiff / nc > 0.1);
    } else
      split = 1;  // maximum degree reached

    if ((m + h * xi[0]) >= (m + h * xi[1]) || (m + h * xi[31]) >= (m + h * xi[32]) ||
        iv_err < fabs(iv_igral) * DBL_EPSILON * 10) {
      // drop the interval: keep its contribution
      err_final += iv_err;
      igral_final += iv_igral;
      if (lane == 0) {
        const int t = heap[nivals - 1]; heap[nivals - 1] = heap[0]; heap[0] = t;
        cq_sift_down(ivals, heap, 0, nivals - 1);
      }
      nivals--;
      __syncwarp();
    } else if (split) {
      d = iv->depth;
      CqIval* ch[2] = {&ivals[heap[nivals]], &ivals[heap[nivals + 1]]};
      // the six new samples of both children in one pass (i = 8, 16, 24 of each)
      if (lane < 6) {
        const int side = lane / 3, i = 8 * (lane % 3 + 1);
        const double ca = side ? m : iv->a, cb = side ? iv->b : m;
        ch[side]->fx[i] = feval((ca + cb) / 2 + xi[i] * h / 2);
      } else if (lane == 6) { ch[0]->fx[0] = iv->fx[0]; ch[0]->fx[32] = iv->fx[16]; }
      else if (lane == 7) { ch[1]->fx[0] = iv->fx[16]; ch[1]->fx[32] = iv->fx[32]; }
      nev += 6;
      settle();
      bool diverged = false;
      for (int side = 0; side < 2; ++side) {
        CqIval* ic = ch[side];
        const double* T = side == 0 ? CQ_TLEFT : CQ_TRIGHT;
        cq_vinvfx(ic->fx, ic->c, 0, lane);
        for (int i = lane; i <= n[d]; i += 32) {
          double acc = 0.0;
          for (int j = i; j <= n[d]; ++j) acc += T[i * 33 + j] * iv->c[idx[d] + j];
          ic->c[idx[d] + i] = acc;
        }
        __syncwarp();
        ncdiff = 0.0;
        for (int i = 0; i <= n[0]; i++) { temp = ic->c[i] - ic->c[idx[d] + i]; ncdiff += temp * temp; }
        for (int i = n[0] + 1; i <= n[d]; i++) { temp = ic->c[idx[d] + i]; ncdiff += temp * temp; }
        ncdiff = sqrt(ncdiff);
        const int ndiv = iv->ndiv + (fabs(iv->c[0]) > 0 && ic->c[0] / iv->c[0] > 2);
        const int rdepth = iv->rdepth + 1;
        const double cig = h * w * ic->c[0];
        __syncwarp();
        if (lane == 0) {
          ic->a = side ? m : iv->a; ic->b = side ? iv->b : m;
          ic->depth = 0; ic->rdepth = rdepth; ic->ndiv = ndiv;
          ic->err = ncdiff * h; ic->igral = cig;
        }
        __syncwarp();
        if (ndiv > ndiv_max && 2 * ndiv > rdepth) { diverged = true; break; }
      }
      if (diverged) {
        result = (igral >= 0) ? CUDART_INF : -CUDART_INF;
        return ST_GSL_EDIVERGE;
      }
      nivals += 2;
      // the parent leaves the heap; the right child takes its place, the left one is the last entry
      if (lane == 0) {
        int t = heap[nivals - 1]; heap[nivals - 1] = heap[0]; heap[0] = t;
        cq_sift_down(ivals, heap, 0, nivals - 2);
        int i = nivals - 2;
        while (i > 0) {
          const int j = (i - 1) / 2;
          if (ivals[heap[j]].err < ivals[heap[i]].err) {
            t = heap[j]; heap[j] = heap[i]; heap[i] = t;
            i = j;
          } else
            break;
        }
      }
      nivals--;
      __syncwarp();
    } else {
      if (lane == 0) cq_sift_down(ivals, heap, 0, nivals);
      __syncwarp();
    }
    // heap about to overflow: retire the last intervals
    while (nivals > wsize - 2) {
      const CqIval* il = &ivals[heap[nivals - 1]];
      err_final += il->err;
      igral_final += il->igral;
      nivals--;
    }
    igral = igral_final;
    err = err_final;
    for (int i = 0; i < nivals; i++) { igral += ivals[heap[i]].igral; err += ivals[heap[i]].err; }
  }
  result = igral;
  if (bad && !st) return ST_GSL_EFAILED;
  return status;
}

// Tasks the QAG kernel handed over (list[0] = count, list[1] = next, entries from list[2]).
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
limber_cquad_kernel(const LimberProblem P, unsigned long long* list, CqIval* pool, int* heaps) {
  __shared__ double s_pref[WARPS_PER_CTA][2 * MAX_SUB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * WARPS_PER_CTA + warp;
  const int wsize = P.par.N_ITERATION;
  CqIval* ivals = pool + (size_t)gw * wsize;
  int* heap = heaps + (size_t)gw * wsize;
  double* pref = s_pref[warp];
  const unsigned long long count = min(list[0], CQ_LIST_CAP);
  for (;;) {
    unsigned long long i = 0;
    if (lane == 0) i = atomicAdd(&list[1], 1ull);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= count) break;
    const long long t = (long long)list[2 + i];
    TaskCtx c;
    int ic, ip, il;
    setup_task(P, t, c, ic, ip, il);
    double lkmin, lkmax;
    get_k_interval(c, P.par, lkmin, lkmax);
    __syncwarp();
    setup_prefactors(c, pref, lane);
    int st = 0;
    unsigned long long nev = 0;
    double result = 0.0;
    const int gsl = cquad_warp(c, pref, lkmin, lkmax, 0.0, P.par.LIMBER_EPSREL, wsize, ivals, heap, lane, result, st, nev);
    if (P.counters && lane == 0) atomicAdd(P.counters, nev);
    finish_task(P, ic, ip, il, result, st ? st : gsl, lane);
    __syncwarp();
  }
}

// integ_cls_limber_qag_quad (src/ccl_cls.c:227-263): gsl_integration_qag(F, lkmin, lkmax, 0,
// epsrel, limit, GSL_INTEG_GAUSS41).  Persistent warps pull tasks from a counter; the
// interval list (a, b, r, e) of the warp lives in its slice of a global workspace.
#ifndef CCLB_QAG_MINB
#define CCLB_QAG_MINB 5
#endif
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CCLB_QAG_MINB)
limber_qag_kernel(const LimberProblem P, long long ntasks, double* ws, unsigned long long* next_task,
                  unsigned long long* cq_list, const unsigned long long* list, unsigned long long list_cap) {
  __shared__ double s_fv[WARPS_PER_CTA][2 * 41];
  __shared__ double s_pref[WARPS_PER_CTA][2 * MAX_SUB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int limit = P.par.N_ITERATION;
  double* alist = ws + ((size_t)blockIdx.x * WARPS_PER_CTA + warp) * 4 * (size_t)limit;
  double* blist = alist + limit;
  double* rlist = blist + limit;
  double* elist = rlist + limit;
  double* fv = s_fv[warp];
  double* pref = s_pref[warp];
  // list mode (redo list of the grid-sharing kernel): list[0] = count, tasks from list[1]; an overflowed
  // list means every task is recomputed
  const bool listed = list != nullptr && list[0] <= list_cap;
  if (list != nullptr) ntasks = listed ? (long long)list[0] : ntasks;
  for (;;) {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(next_task, 1ull);
    t = __shfl_sync(0xffffffffu, t, 0);
    if ((long long)t >= ntasks) break;
    if (listed) t = list[1 + t];
    TaskCtx c;
    int ic, ip, il;
    setup_task(P, (long long)t, c, ic, ip, il);
    if (!c.cs->has_achi) {
      finish_task(P, ic, ip, il, 0.0, ST_ERR_DISTANCES_INIT, lane);
      continue;
    }
    double lkmin, lkmax;
    get_k_interval(c, P.par, lkmin, lkmax);
    __syncwarp();
    setup_prefactors(c, pref, lane);
    const double epsabs = 0.0, epsrel = P.par.LIMBER_EPSREL;
    int st = 0, gsl = 0;
    unsigned long long nev = 41;
    double result = 0.0;
    if (epsabs <= 0 && (epsrel < 50 * DBL_EPSILON || epsrel < 0.5e-28)) {
      gsl = ST_GSL_EBADTOL;
    } else if (P.par.force_cquad) {
      gsl = ST_GSL_EROUND;  // tests of the fallback branch: skip QAG
      nev = 0;
    } else {
      double r0[1], e0[1], ab0[1], as0[1];
      const double pa0[1] = {lkmin}, pb0[1] = {lkmax};
      qk41_warp<1>(c, pref, pa0, pb0, fv, lane, r0, e0, ab0, as0, st);
      double tolerance = fmax(epsabs, epsrel * fabs(r0[0]));
      const double round_off = 50 * DBL_EPSILON * ab0[0];
      result = r0[0];
      if (e0[0] <= round_off && e0[0] > tolerance) gsl = ST_GSL_EROUND;
      else if ((e0[0] <= tolerance && e0[0] != as0[0]) || e0[0] == 0.0) gsl = 0;
      else if (limit == 1) gsl = ST_GSL_EMAXITER;
      else {
        if (lane == 0) { alist[0] = lkmin; blist[0] = lkmax; rlist[0] = r0[0]; elist[0] = e0[0]; }
        __syncwarp();
        int size = 1;
        double area = r0[0], errsum = e0[0];
        int iteration = 1, roundoff_type1 = 0, roundoff_type2 = 0, error_type = 0;
        do {
          // retrieve(): interval with the largest error estimate
          double emax = -1.0;
          int imax = 0;
          for (int i = lane; i < size; i += 32) {
            const double e = elist[i];
            if (e > emax) { emax = e; imax = i; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double eo = __shfl_xor_sync(0xffffffffu, emax, o);
            const int io = __shfl_xor_sync(0xffffffffu, imax, o);
            if (eo > emax || (eo == emax && io < imax)) { emax = eo; imax = io; }
          }
          const double a_i = alist[imax], b_i = blist[imax], r_i = rlist[imax], e_i = elist[imax];
          const double a1 = a_i, b1 = 0.5 * (a_i + b_i), a2 = b1, b2 = b_i;
          double ar[2], er[2], rab[2], ras[2];
          const double pa[2] = {a1, a2}, pb[2] = {b1, b2};
          qk41_warp<2>(c, pref, pa, pb, fv, lane, ar, er, rab, ras, st);
          nev += 82;
          const double area12 = ar[0] + ar[1];
          const double error12 = er[0] + er[1];
          errsum += (error12 - e_i);
          area += area12 - r_i;
          if (ras[0] != er[0] && ras[1] != er[1]) {
            const double delta = r_i - area12;
            if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) roundoff_type1++;
            if (iteration >= 10 && error12 > e_i) roundoff_type2++;
          }
          tolerance = fmax(epsabs, epsrel * fabs(area));
          if (errsum > tolerance) {
            if (roundoff_type1 >= 6 || roundoff_type2 >= 20) error_type = 2;
            if (subinterval_too_small(a1, a2, b2)) error_type = 3;
          }
          // update(): larger-error half stays at imax, the other is appended
          __syncwarp();
          if (lane == 0) {
            if (er[1] > er[0]) {
              alist[imax] = a2; rlist[imax] = ar[1]; elist[imax] = er[1];
              alist[size] = a1; blist[size] = b1; rlist[size] = ar[0]; elist[size] = er[0];
            } else {
              blist[imax] = b1; rlist[imax] = ar[0]; elist[imax] = er[0];
              alist[size] = a2; blist[size] = b2; rlist[size] = ar[1]; elist[size] = er[1];
            }
          }
          __syncwarp();
          size++;
          iteration++;
          if (st) break;  // integrand status poisons the C_ell anyway
        } while (iteration < limit && !error_type && errsum > tolerance);
        double sum = 0.0;
        for (int i = lane; i < size; i += 32) sum += rlist[i];
        result = warp_sum(sum);
        if (errsum <= tolerance) gsl = 0;
        else if (error_type == 2) gsl = ST_GSL_EROUND;
        else if (error_type == 3) gsl = ST_GSL_ESING;
        else if (iteration == limit) gsl = ST_GSL_EMAXITER;
        else gsl = ST_GSL_EFAILED;
      }
    }
    if (P.counters && lane == 0) atomicAdd(P.counters, nev);
    // GSL_EROUND switches the reference to CQUAD (src/ccl_cls.c:245-259): the task is handed to
    // limber_cquad_kernel, which runs after this kernel in the same stream
    if (gsl == ST_GSL_EROUND && !st && cq_list) {
      unsigned long long at = CQ_LIST_CAP;
      if (lane == 0) {
        at = atomicAdd(&cq_list[0], 1ull);
        if (at < CQ_LIST_CAP) cq_list[2 + at] = t;
      }
      at = __shfl_sync(0xffffffffu, at, 0);
      if (at < CQ_LIST_CAP) continue;
    }
    finish_task(P, ic, ip, il, result, st ? st : gsl, lane);
  }
}

// ---- fp64 vector-pipe peak: 16 independent DFMA chains per thread ------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i * 1e-3 + threadIdx.x * 1e-6;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456) out[0] = s;  // keep the chains alive
}

double measure_fp64_peak_tflops(cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) return -1.0;
  const int iters = 1 << 15, blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dfma_peak_kernel<<<blocks, threads, 0, stream>>>(d, 1 << 10, 1.0);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, stream);
    dfma_peak_kernel<<<blocks, threads, 0, stream>>>(d, iters, 1.0 + rep);
    cudaEventRecord(e1, stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * threads;
    best = fmax(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}

// workspace layout: [QAG interval lists | task counter (64 B) | CQUAD task list | CQUAD interval pool | heaps]
static size_t up256(size_t b) { return (b + 255) / 256 * 256; }
static size_t qag_region_bytes(size_t warps, int n_iteration) {
  return up256(warps * 4 * (size_t)n_iteration * sizeof(double)) + 256;
}
static size_t cq_list_bytes() { return up256((2 + CQ_LIST_CAP) * sizeof(unsigned long long)); }
static size_t cq_pool_bytes(int n_iteration) { return up256((size_t)CQ_WARPS * n_iteration * sizeof(CqIval)); }
static size_t cq_heap_bytes(int n_iteration) { return up256((size_t)CQ_WARPS * n_iteration * sizeof(int)); }

size_t limber_qag_workspace_bytes(int n_iteration) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t warps = (size_t)sms * 8 * WARPS_PER_CTA;
  return qag_region_bytes(warps, n_iteration) + cq_list_bytes() + cq_pool_bytes(n_iteration) + cq_heap_bytes(n_iteration);
}

int limber_nk_cap(const LimberParams& par) {
  // upper bound of nk over all (pair, ell): lkmax - lkmin <= ln(K_MAX / K_MIN)
  const double span = log(par.K_MAX / par.K_MIN);
  int cap = (int)(fmax(span / par.DLOGK_INTEGRATION + 0.5, 1.0)) + 2;
  return (cap + 3) & ~3;
}

void launch_limber_redo(const LimberProblem& prob, const FastPlan& plan, cudaStream_t stream) {
  const long long ntasks = (long long)prob.ncosmo * prob.npair * prob.nl;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = (size_t)WARPS_PER_CTA * (prob.par.nk_cap + 2 * MAX_SUB) * sizeof(double);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(limber_spline_redo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  limber_spline_redo_kernel<<<sms * 4, WARPS_PER_CTA * 32, smem, stream>>>(prob, ntasks, plan.redo, plan.redo_cap);
}

// The CQUAD hand-over list inside the QAG workspace (zeroed here: call once per angular_cl call, before the
// integrators that append to it).
void limber_qag_cq_list(double* qag_ws, int n_iteration, unsigned long long** list, unsigned long long* cap,
                        cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t o_next = qag_region_bytes((size_t)sms * 8 * WARPS_PER_CTA, n_iteration) - 256;
  char* wsb = reinterpret_cast<char*>(qag_ws);
  *list = reinterpret_cast<unsigned long long*>(wsb + o_next + 256);
  *cap = CQ_LIST_CAP;
  cudaMemsetAsync(*list, 0, 2 * sizeof(unsigned long long), stream);
}

// General QAG kernel over every task (list == nullptr) or over a redo list, then the CQUAD fallback over the
// hand-over list (which the caller zeroed with limber_qag_cq_list before the first integrator ran).
void launch_limber_qag_list(const LimberProblem& prob, double* qag_ws, size_t qag_ws_bytes,
                            const unsigned long long* list, unsigned long long cap, cudaStream_t stream) {
  const long long ntasks = (long long)prob.ncosmo * prob.npair * prob.nl;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!g_gk_uploaded[dev & 63]) {
    cudaMemcpyToSymbol(c_xgk, GK41_XGK, sizeof(double) * 21);
    cudaMemcpyToSymbol(c_wgk, GK41_WGK, sizeof(double) * 21);
    cudaMemcpyToSymbol(c_wg, GK41_WG, sizeof(double) * 10);
    g_gk_uploaded[dev & 63] = true;
  }
  const int nblk = sms * 8;
  const int nit = prob.par.N_ITERATION;
  const size_t o_next = qag_region_bytes((size_t)nblk * WARPS_PER_CTA, nit) - 256;
  const size_t o_list = o_next + 256, o_pool = o_list + cq_list_bytes(), o_heap = o_pool + cq_pool_bytes(nit);
  (void)qag_ws_bytes;
  char* wsb = reinterpret_cast<char*>(qag_ws);
  unsigned long long* next = reinterpret_cast<unsigned long long*>(wsb + o_next);
  unsigned long long* cq_list = reinterpret_cast<unsigned long long*>(wsb + o_list);
  cudaMemsetAsync(next, 0, sizeof(unsigned long long), stream);
  limber_qag_kernel<<<nblk, WARPS_PER_CTA * 32, 0, stream>>>(prob, ntasks, qag_ws, next, cq_list, list, cap);
  // the reference's fallback for GSL_EROUND; its warps exit at once when the list is empty
  limber_cquad_kernel<<<CQ_WARPS / WARPS_PER_CTA, WARPS_PER_CTA * 32, 0, stream>>>(
      prob, cq_list, reinterpret_cast<CqIval*>(wsb + o_pool), reinterpret_cast<int*>(wsb + o_heap));
}

cudaError_t launch_limber(const LimberProblem& prob, int method, double* qag_ws, size_t qag_ws_bytes,
                          cudaStream_t stream) {
  const long long ntasks = (long long)prob.ncosmo * prob.npair * prob.nl;
  if (ntasks == 0) return cudaSuccess;
  if (method == INTEG_SPLINE) {
    const size_t smem = (size_t)WARPS_PER_CTA * (prob.par.nk_cap + 2 * MAX_SUB) * sizeof(double);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(limber_spline_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long nblk = (ntasks + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (nblk > 0x7fffffffLL) return cudaErrorInvalidValue;
    limber_spline_kernel<<<(unsigned)nblk, WARPS_PER_CTA * 32, smem, stream>>>(prob, ntasks);
  } else {
    if (limber_qag_workspace_bytes(prob.par.N_ITERATION) > qag_ws_bytes) return cudaErrorInvalidValue;
    unsigned long long* cq_list = nullptr;
    unsigned long long cq_cap = 0;
    limber_qag_cq_list(qag_ws, prob.par.N_ITERATION, &cq_list, &cq_cap, stream);
    launch_limber_qag_list(prob, qag_ws, qag_ws_bytes, nullptr, 0, stream);
  }
  return cudaGetLastError();
}

}  // namespace cclb
