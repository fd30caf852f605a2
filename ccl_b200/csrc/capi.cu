// C ABI of the B200 Limber path (include/ccl_b200.h).  Host-side bookkeeping only: flat input
// arrays are staged into one arena, uploaded with a single copy, and every derived table
// (Akima / natural-cspline coefficients, bicubic derivative planes, index LUTs) is built by
// device kernels.  No numerical work of the hot path happens on the CPU.
#include <sched.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/ccl_b200.h"
#include "tables.h"

using namespace cclb;

namespace {

thread_local std::string g_err;
thread_local int g_device = -1;

void set_err(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

bool cuda_ok(cudaError_t e, const char* what, int* status) {
  if (e == cudaSuccess) return true;
  set_err("ccl_b200: CUDA failure in %s: %s", what, cudaGetErrorString(e));
  if (status) *status = CCL_B200_ERROR_CUDA;
  return false;
}

bool use_device(int* status) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_err("ccl_b200: no CUDA device available (%s); this library has no CPU fallback",
            e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (status) *status = CCL_B200_ERROR_CUDA;
    return false;
  }
  if (g_device >= 0) return cuda_ok(cudaSetDevice(g_device), "cudaSetDevice", status);
  return true;
}

constexpr size_t ALIGN_D = 4;  // raw region alignment in doubles (32 B)

// Helper threads of the host-side staging / node checks: at most 16, at most the cores this process may run on
// (sched_getaffinity: a launcher that pins ranks to core sets is respected), divided by the ranks that share
// the node (LOCAL_WORLD_SIZE, exported by torchrun) so that 8 ranks do not start 128 threads on 32 cores.
// CCL_B200_HOST_THREADS overrides.
unsigned host_threads() {
  static const unsigned n = [] {
    if (const char* e = getenv("CCL_B200_HOST_THREADS")) return (unsigned)std::max(1, atoi(e));
    unsigned cores = std::max(1u, std::thread::hardware_concurrency());
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = std::max(1, CPU_COUNT(&set));
    unsigned ranks = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = (unsigned)std::max(1, atoi(e));
    // a rank pinned to its own cores keeps them all; otherwise share the node's cores evenly
    if (cores >= std::thread::hardware_concurrency()) cores = std::max(1u, cores / ranks);
    return std::min(16u, cores);
  }();
  return n;
}

// ---- plans: offsets into the arena, turned into device descriptors at commit ----
constexpr size_t NO_LAND = (size_t)-1;

struct S1Plan {
  bool present = false;
  int kind = 0;  // 0 akima, 1 natural cspline, 2 bicubic axis (nodes + LUT)
  int n = 0, nlut = 0, uniform = 0;
  size_t x_off = 0, y_off = 0;               // raw region, in doubles
  size_t co_off = 0, lut_off = 0, scr_off = 0;  // derived region, in bytes
  double x0 = 0, xn = 0;
  // tables built on the device (ccl_b200_tables) are read in place by the preparation kernels
  const double* x_dev = nullptr;
  const double* y_dev = nullptr;
  // stacked setters on pinned host arrays: the whole stacked array is copied by ONE DMA into the
  // landing region at commit and the table is read there (offset in doubles; NO_LAND = staged)
  size_t x_land = NO_LAND, y_land = NO_LAND;
  int check_id = -1;  // stacked setters: node checks deferred to TableSet::run_checks (host threads)
};

struct F2DPlan {
  F2D flags;  // pointer members unused
  S1Plan fk, fa, gk, ga;
  size_t z_off = 0, tab_off = 0, scr_off = 0;
  int nk = 0, na = 0;
  const double* z_dev = nullptr;
  size_t z_land = NO_LAND;
};

// Where the arrays of one stacked-setter slot sit in the landing region (NO_LAND: stage through raw)
struct Lands {
  size_t chi = NO_LAND, w = NO_LAND, a = NO_LAND, lk = NO_LAND, fka = NO_LAND, fk = NO_LAND, fa = NO_LAND;
};

struct TracerPlan {
  int der_bessel = 0, der_angles = 0;
  double chi_min = 0, chi_max = 1e15;
  S1Plan kernel;
  bool has_transfer = false;
  F2DPlan transfer;
  bool const_transfer = false;  // factorised, a-only, linear, all samples equal
  double const_transfer_value = 0;
};

struct CosmoPlan {
  S1Plan achi, growth;
  double chi_max_nokernel = 1e15;
  bool has_psp = false;
  F2DPlan psp;
  std::vector<TracerPlan> tracers;
};

struct HostBuf {  // growable host staging; pinned once it is large
  double* p = nullptr;
  size_t size = 0, cap = 0;
  bool pinned = false;
  ~HostBuf() { release(); }
  void release() {
    if (p) {
      if (pinned) cudaFreeHost(p); else free(p);
    }
    p = nullptr; size = cap = 0;
  }
  bool reserve(size_t n) {
    if (n <= cap) return true;
    size_t ncap = std::max(n, cap * 2);
    bool want_pinned = ncap * sizeof(double) >= (size_t(1) << 20);
    double* q = nullptr;
    if (want_pinned) {
      if (cudaHostAlloc((void**)&q, ncap * sizeof(double), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        want_pinned = false;
      }
    }
    if (!want_pinned) q = (double*)malloc(ncap * sizeof(double));
    if (!q) return false;
    if (p) memcpy(q, p, size * sizeof(double));
    if (p) { if (pinned) cudaFreeHost(p); else free(p); }
    p = q; cap = ncap; pinned = want_pinned;
    return true;
  }
  // With defer set, put() only reserves the destination; flush() then copies every pending block
  // with several host threads (the stacked setters stage GBs of tables per call).
  struct Pending { size_t off; const double* src; size_t n; };
  bool defer = false;
  std::vector<Pending> pending;
  size_t put(const double* src, size_t n) {
    size_t off = (size + ALIGN_D - 1) / ALIGN_D * ALIGN_D;
    if (!reserve(off + n)) return (size_t)-1;
    if (off > size) memset(p + size, 0, (off - size) * sizeof(double));
    if (defer) pending.push_back(Pending{off, src, n});
    else memcpy(p + off, src, n * sizeof(double));
    size = off + n;
    return off;
  }
  void flush() {
    if (pending.empty()) return;
    size_t total = 0;
    for (const Pending& q : pending) total += q.n;
    unsigned nt = host_threads();
    if (total * sizeof(double) < (size_t(8) << 20)) nt = 1;
    auto work = [&](unsigned t) {
      for (size_t i = t; i < pending.size(); i += nt)
        memcpy(p + pending[i].off, pending[i].src, pending[i].n * sizeof(double));
    };
    if (nt == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
      for (auto& x : th) x.join();
    }
    pending.clear();
  }
};

struct DevBuf {
  char* p = nullptr;
  size_t cap = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  bool reserve(size_t n, int* status) {
    if (n <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    if (!cuda_ok(cudaMalloc((void**)&p, n), "cudaMalloc(arena)", status)) return false;
    cap = n;
    return true;
  }
};

// One arena with the table packs of `cosmos.size()` cosmologies.
struct TableSet {
  int device = 0;
  HostBuf raw;
  size_t derived = 0;  // bytes planned in the derived region
  std::vector<CosmoPlan> cosmos;
  DevBuf dev;          // [raw | derived | descriptors]
  DevBuf jobs;         // prep job arrays
  // Landing region: stacked arrays that live in pinned host memory are not staged; commit() copies
  // each of them with one DMA straight from the caller's buffer (which must stay valid until
  // commit returns).  Small shared grids are kept in `owned` so that they take the same route.
  struct DirectCopy { size_t off; const double* src; size_t n; bool done; };
  DevBuf landing;
  size_t land_size = 0;  // doubles
  std::vector<DirectCopy> direct;
  std::deque<std::vector<double>> owned;

  static bool is_pinned(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
  }
  size_t land_put(const double* src, size_t n) {
    const size_t off = (land_size + ALIGN_D - 1) / ALIGN_D * ALIGN_D;
    direct.push_back(DirectCopy{off, src, n, false});
    land_size = off + n;
    return off;
  }
  size_t land_put_owned(const double* src, size_t n) {
    owned.emplace_back(src, src + n);
    return land_put(owned.back().data(), n);
  }
  // landing offset of a stacked array if it is pinned, else NO_LAND
  size_t land_if_pinned(const double* src, size_t n) {
    static const bool off = getenv("CCL_B200_NO_DIRECT") != nullptr;
    return (src && n && !off && is_pinned(src)) ? land_put(src, n) : NO_LAND;
  }
  size_t raw_bytes = 0, desc_off = 0;
  Cosmo* d_cosmos = nullptr;
  F2D* d_psp = nullptr;
  Tracer* d_tracers = nullptr;
  std::vector<Cosmo> h_cosmos;
  std::vector<F2D> h_psp;
  std::vector<Tracer> h_tracers;
  std::vector<size_t> tracer_base;  // first tracer of cosmology i in h_tracers
  bool committed = false;
  unsigned long long commit_gen = 0;  // bumped by every commit(): invalidates plans derived from the tables
  unsigned long long last_h2d = 0;

  size_t derive(size_t bytes) {
    size_t off = (derived + 63) / 64 * 64;
    derived = off + bytes;
    return off;
  }

  // deferred node checks of the stacked setters (strictly increasing nodes; quasi-uniform grid)
  struct NodeCheck { size_t x_off; const double* xsrc; int n; int uniform; int ok; };
  std::vector<NodeCheck> checks;
  size_t checks_done = 0;
  std::unordered_map<const double*, int> shared_checks;  // source pointer -> check, within one setter call

  void run_checks() {
    const size_t lo = checks_done, hi = checks.size();
    if (hi == lo) { shared_checks.clear(); return; }
    size_t nodes = 0;
    for (size_t c = lo; c < hi; ++c) nodes += (size_t)checks[c].n;
    // a thread is worth spawning per ~256k nodes (a stacked call on a few hundred slots stays inline)
    unsigned nt = host_threads();
    nt = (unsigned)std::min<size_t>(nt, nodes / (size_t(1) << 18) + 1);
    auto work = [&](unsigned t) {
      for (size_t c = lo + t; c < hi; c += nt) {
        NodeCheck& k = checks[c];
        const double* x = k.xsrc ? k.xsrc : raw.p + k.x_off;
        k.ok = 1;
        for (int i = 1; i < k.n; ++i)
          if (!(x[i] > x[i - 1])) { k.ok = 0; break; }
        k.uniform = k.ok ? quasi_uniform(x, k.n) : 0;
      }
    };
    if (nt == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
      for (auto& x : th) x.join();
    }
    checks_done = hi;
    shared_checks.clear();
  }

  bool resolve(S1Plan& s, int* status) {
    if (!s.present || s.check_id < 0) return true;
    const NodeCheck& k = checks[s.check_id];
    s.check_id = -1;
    if (!k.ok) {
      set_err("ccl_b200: spline nodes must be strictly increasing");
      *status = CCL_B200_ERROR_SPLINE;
      return false;
    }
    s.uniform = k.uniform;
    return true;
  }
  bool resolve_f2d(F2DPlan& f, int* status) {
    return resolve(f.fk, status) && resolve(f.fa, status) && resolve(f.gk, status) && resolve(f.ga, status);
  }
  bool resolve_all(int* status) {
    for (CosmoPlan& c : cosmos) {
      if (!resolve(c.achi, status) || !resolve(c.growth, status)) return false;
      if (c.has_psp && !resolve_f2d(c.psp, status)) return false;
      for (TracerPlan& t : c.tracers) {
        if (!resolve(t.kernel, status)) return false;
        if (t.has_transfer && !resolve_f2d(t.transfer, status)) return false;
      }
    }
    return true;
  }

  // interval records of an n-node spline: {x0, x1}[n-1] then, on a 32-byte boundary, {y, b, c, d}[n-1]
  static size_t rec_x_bytes(int n) { return ((size_t)(n - 1) * sizeof(double2) + 31) / 32 * 32; }
  static size_t rec_bytes(int n) { return rec_x_bytes(n) + (size_t)(n - 1) * sizeof(double4); }

  static int lut_size(int n) { return std::min(std::max(n > 1000 ? 2 * n : 4 * n, 64), 1 << 15); }

  bool plan_s1(S1Plan& s, int kind, int n, const double* x, const double* y, int* status,
               size_t x_land = NO_LAND, size_t y_land = NO_LAND) {
    s = S1Plan();
    s.present = true; s.kind = kind; s.n = n;
    s.x_land = x_land;
    if (x_land == NO_LAND) s.x_off = raw.put(x, n);
    if (y) {
      s.y_land = y_land;
      if (y_land == NO_LAND) s.y_off = raw.put(y, n);
    }
    if (s.x_off == (size_t)-1 || s.y_off == (size_t)-1) { *status = CCL_B200_ERROR_MEMORY; return false; }
    s.x0 = x[0]; s.xn = x[n - 1];
    if (raw.defer) {
      // stacked setters: the O(n) node checks run on several host threads after staging
      // (a grid shared by the slots of one stacked call is checked once)
      auto seen = x_land == NO_LAND ? shared_checks.end() : shared_checks.find(x);
      if (seen != shared_checks.end() && checks[seen->second].n == n) {
        s.check_id = seen->second;
      } else {
        s.check_id = (int)checks.size();
        checks.push_back(NodeCheck{s.x_off, x_land == NO_LAND ? nullptr : x, n, 0, 1});
        if (x_land != NO_LAND) shared_checks[x] = s.check_id;
      }
    } else {
      for (int i = 1; i < n; ++i)
        if (!(x[i] > x[i - 1])) {  // gsl_spline_init requires strictly increasing x
          set_err("ccl_b200: spline nodes must be strictly increasing");
          *status = CCL_B200_ERROR_SPLINE;
          return false;
        }
      // quasi-uniform grid: the analytic index guess is within +-1 interval of the true one
      s.uniform = quasi_uniform(x, n);
    }
    s.nlut = lut_size(n);
    s.lut_off = derive(sizeof(int) * s.nlut);
    if (kind != 2) s.co_off = derive(rec_bytes(n));
    else s.co_off = derive(sizeof(AxisRec) * (n - 1));
    if (kind == 1) s.scr_off = derive(sizeof(double) * 3 * n);
    return true;
  }

  // Same bookkeeping as plan_s1 for a table that already sits in device memory: the host only needs
  // its size, end nodes and whether its grid is quasi-uniform.
  void plan_s1_dev(S1Plan& s, int kind, int n, double x0, double xn, int uniform, const double* d_x,
                   const double* d_y) {
    s = S1Plan();
    s.present = true; s.kind = kind; s.n = n; s.x0 = x0; s.xn = xn; s.uniform = uniform;
    s.x_dev = d_x; s.y_dev = d_y;
    s.nlut = lut_size(n);
    s.lut_off = derive(sizeof(int) * s.nlut);
    if (kind != 2) s.co_off = derive(rec_bytes(n));
    else s.co_off = derive(sizeof(AxisRec) * (n - 1));
    if (kind == 1) s.scr_off = derive(sizeof(double) * 3 * n);
  }

  static int quasi_uniform(const double* x, int n) {
    const double step = (x[n - 1] - x[0]) / (n - 1.);
    for (int i = 0; i < n; ++i)
      if (fabs(x[i] - (x[0] + step * i)) > 0.3 * step) return 0;
    return 1;
  }

  // set_pk2d_new_from_arrays flags on a device-resident log P(k,a) table with shared grids
  void plan_psp_dev(F2DPlan& f, int nk, const double* h_lk, const double* d_lk, int na, const double* h_a,
                    const double* d_a, const double* d_z, int order_lok, int order_hik, int is_log) {
    memset(&f.flags, 0, sizeof(F2D));
    F2D& g = f.flags;
    g.extrap_order_lok = order_lok; g.extrap_order_hik = order_hik;
    g.extrap_linear_growth = GROWTH_CCL; g.is_log = is_log; g.growth_factor_0 = 0; g.growth_exponent = 2;
    g.lkmin = h_lk[0]; g.lkmax = h_lk[nk - 1]; g.amin = h_a[0]; g.amax = h_a[na - 1];
    plan_s1_dev(f.gk, 2, nk, h_lk[0], h_lk[nk - 1], quasi_uniform(h_lk, nk), d_lk, nullptr);
    plan_s1_dev(f.ga, 2, na, h_a[0], h_a[na - 1], quasi_uniform(h_a, na), d_a, nullptr);
    f.fk = S1Plan(); f.fa = S1Plan();
    f.nk = nk; f.na = na;
    f.z_dev = d_z;
    f.tab_off = derive(sizeof(double4) * (size_t)nk * na);
    f.scr_off = derive(sizeof(double) * 4 * (size_t)(nk + na));
    g.has_fka = 1;
  }

  // ccl_f2d_t_new (src/ccl_f2d.c:92-201)
  bool plan_f2d(F2DPlan& f, int na, const double* a_arr, int nk, const double* lk_arr,
                const double* fka_arr, const double* fk_arr, const double* fa_arr, int is_factorizable,
                int extrap_order_lok, int extrap_order_hik, int extrap_linear_growth, int is_fka_log,
                double growth_factor_0, int growth_exponent, int* status, const Lands& L = Lands()) {
    memset(&f.flags, 0, sizeof(F2D));
    f.z_dev = nullptr;
    f.z_land = NO_LAND;
    F2D& g = f.flags;
    is_factorizable = is_factorizable || (a_arr == nullptr) || (lk_arr == nullptr) || (fka_arr == nullptr);
    g.is_factorizable = is_factorizable;
    g.is_k_constant = ((lk_arr == nullptr) || ((fka_arr == nullptr) && (fk_arr == nullptr)));
    g.is_a_constant = ((a_arr == nullptr) || ((fka_arr == nullptr) && (fa_arr == nullptr)));
    g.extrap_order_lok = extrap_order_lok;
    g.extrap_order_hik = extrap_order_hik;
    g.extrap_linear_growth = extrap_linear_growth;
    g.is_log = is_fka_log;
    g.growth_factor_0 = growth_factor_0;
    g.growth_exponent = growth_exponent;
    if (!g.is_k_constant) { g.lkmin = lk_arr[0]; g.lkmax = lk_arr[nk - 1]; }
    if (!g.is_a_constant) { g.amin = a_arr[0]; g.amax = a_arr[na - 1]; }
    if ((extrap_order_lok > 2) || (extrap_order_lok < 0) || (extrap_order_hik > 2) || (extrap_order_hik < 0))
      *status = CCL_B200_ERROR_INCONSISTENT;
    if ((extrap_linear_growth != GROWTH_CCL) && (extrap_linear_growth != GROWTH_CONST) &&
        (extrap_linear_growth != GROWTH_NONE))
      *status = CCL_B200_ERROR_INCONSISTENT;
    if (*status) { set_err("ccl_b200: ccl_f2d_t_new(): inconsistent extrapolation settings"); return false; }
    if (g.is_factorizable) {
      if (!g.is_k_constant) {
        if (nk < 3) { *status = CCL_B200_ERROR_MEMORY; set_err("ccl_b200: cspline needs >= 3 nodes"); return false; }
        if (!plan_s1(f.fk, 1, nk, lk_arr, fk_arr, status, L.lk, L.fk)) return false;
        g.has_fk = 1;
      }
      if (!g.is_a_constant) {
        if (na < 3) { *status = CCL_B200_ERROR_MEMORY; set_err("ccl_b200: cspline needs >= 3 nodes"); return false; }
        if (!plan_s1(f.fa, 1, na, a_arr, fa_arr, status, L.a, L.fa)) return false;
        g.has_fa = 1;
      }
    } else if (!(g.is_k_constant || g.is_a_constant)) {
      if (nk < 4 || na < 4) {  // gsl_interp2d_bicubic min_size
        *status = CCL_B200_ERROR_MEMORY;
        set_err("ccl_b200: bicubic interpolation needs >= 4 nodes per axis");
        return false;
      }
      if (!plan_s1(f.gk, 2, nk, lk_arr, nullptr, status, L.lk)) return false;
      if (!plan_s1(f.ga, 2, na, a_arr, nullptr, status, L.a)) return false;
      f.nk = nk; f.na = na;
      f.z_land = L.fka;
      if (L.fka == NO_LAND) f.z_off = raw.put(fka_arr, (size_t)nk * na);
      if (f.z_off == (size_t)-1) { *status = CCL_B200_ERROR_MEMORY; return false; }
      f.tab_off = derive(sizeof(double4) * (size_t)nk * na);
      f.scr_off = derive(sizeof(double) * 4 * (size_t)(nk + na));
      g.has_fka = 1;
    }
    return true;
  }

  // ccl_cl_tracer_t_new (src/ccl_tracers.c:569-691)
  bool plan_tracer(TracerPlan& t, double chi_max_nokernel, int der_bessel, int der_angles, int n_w,
                   const double* chi_w, const double* w_w, int na_ka, const double* a_ka, int nk_ka,
                   const double* lk_ka, const double* fka_arr, const double* fk_arr, const double* fa_arr,
                   int is_fka_log, int is_factorizable, int extrap_order_lok, int extrap_order_hik,
                   int* status, const Lands& L = Lands()) {
    if ((der_angles < 0) || (der_angles > 2)) {
      *status = CCL_B200_ERROR_INCONSISTENT;
      set_err("ccl_tracers.c: ccl_cl_tracer_new(): der_angles must be between 0 and 2\n");
    }
    if ((der_bessel < -1) || (der_bessel > 2)) {
      *status = CCL_B200_ERROR_INCONSISTENT;
      set_err("ccl_tracers.c: ccl_cl_tracer_new(): der_bessel must be between -1 and 2\n");
    }
    if (*status) return false;
    t.der_angles = der_angles;
    t.der_bessel = der_bessel;
    t.chi_min = 0;
    t.chi_max = 1E15;
    const bool has_kernel = (n_w > 0) && (chi_w != nullptr) && (w_w != nullptr);
    if (has_kernel) {
      if (n_w < 5) {  // gsl_spline_alloc(akima) fails below 5 points -> CCL_ERROR_MEMORY
        *status = CCL_B200_ERROR_MEMORY;
        set_err("ccl_b200: tracer kernel needs >= 5 nodes (akima)");
        return false;
      }
      if (!plan_s1(t.kernel, 0, n_w, chi_w, w_w, status, L.chi, L.w)) return false;
      double w_max = fabs(w_w[0]);
      for (int i = 0; i < n_w; ++i)
        if (fabs(w_w[i]) >= w_max) w_max = fabs(w_w[i]);
      w_max *= 5E-4;  // CCL_FRAC_RELEVANT, include/ccl_tracers.h:13
      t.chi_min = chi_w[0];
      t.chi_max = chi_w[n_w - 1];
      for (int i = 0; i < n_w - 1; ++i)
        if (fabs(w_w[i + 1]) >= w_max) { t.chi_min = chi_w[i]; break; }
      for (int i = n_w - 1; i >= 1; --i)
        if (fabs(w_w[i - 1]) >= w_max) { t.chi_max = chi_w[i]; break; }
    } else {
      t.chi_min = 0;
      t.chi_max = chi_max_nokernel;
    }
    if ((fka_arr != nullptr) || (fk_arr != nullptr) || (fa_arr != nullptr)) {
      t.has_transfer = true;
      if (!plan_f2d(t.transfer, na_ka, a_ka, nk_ka, lk_ka, fka_arr, fk_arr, fa_arr, is_factorizable,
                    extrap_order_lok, extrap_order_hik, GROWTH_CONST, is_fka_log, 1.0, 0, status, L))
        return false;
      const F2D& g = t.transfer.flags;
      if (g.is_factorizable && !g.has_fk && g.has_fa && !g.is_log && fa_arr != nullptr) {
        bool same = true;
        for (int i = 1; i < na_ka && same; ++i) same = (fa_arr[i] == fa_arr[0]);
        t.const_transfer = same;
        t.const_transfer_value = fa_arr[0];
      }
    }
    return true;
  }

  void reset() {
    raw.size = 0;
    derived = 0;
    land_size = 0;
    direct.clear();
    owned.clear();
    checks.clear();
    shared_checks.clear();
    checks_done = 0;
    for (auto& c : cosmos) c = CosmoPlan();
    committed = false;
  }

  // ---- commit: upload + device preparation ----
  Spline1D mk_s1(const S1Plan& s, char* base_raw, char* base_der) const {
    Spline1D d;
    memset(&d, 0, sizeof(d));
    if (!s.present) return d;
    d.xr = reinterpret_cast<const double2*>(base_der + s.co_off);
    d.cr = reinterpret_cast<const double4*>(base_der + s.co_off + rec_x_bytes(s.n));
    d.lut = reinterpret_cast<const int*>(base_der + s.lut_off);
    d.n = s.n; d.nlut = s.nlut; d.uniform = s.uniform; d.x0 = s.x0; d.xn = s.xn;
    d.inv_step = (s.n - 1) / (s.xn - s.x0);
    d.inv_bin = s.nlut / (s.xn - s.x0);
    return d;
  }

  Axis mk_axis(const S1Plan& s, char* base_der) const {
    Axis d;
    memset(&d, 0, sizeof(d));
    if (!s.present) return d;
    d.rec = reinterpret_cast<const AxisRec*>(base_der + s.co_off);
    d.lut = reinterpret_cast<const int*>(base_der + s.lut_off);
    d.n = s.n; d.nlut = s.nlut; d.uniform = s.uniform; d.x0 = s.x0; d.xn = s.xn;
    d.inv_step = (s.n - 1) / (s.xn - s.x0);
    d.inv_bin = s.nlut / (s.xn - s.x0);
    return d;
  }

  F2D mk_f2d(const F2DPlan& f, char* br, char* bd) const {
    F2D d = f.flags;
    d.fk = mk_s1(f.fk, br, bd);
    d.fa = mk_s1(f.fa, br, bd);
    d.gk = mk_axis(f.gk, bd);
    d.ga = mk_axis(f.ga, bd);
    d.tab = d.has_fka ? reinterpret_cast<const double4*>(bd + f.tab_off) : nullptr;
    return d;
  }

  const double* land_base() const { return reinterpret_cast<const double*>(landing.p); }
  const double* xptr(const S1Plan& s, char* br) const {
    return s.x_dev ? s.x_dev
                   : (s.x_land != NO_LAND ? land_base() + s.x_land : reinterpret_cast<const double*>(br) + s.x_off);
  }

  void jobs_s1(const S1Plan& s, char* br, char* bd, std::vector<AkimaJob>& ak, std::vector<CsplineJob>& cs,
               std::vector<LutJob>& lu) {
    if (!s.present) return;
    const double* x = xptr(s, br);
    const double* y = s.y_dev ? s.y_dev
                              : (s.y_land != NO_LAND ? land_base() + s.y_land
                                                     : reinterpret_cast<const double*>(br) + s.y_off);
    // the LUT is built for every grid: the fast kernel indexes tracer splines through it only
    lu.push_back(LutJob{x, reinterpret_cast<int*>(bd + s.lut_off), s.n, s.nlut});
    double2* xr = reinterpret_cast<double2*>(bd + s.co_off);
    double4* cr = reinterpret_cast<double4*>(bd + s.co_off + rec_x_bytes(s.n));
    if (s.kind == 0) ak.push_back(AkimaJob{x, y, xr, cr, s.n});
    else if (s.kind == 1)
      cs.push_back(CsplineJob{x, y, xr, cr, reinterpret_cast<double*>(bd + s.scr_off), s.n});
    else
      ax_jobs.push_back(AxisJob{x, reinterpret_cast<AxisRec*>(bd + s.co_off), s.n});
  }

  std::vector<AxisJob> ax_jobs;

  void jobs_f2d(const F2DPlan& f, char* br, char* bd, std::vector<AkimaJob>& ak, std::vector<CsplineJob>& cs,
                std::vector<LutJob>& lu, std::vector<BicubicJob>& bc, int& max_nk, int& max_na) {
    jobs_s1(f.fk, br, bd, ak, cs, lu);
    jobs_s1(f.fa, br, bd, ak, cs, lu);
    jobs_s1(f.gk, br, bd, ak, cs, lu);
    jobs_s1(f.ga, br, bd, ak, cs, lu);
    if (f.flags.has_fka) {
      // tables on the same pair of grids (same device arrays) share one set of axis factors
      const double* xk = xptr(f.gk, br);
      const double* ya = xptr(f.ga, br);
      double* scratch = reinterpret_cast<double*>(bd + f.scr_off);
      int factor = 1;
      if (!bc.empty() && bc.back().xk == xk && bc.back().ya == ya && bc.back().nk == f.nk && bc.back().na == f.na) {
        scratch = bc.back().scratch;
        factor = 0;
      }
      bc.push_back(BicubicJob{xk, ya,
                              f.z_dev ? f.z_dev
                                      : (f.z_land != NO_LAND ? land_base() + f.z_land
                                                             : reinterpret_cast<const double*>(br) + f.z_off),
                              reinterpret_cast<double4*>(bd + f.tab_off), scratch, f.nk, f.na, factor});
      max_nk = std::max(max_nk, f.nk);
      max_na = std::max(max_na, f.na);
    }
  }

  bool commit(cudaStream_t stream, int* status) {
    if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice", status)) return false;
    raw.flush();
    run_checks();
    if (!resolve_all(status)) return false;
    raw_bytes = (raw.size * sizeof(double) + 255) / 256 * 256;
    const size_t der_bytes = (derived + 255) / 256 * 256;
    size_t ntr = 0;
    for (auto& c : cosmos) ntr += c.tracers.size();
    const size_t nc = cosmos.size();
    const size_t desc_bytes = nc * sizeof(Cosmo) + nc * sizeof(F2D) + ntr * sizeof(Tracer) + 256;
    desc_off = raw_bytes + der_bytes;
    if (!dev.reserve(desc_off + desc_bytes, status)) return false;
    char* br = dev.p;
    char* bd = dev.p + raw_bytes;
    d_cosmos = reinterpret_cast<Cosmo*>(dev.p + desc_off);
    d_psp = reinterpret_cast<F2D*>(d_cosmos + nc);
    d_tracers = reinterpret_cast<Tracer*>(d_psp + nc);
    if (raw.size &&
        !cuda_ok(cudaMemcpyAsync(br, raw.p, raw.size * sizeof(double), cudaMemcpyHostToDevice, stream),
                 "H2D(raw tables)", status))
      return false;
    last_h2d = raw.size * sizeof(double);
    if (land_size) {
      const size_t need = land_size * sizeof(double);
      if (need > landing.cap && landing.p) {
        // growing after an earlier commit: blocks already landed are kept (their sources may be gone)
        DevBuf grown;
        if (!grown.reserve(need, status)) return false;
        if (!cuda_ok(cudaMemcpyAsync(grown.p, landing.p, landing.cap, cudaMemcpyDeviceToDevice, stream),
                     "D2D(landing)", status) ||
            !cuda_ok(cudaStreamSynchronize(stream), "D2D(landing)", status))
          return false;
        std::swap(grown.p, landing.p);
        std::swap(grown.cap, landing.cap);
      } else if (!landing.reserve(need, status))
        return false;
      for (DirectCopy& c : direct) {
        if (c.done) continue;
        if (!cuda_ok(cudaMemcpyAsync(landing.p + c.off * sizeof(double), c.src, c.n * sizeof(double),
                                     cudaMemcpyHostToDevice, stream), "H2D(pinned stacked tables)", status))
          return false;
        last_h2d += c.n * sizeof(double);
      }
    }
    // descriptors + jobs
    h_cosmos.assign(nc, Cosmo());
    h_psp.assign(nc, F2D());
    h_tracers.clear();
    h_tracers.reserve(ntr);
    tracer_base.assign(nc, 0);
    std::vector<AkimaJob> ak;
    std::vector<CsplineJob> cs;
    std::vector<LutJob> lu;
    std::vector<BicubicJob> bc;
    ax_jobs.clear();
    int max_nk = 0, max_na = 0;
    for (size_t i = 0; i < nc; ++i) {
      const CosmoPlan& c = cosmos[i];
      Cosmo& d = h_cosmos[i];
      memset(&d, 0, sizeof(d));
      d.achi = mk_s1(c.achi, br, bd);
      d.growth = mk_s1(c.growth, br, bd);
      d.has_achi = c.achi.present;
      d.has_growth = c.growth.present;
      jobs_s1(c.achi, br, bd, ak, cs, lu);
      jobs_s1(c.growth, br, bd, ak, cs, lu);
      if (c.has_psp) {
        h_psp[i] = mk_f2d(c.psp, br, bd);
        jobs_f2d(c.psp, br, bd, ak, cs, lu, bc, max_nk, max_na);
      } else
        memset(&h_psp[i], 0, sizeof(F2D));
      d.psp = c.has_psp ? d_psp + i : nullptr;
      tracer_base[i] = h_tracers.size();
      d.tracers = d_tracers + h_tracers.size();
      d.ntracers = (int)c.tracers.size();
      for (const TracerPlan& t : c.tracers) {
        Tracer td;
        memset(&td, 0, sizeof(td));
        td.der_bessel = t.der_bessel; td.der_angles = t.der_angles;
        td.has_kernel = t.kernel.present; td.has_transfer = t.has_transfer;
        td.chi_min = t.chi_min; td.chi_max = t.chi_max;
        td.const_transfer = t.const_transfer; td.const_transfer_value = t.const_transfer_value;
        td.kernel = mk_s1(t.kernel, br, bd);
        jobs_s1(t.kernel, br, bd, ak, cs, lu);
        if (t.has_transfer) {
          td.transfer = mk_f2d(t.transfer, br, bd);
          jobs_f2d(t.transfer, br, bd, ak, cs, lu, bc, max_nk, max_na);
        }
        h_tracers.push_back(td);
      }
    }
    // upload descriptors
    if (!cuda_ok(cudaMemcpyAsync(d_cosmos, h_cosmos.data(), nc * sizeof(Cosmo), cudaMemcpyHostToDevice, stream),
                 "H2D(cosmo desc)", status)) return false;
    if (!cuda_ok(cudaMemcpyAsync(d_psp, h_psp.data(), nc * sizeof(F2D), cudaMemcpyHostToDevice, stream),
                 "H2D(psp desc)", status)) return false;
    if (ntr && !cuda_ok(cudaMemcpyAsync(d_tracers, h_tracers.data(), ntr * sizeof(Tracer),
                                        cudaMemcpyHostToDevice, stream), "H2D(tracer desc)", status))
      return false;
    last_h2d += nc * (sizeof(Cosmo) + sizeof(F2D)) + ntr * sizeof(Tracer);
    // upload jobs
    const size_t jb = ak.size() * sizeof(AkimaJob) + cs.size() * sizeof(CsplineJob) +
                      lu.size() * sizeof(LutJob) + bc.size() * sizeof(BicubicJob) +
                      ax_jobs.size() * sizeof(AxisJob) + 2048;
    if (!jobs.reserve(jb, status)) return false;
    char* q = jobs.p;
    auto up = [&](const void* src, size_t bytes) -> char* {
      char* at = q;
      if (bytes) cudaMemcpyAsync(at, src, bytes, cudaMemcpyHostToDevice, stream);
      q += (bytes + 255) / 256 * 256;
      return at;
    };
    // (job arrays are pageable vectors: the async copies complete before returning to the host
    //  only for pageable memory staged by the driver, which is the documented behaviour)
    AkimaJob* d_ak = (AkimaJob*)up(ak.data(), ak.size() * sizeof(AkimaJob));
    CsplineJob* d_cs = (CsplineJob*)up(cs.data(), cs.size() * sizeof(CsplineJob));
    LutJob* d_lu = (LutJob*)up(lu.data(), lu.size() * sizeof(LutJob));
    BicubicJob* d_bc = (BicubicJob*)up(bc.data(), bc.size() * sizeof(BicubicJob));
    AxisJob* d_ax = (AxisJob*)up(ax_jobs.data(), ax_jobs.size() * sizeof(AxisJob));
    last_h2d += jb - 2048;
    int max_ncs = 0;
    for (const CsplineJob& j : cs) max_ncs = std::max(max_ncs, j.n);
    launch_prep(d_ak, (int)ak.size(), d_cs, (int)cs.size(), d_lu, (int)lu.size(), d_ax, (int)ax_jobs.size(),
                d_bc, (int)bc.size(), max_nk, max_na, max_ncs, stream);
    if (!cuda_ok(cudaGetLastError(), "table preparation launch", status)) return false;
    if (!cuda_ok(cudaStreamSynchronize(stream), "table preparation", status)) return false;
    for (DirectCopy& c : direct) c.done = true;  // the callers' pinned buffers are no longer needed
    owned.clear();
    committed = true;
    ++commit_gen;
    return true;
  }
};

// grow-only device scratch shared per device (qag workspace, call temporaries)
struct DeviceScratch {
  std::mutex mu;
  DevBuf qag;
  DevBuf call;
  HostBuf pinned;
};
DeviceScratch g_scratch[16];

DeviceScratch& scratch_for(int device) { return g_scratch[device & 15]; }

LimberParams to_limber_params(const ccl_b200_params& p) {
  LimberParams lp;
  lp.K_MAX = p.K_MAX; lp.K_MIN = p.K_MIN; lp.DLOGK_INTEGRATION = p.DLOGK_INTEGRATION;
  lp.LIMBER_EPSREL = p.INTEGRATION_LIMBER_EPSREL; lp.N_ITERATION = p.N_ITERATION;
  lp.nk_cap = limber_nk_cap(lp);
  lp.force_cquad = getenv("CCL_B200_FORCE_CQUAD") != nullptr;
  return lp;
}

int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}

}  // namespace

// ------------------------------------------------------------------------------------------
struct ccl_b200_cosmology {
  ccl_b200_params params;
  TableSet ts;
  std::mutex mu;
  bool dirty = true;
};
struct ccl_b200_f2d { TableSet ts; };
struct ccl_b200_f3d {  // ccl_f3d_t: the planes are the P tables of a TableSet with one "cosmology" per plane
  TableSet ts;
  F3D desc;
};
struct ccl_b200_tracer { TableSet ts; double chi_min, chi_max; int der_bessel, der_angles; };
struct ccl_b200_collection { std::vector<ccl_b200_tracer*> ts; };
struct ccl_b200_tables;
struct ccl_b200_batch {
  ccl_b200_params params;
  TableSet ts;
  unsigned long long last_qag_evals = 0, last_qag_nodes = 0;
  bool last_fast = false;  // the last 'spline' launch used the grid-sharing kernels
  cudaStream_t stream = nullptr;  // non-blocking: batches driven from different host threads overlap
  // commit() runs on its own high-priority stream: while another batch's Limber grid (tens of
  // thousands of CTAs) is being dispatched, the small table-preparation kernels of this batch are
  // scheduled ahead of its remaining CTAs instead of behind its last wave (PipelinedBatch)
  cudaStream_t prep_stream = nullptr;
  DevBuf out;   // device result buffer of the host-buffer entry (kept across calls)
  DevBuf call;  // per-batch call buffers (colls, pairs, ells, counters, outputs)
  // The grid tasks of the grid-sharing kernels depend only on the committed tables and on (collections, pairs):
  // they are kept across calls (a likelihood evaluates the same geometry again and again; for one cosmology the
  // grouping pass would otherwise be half of the call) and rebuilt when either changes.
  HostBuf host_status;       // pinned landing buffer of the status words of a started call (_start / _wait)
  size_t pending_status = 0; // entries a started call will deliver
  DevBuf plan;
  std::vector<int> plan_key;
  unsigned long long plan_commit = ~0ull;
  cudaEvent_t plan_event = nullptr;
  // table builders whose halofit pass this batch's next commit waits for (ccl_b200_batch_set_from_tables)
  std::vector<const ccl_b200_tables*> table_deps;
};

struct ccl_b200_tables {
  int device = 0;
  int ncosmo = 0, na = 0, stride = 0, nk = 0, na_pk = 0, nz = 0, ntr = 0, nchi = 0;
  bool has_nl = false;
  // halofit runs on its own stream behind the linear table and is still in flight when ccl_b200_tables_build
  // returns: the caller plans its batch (host work) meanwhile.  Consumers of d_nl wait for ev_nl on their stream
  // (ccl_b200_batch_set_from_tables); the kernel's failure flag is read by the commit that follows.
  cudaStream_t s_nl = nullptr;
  cudaEvent_t ev_nl = nullptr;
  ~ccl_b200_tables() {
    if (s_nl) { cudaStreamSynchronize(s_nl); cudaStreamDestroy(s_nl); }
    if (ev_nl) cudaEventDestroy(ev_nl);
  }
  DevBuf buf;
  double *d_par = nullptr, *d_a = nullptr, *d_E = nullptr, *d_chi = nullptr, *d_gx = nullptr, *d_ga = nullptr,
         *d_D = nullptr, *d_f = nullptr, *d_g0 = nullptr, *d_k = nullptr, *d_lk = nullptr, *d_apk = nullptr,
         *d_lin = nullptr, *d_nl = nullptr, *d_s8 = nullptr, *d_rs = nullptr, *d_z = nullptr, *d_nz = nullptr,
         *d_sz = nullptr, *d_in = nullptr, *d_chiz = nullptr, *d_wnc = nullptr, *d_chil = nullptr,
         *d_wl = nullptr, *d_rng = nullptr, *d_cmax = nullptr;
  int *d_nachi = nullptr, *d_fail = nullptr;
  // host-side metadata the arena planner needs (sizes and end nodes; never the tables themselves)
  std::vector<double> h_a, h_lk, h_apk, h_chi_amin, h_cmax, h_rng_nc, h_rng_l;
  std::vector<int> h_nachi;
  // kernels built after the object (CMB-lensing kappa kernels): kept until the object is freed
  std::deque<DevBuf> extras;
};

static void cosmology_rebuild(ccl_b200_cosmology* c, const double* chi, const double* a, int n,
                              const double* ag, const double* g, int ng, int* status) {
  // re-plan from scratch keeping whichever table is not being replaced
  std::vector<double> kx, ky;
  CosmoPlan old = c->ts.cosmos[0];
  auto grab = [&](const S1Plan& s, std::vector<double>& x, std::vector<double>& y) {
    if (!s.present) return;
    x.assign(c->ts.raw.p + s.x_off, c->ts.raw.p + s.x_off + s.n);
    y.assign(c->ts.raw.p + s.y_off, c->ts.raw.p + s.y_off + s.n);
  };
  std::vector<double> ax, ay, gx, gy;
  if (!chi) grab(old.achi, ax, ay);
  if (!ag) grab(old.growth, gx, gy);
  c->ts.reset();
  CosmoPlan& p = c->ts.cosmos[0];
  p.chi_max_nokernel = old.chi_max_nokernel;
  if (chi) { ax.assign(chi, chi + n); ay.assign(a, a + n); }
  if (ag) { gx.assign(ag, ag + ng); gy.assign(g, g + ng); }
  if (!ax.empty()) {
    if (ax.size() < 5) { *status = CCL_B200_ERROR_MEMORY; set_err("ccl_b200: a(chi) spline needs >= 5 nodes"); return; }
    if (!c->ts.plan_s1(p.achi, 0, (int)ax.size(), ax.data(), ay.data(), status)) return;
  }
  if (!gx.empty()) {
    if (gx.size() < 5) { *status = CCL_B200_ERROR_MEMORY; set_err("ccl_b200: growth spline needs >= 5 nodes"); return; }
    if (!c->ts.plan_s1(p.growth, 0, (int)gx.size(), gx.data(), gy.data(), status)) return;
  }
  c->ts.commit(0, status);
}

// helper: run a vectorised accessor with host in/out arrays
template <class F>
static void run_vec(int device, int nin_arrays, const double* const* in, const size_t* in_n, size_t nout,
                    double* out, int* status, F&& launch) {
  if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice", status)) return;
  size_t tot = 0;
  for (int i = 0; i < nin_arrays; ++i) tot += in_n[i];
  DevBuf buf;
  if (!buf.reserve((tot + nout) * sizeof(double) + 64, status)) return;
  double* d = reinterpret_cast<double*>(buf.p);
  const double* dins[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t off = 0;
  for (int i = 0; i < nin_arrays; ++i) {
    if (in_n[i]) cudaMemcpy(d + off, in[i], in_n[i] * sizeof(double), cudaMemcpyHostToDevice);
    dins[i] = d + off;
    off += in_n[i];
  }
  double* dout = d + off;
  int* dst = reinterpret_cast<int*>(dout + nout);
  cudaMemset(dst, 0, sizeof(int));
  launch(dins, dout, dst);
  if (!cuda_ok(cudaGetLastError(), "accessor launch", status)) return;
  if (!cuda_ok(cudaDeviceSynchronize(), "accessor kernel", status)) return;
  cudaMemcpy(out, dout, nout * sizeof(double), cudaMemcpyDeviceToHost);
  int st = 0;
  cudaMemcpy(&st, dst, sizeof(int), cudaMemcpyDeviceToHost);
  if (st) *status = st;
}

static bool batch_check(ccl_b200_batch* b, int ncoll, const int* coll_start, const int* coll_count,
                        int npair, const int* pair1, const int* pair2, int* status) {
  if (!b->ts.committed) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_b200: batch not committed");
    return false;
  }
  for (size_t i = 0; i < b->ts.cosmos.size(); ++i) {
    const CosmoPlan& c = b->ts.cosmos[i];
    if (!c.achi.present) {
      *status = CCL_B200_ERROR_DISTANCES_INIT;
      set_err("ccl_cls.c: ccl_angular_cls_limber(): distance splines have not been precomputed!");
      return false;
    }
    if (!c.has_psp) { *status = CCL_B200_ERROR_INCONSISTENT; set_err("ccl_b200: cosmology %zu has no P(k)", i); return false; }
    for (int k = 0; k < ncoll; ++k)
      if (coll_start[k] < 0 || coll_count[k] < 0 || coll_start[k] + coll_count[k] > (int)c.tracers.size()) {
        *status = CCL_B200_ERROR_INCONSISTENT;
        set_err("ccl_b200: collection %d exceeds the tracers of cosmology %zu", k, i);
        return false;
      }
  }
  for (int p = 0; p < npair; ++p)
    if (pair1[p] < 0 || pair1[p] >= ncoll || pair2[p] < 0 || pair2[p] >= ncoll) {
      *status = CCL_B200_ERROR_INCONSISTENT;
      set_err("ccl_b200: pair %d names an unknown collection", p);
      return false;
    }
  return true;
}

// The fast 'spline' kernels (limber_fast.cu) cover the table forms pyccl's own tracers produce on
// the default path: log bicubic P(k,a); sub-tracers with an Akima kernel (any der_bessel) and
// either no transfer or an a-only (factorised, linear) one.  Anything else runs on the
// general kernel.
static bool fast_eligible(const ccl_b200_batch* b, int ncoll, int npair) {
  if (getenv("CCL_B200_NO_FAST")) return false;
  if (ncoll > FAST_MAX_COLLS || npair > FAST_MAX_PAIRS) return false;
  for (const CosmoPlan& c : b->ts.cosmos) {
    const F2D& g = c.psp.flags;
    if (!c.has_psp || !g.has_fka || !g.is_log || g.is_factorizable) return false;
    if ((int)c.tracers.size() > FAST_MAX_TRACERS) return false;
    for (const TracerPlan& t : c.tracers) {
      if (!t.kernel.present) return false;
      if (t.has_transfer) {
        const F2D& f = t.transfer.flags;
        if (!f.is_factorizable || f.has_fk || !f.has_fa || f.is_log) return false;
      }
    }
  }
  return true;
}

extern "C" {

void ccl_b200_set_device(int device, int* status) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    set_err("ccl_b200_set_device: device %d not available (%d visible)", device, n);
    if (status) *status = CCL_B200_ERROR_CUDA;
    return;
  }
  g_device = device;
  cuda_ok(cudaSetDevice(device), "cudaSetDevice", status);
}

int ccl_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* ccl_b200_last_error(void) { return g_err.c_str(); }
const char* ccl_b200_version(void) { return "ccl_b200 0.1 (sm_100a)"; }

double ccl_b200_measure_fp64_peak(int* status) {
  if (*status) return -1.0;
  if (!use_device(status)) return -1.0;
  const double t = measure_fp64_peak_tflops(0);
  if (t <= 0) { *status = CCL_B200_ERROR_CUDA; set_err("ccl_b200: fp64 peak measurement failed"); }
  return t;
}

void ccl_b200_default_params(ccl_b200_params* p) {
  p->K_MAX = 1e3; p->K_MIN = 5e-5; p->DLOGK_INTEGRATION = 0.025;
  p->INTEGRATION_LIMBER_EPSREL = 1e-4; p->N_ITERATION = 1000; p->A_SPLINE_MIN = 0.1;
  p->INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS = 4;  // GSL_INTEG_GAUSS41 (src/ccl_core.c:40)
  p->DCHI_INTEGRATION = 5.0;
  p->ELL_MIN_CORR = 0.01; p->ELL_MAX_CORR = 60000; p->N_ELL_CORR = 5000;  // src/ccl_core.c:129-131
  p->INTEGRATION_GAUSS_KRONROD_POINTS = 4; p->INTEGRATION_EPSREL = 1e-4;  // src/ccl_core.c:69-70
}

// gsl_integration_qag takes the rule key from the parameters (src/ccl_cls.c:240); the device has GK41 only
static bool qag_rule_ok(const ccl_b200_params& p, int* status) {
  if (p.INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS == 4) return true;
  *status = CCL_B200_ERROR_NOT_IMPLEMENTED;
  set_err("ccl_b200: INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS = %d: only GSL_INTEG_GAUSS41 (4) is implemented "
          "for limber_integration_method='qag_quad'", p.INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS);
  return false;
}

// ---- cosmology ----
ccl_b200_cosmology* ccl_b200_cosmology_new(const ccl_b200_params* params, int* status) {
  if (*status) return nullptr;
  if (!use_device(status)) return nullptr;
  auto* c = new ccl_b200_cosmology();
  if (params) c->params = *params; else ccl_b200_default_params(&c->params);
  c->ts.device = current_device();
  c->ts.cosmos.resize(1);
  return c;
}

void ccl_b200_cosmology_free(ccl_b200_cosmology* cosmo) {
  if (!cosmo) return;
  cudaSetDevice(cosmo->ts.device);
  delete cosmo;
}

void ccl_b200_cosmology_set_achi(ccl_b200_cosmology* cosmo, int n, const double* chi, const double* a,
                                 int* status) {
  if (*status || !cosmo) return;
  std::lock_guard<std::mutex> lk(cosmo->mu);
  cosmology_rebuild(cosmo, chi, a, n, nullptr, nullptr, 0, status);
}

void ccl_b200_cosmology_distances_from_input(ccl_b200_cosmology* cosmo, int na, const double* a,
                                             const double* chi_a, int* status) {
  if (*status || !cosmo) return;
  std::vector<double> cr(na), ar(na);
  for (int i = 0; i < na; ++i) { cr[i] = chi_a[na - 1 - i]; ar[i] = a[na - 1 - i]; }
  cosmo->params.A_SPLINE_MIN = a[0];  // src/ccl_background.c:617
  std::lock_guard<std::mutex> lk(cosmo->mu);
  cosmology_rebuild(cosmo, cr.data(), ar.data(), na, nullptr, nullptr, 0, status);
}

void ccl_b200_cosmology_set_growth(ccl_b200_cosmology* cosmo, int n, const double* a, const double* growth,
                                   int* status) {
  if (*status || !cosmo) return;
  std::lock_guard<std::mutex> lk(cosmo->mu);
  cosmology_rebuild(cosmo, nullptr, nullptr, 0, a, growth, n, status);
}

void ccl_b200_cosmology_set_chi_max_nokernel(ccl_b200_cosmology* cosmo, double chi) {
  if (cosmo) cosmo->ts.cosmos[0].chi_max_nokernel = chi;
}

void ccl_b200_scale_factor_of_chis(ccl_b200_cosmology* cosmo, int n, const double* chi, double* a_out,
                                   int* status) {
  if (*status || !cosmo) return;
  if (!cosmo->ts.committed || !cosmo->ts.cosmos[0].achi.present) {
    *status = CCL_B200_ERROR_DISTANCES_INIT;
    set_err("ccl_background.c: ccl_scale_factor_of_chi(): distance splines have not been precomputed!");
    return;
  }
  const double* in[1] = {chi};
  const size_t nn[1] = {(size_t)n};
  run_vec(cosmo->ts.device, 1, in, nn, n, a_out, status, [&](const double* const* di, double* dout, int* dst) {
    launch_eval_achi(cosmo->ts.d_cosmos, n, di[0], dout, dst, 0);
  });
}

// ---- f2d ----
ccl_b200_f2d* ccl_b200_f2d_new(int na, const double* a_arr, int nk, const double* lk_arr,
                               const double* fka_arr, const double* fk_arr, const double* fa_arr,
                               int is_factorizable, int extrap_order_lok, int extrap_order_hik,
                               int extrap_linear_growth, int is_fka_log, double growth_factor_0,
                               int growth_exponent, int* status) {
  if (*status) return nullptr;
  if (!use_device(status)) return nullptr;
  auto* f = new ccl_b200_f2d();
  f->ts.device = current_device();
  f->ts.cosmos.resize(1);
  CosmoPlan& p = f->ts.cosmos[0];
  p.has_psp = true;
  if (!f->ts.plan_f2d(p.psp, na, a_arr, nk, lk_arr, fka_arr, fk_arr, fa_arr, is_factorizable,
                      extrap_order_lok, extrap_order_hik, extrap_linear_growth, is_fka_log,
                      growth_factor_0, growth_exponent, status) ||
      !f->ts.commit(0, status)) {
    delete f;
    return nullptr;
  }
  return f;
}

ccl_b200_f2d* ccl_b200_pk2d_new_from_arrays(const double* lk_arr, int nk, const double* a_arr, int na,
                                            const double* pk_arr, int order_lok, int order_hik,
                                            int is_logp, int* status) {
  return ccl_b200_f2d_new(na, a_arr, nk, lk_arr, pk_arr, nullptr, nullptr, 0, order_lok, order_hik,
                          GROWTH_CCL, is_logp, 0, 2, status);
}

void ccl_b200_f2d_free(ccl_b200_f2d* f2d) {
  if (!f2d) return;
  cudaSetDevice(f2d->ts.device);
  delete f2d;
}

static void f2d_eval_multi(ccl_b200_f2d* f2d, const double* lk, int nk, double a, ccl_b200_cosmology* cosmo,
                          double* out, int* status, int deriv);
void ccl_b200_f2d_eval_multi(ccl_b200_f2d* f2d, const double* lk, int nk, double a, ccl_b200_cosmology* cosmo,
                             double* out, int* status) {
  f2d_eval_multi(f2d, lk, nk, a, cosmo, out, status, 0);
}
// pk2d_der_eval_multi (pyccl/ccl_pk2d.i:58-64) -> ccl_f2d_t_dlogf_dlk_eval
void ccl_b200_f2d_dlogf_dlk_eval_multi(ccl_b200_f2d* f2d, const double* lk, int nk, double a, ccl_b200_cosmology* cosmo,
                                       double* out, int* status) {
  f2d_eval_multi(f2d, lk, nk, a, cosmo, out, status, 1);
}
static void f2d_eval_multi(ccl_b200_f2d* f2d, const double* lk, int nk, double a, ccl_b200_cosmology* cosmo,
                          double* out, int* status, int deriv) {
  if (*status || !f2d) return;
  std::vector<double> av(nk, a);
  const double* in[2] = {lk, av.data()};
  const size_t nn[2] = {(size_t)nk, (size_t)nk};
  const Cosmo* dcs = (cosmo && cosmo->ts.committed) ? cosmo->ts.d_cosmos : nullptr;
  run_vec(f2d->ts.device, 2, in, nn, nk, out, status, [&](const double* const* di, double* dout, int* dst) {
    launch_eval_f2d(f2d->ts.d_psp, dcs, nk, di[0], di[1], dout, dst, 0, deriv);
  });
  if (*status == CCL_B200_ERROR_GROWTH_INIT)
    set_err("ccl_f2d.c: ccl_f2d_t_eval(): growth factor splines have not been precomputed!");
}

// ---- ccl_f3d_t and the Limber covariance (src/ccl_f3d.c, src/ccl_cls.c:377-612) ----
ccl_b200_f3d* ccl_b200_f3d_new(int na, const double* a_arr, int nk, const double* lk_arr, const double* tkka_arr,
                               const double* fka1_arr, const double* fka2_arr, int is_product,
                               int extrap_order_lok, int extrap_order_hik, int extrap_linear_growth,
                               int is_tkka_log, double growth_factor_0, int growth_exponent, int* status) {
  if (*status) return nullptr;
  if (!use_device(status)) return nullptr;
  is_product = is_product || (tkka_arr == nullptr);
  if ((extrap_order_lok > 1) || (extrap_order_lok < 0) || (extrap_order_hik > 1) || (extrap_order_hik < 0) ||
      ((extrap_linear_growth != GROWTH_CCL) && (extrap_linear_growth != GROWTH_CONST) &&
       (extrap_linear_growth != GROWTH_NONE)) ||
      (is_product ? (fka1_arr == nullptr || fka2_arr == nullptr) : (tkka_arr == nullptr)) || na < 1 || nk < 4 ||
      !a_arr || !lk_arr) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_b200: ccl_f3d_t_new(): inconsistent arguments");
    return nullptr;
  }
  auto* f = new ccl_b200_f3d();
  f->ts.device = current_device();
  const int nplanes = is_product ? 2 : na;
  f->ts.cosmos.resize(nplanes);
  const size_t a_off = f->ts.raw.put(a_arr, na);
  bool ok = a_off != (size_t)-1;
  for (int i = 0; i < nplanes && ok; ++i) {
    CosmoPlan& c = f->ts.cosmos[i];
    c.has_psp = true;
    if (is_product)  // two ccl_f2d_t with half the growth exponent (src/ccl_f3d.c:226-241)
      ok = f->ts.plan_f2d(c.psp, na, a_arr, nk, lk_arr, i == 0 ? fka1_arr : fka2_arr, nullptr, nullptr, 0,
                          extrap_order_lok, extrap_order_hik, extrap_linear_growth, is_tkka_log, growth_factor_0,
                          growth_exponent / 2, status);
    else  // gsl_spline2d_init(tkka[ia], lk_arr, lk_arr, tkk, nk, nk): a bicubic table on (ln k1, ln k2)
      ok = f->ts.plan_f2d(c.psp, nk, lk_arr, nk, lk_arr, tkka_arr + (size_t)i * nk * nk, nullptr, nullptr, 0, 0, 0,
                          GROWTH_CONST, 0, 1.0, 0, status);
  }
  if (!ok || !f->ts.commit(0, status)) {
    if (!*status) *status = CCL_B200_ERROR_MEMORY;
    delete f;
    return nullptr;
  }
  F3D& d = f->desc;
  memset(&d, 0, sizeof(d));
  d.na = na; d.is_product = is_product;
  d.extrap_order_lok = extrap_order_lok; d.extrap_order_hik = extrap_order_hik;
  d.extrap_linear_growth = extrap_linear_growth; d.is_log = is_tkka_log; d.growth_exponent = growth_exponent;
  d.growth_factor_0 = growth_factor_0; d.lkmin = lk_arr[0]; d.lkmax = lk_arr[nk - 1];
  d.a_arr = reinterpret_cast<const double*>(f->ts.dev.p) + a_off;
  d.planes = f->ts.d_psp;
  return f;
}

void ccl_b200_f3d_free(ccl_b200_f3d* f3d) {
  if (!f3d) return;
  cudaSetDevice(f3d->ts.device);
  delete f3d;
}

void ccl_b200_angular_cl_covariance(ccl_b200_cosmology* cosmo, ccl_b200_collection* trc1, ccl_b200_collection* trc2,
                                    ccl_b200_collection* trc3, ccl_b200_collection* trc4, ccl_b200_f3d* tsp,
                                    int nl1_out, const double* l1_out, int nl2_out, const double* l2_out,
                                    double* cov_out, int integration_method, int chi_exponent, int n_ker,
                                    const double* a_ker, const double* f_ker, double prefactor_extra, int* status) {
  if (!cosmo || !trc1 || !trc2 || !trc3 || !trc4 || !tsp) { *status = CCL_B200_ERROR_INCONSISTENT; return; }
  std::lock_guard<std::mutex> lk(cosmo->mu);
  if (!cosmo->ts.committed || !cosmo->ts.cosmos[0].achi.present) {
    *status = CCL_B200_ERROR_DISTANCES_INIT;
    set_err("ccl_cls.c: ccl_angular_cl_limber(): distance splines have not been precomputed!");
    return;
  }
  if (*status) return;
  const size_t nout = (size_t)nl1_out * nl2_out;
  auto fail_all = [&](int code, const char* msg) {
    for (size_t i = 0; i < nout; ++i) cov_out[i] = NAN;
    *status = code;
    set_err("%s", msg);
  };
  if (integration_method != INTEG_QAG && integration_method != INTEG_SPLINE)
    return fail_all(CCL_B200_ERROR_INTEG, "ccl_cls.c: ccl_angular_cov_limber(); integration error\n");
  if (integration_method == INTEG_QAG && !qag_rule_ok(cosmo->params, status)) {
    for (size_t i = 0; i < nout; ++i) cov_out[i] = NAN;
    return;
  }
  const int device = cosmo->ts.device;
  if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice", status)) return;
  // kernel_extra: a ccl_f1d_t of (a, f) with constant extrapolation (pyccl/ccl_covs.i:95-118), built like any Akima table
  TableSet ker;
  ker.device = device;
  if (n_ker > 0) {
    if (n_ker < 5 || !a_ker || !f_ker) {
      *status = CCL_B200_ERROR_SPLINE;
      set_err("ccl_b200: kernel_extra needs >= 5 nodes (akima)");
      return;
    }
    ker.cosmos.resize(1);
    if (!ker.plan_s1(ker.cosmos[0].achi, 0, n_ker, a_ker, f_ker, status) || !ker.commit(0, status)) return;
  }
  DeviceScratch& sc = scratch_for(device);
  std::lock_guard<std::mutex> lk2(sc.mu);
  // update_chi_limits (src/ccl_cls.c:34-62): union over trc1, then intersections
  ccl_b200_collection* cols[4] = {trc1, trc2, trc3, trc4};
  double chimin = 1E15, chimax = -1E15;
  std::vector<Tracer> trs;
  int2 coll[4];
  for (int c = 0; c < 4; ++c) {
    double lo = 1E15, hi = -1E15;
    coll[c] = make_int2((int)trs.size(), (int)cols[c]->ts.size());
    for (auto* t : cols[c]->ts) {
      lo = std::min(lo, t->chi_min);
      hi = std::max(hi, t->chi_max);
      trs.push_back(t->ts.h_tracers[0]);
    }
    if (c == 0) { chimin = std::min(chimin, lo); chimax = std::max(chimax, hi); }
    else { chimin = std::max(chimin, lo); chimax = std::min(chimax, hi); }
  }
  const double dchi = cosmo->params.DCHI_INTEGRATION;
  const int nchi = (int)(fmax((chimax - chimin) / dchi + 0.5, 1)) + 1;
  const int nit = cosmo->params.N_ITERATION;
  // call buffer: Tracer[] | Cosmo | F3D | l1 | l2 | cov | status[2] | scratch
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t o_tr = 0;
  const size_t o_cs = o_tr + up(trs.size() * sizeof(Tracer));
  const size_t o_f3 = o_cs + up(sizeof(Cosmo));
  const size_t o_l1 = o_f3 + up(sizeof(F3D));
  const size_t o_l2 = o_l1 + up((size_t)nl1_out * 8);
  const size_t o_cv = o_l2 + up((size_t)nl2_out * 8);
  const size_t o_st = o_cv + up(nout * 8);
  const size_t o_sc = o_st + 256;
  const size_t total = o_sc + covariance_scratch_bytes(nchi, nit);
  if (!sc.call.reserve(total, status)) return;
  char* base = sc.call.p;
  Cosmo cd = cosmo->ts.h_cosmos[0];
  cd.psp = nullptr;
  cd.tracers = reinterpret_cast<const Tracer*>(base + o_tr);
  cd.ntracers = (int)trs.size();
  cudaMemcpyAsync(base + o_tr, trs.data(), trs.size() * sizeof(Tracer), cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(base + o_cs, &cd, sizeof(Cosmo), cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(base + o_f3, &tsp->desc, sizeof(F3D), cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(base + o_l1, l1_out, (size_t)nl1_out * 8, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(base + o_l2, l2_out, (size_t)nl2_out * 8, cudaMemcpyHostToDevice, 0);
  cudaMemsetAsync(base + o_st, 0, 256, 0);
  CovProblem P;
  memset(&P, 0, sizeof(P));
  P.cosmo = reinterpret_cast<const Cosmo*>(base + o_cs);
  P.tracers = reinterpret_cast<const Tracer*>(base + o_tr);
  for (int c = 0; c < 4; ++c) P.coll[c] = coll[c];
  P.tsp = reinterpret_cast<const F3D*>(base + o_f3);
  P.l1 = reinterpret_cast<const double*>(base + o_l1);
  P.l2 = reinterpret_cast<const double*>(base + o_l2);
  P.nl1 = nl1_out; P.nl2 = nl2_out; P.chipow = chi_exponent;
  P.has_ker = n_ker > 0;
  if (n_ker > 0) { P.ker = ker.h_cosmos[0].achi; P.ker_y0 = f_ker[0]; P.ker_yf = f_ker[n_ker - 1]; }
  P.chimin = chimin; P.chimax = chimax; P.prefactor = prefactor_extra;
  P.cov = reinterpret_cast<double*>(base + o_cv);
  P.status = reinterpret_cast<int*>(base + o_st);
  if (!cuda_ok(launch_covariance(P, integration_method, nchi, nit, cosmo->params.INTEGRATION_LIMBER_EPSREL,
                                 reinterpret_cast<double*>(base + o_sc), 0), "covariance launch", status))
    return;
  int st[2] = {0, 0};
  if (!cuda_ok(cudaMemcpy(cov_out, P.cov, nout * 8, cudaMemcpyDeviceToHost), "covariance kernel", status)) return;
  cudaMemcpy(st, P.status, sizeof(st), cudaMemcpyDeviceToHost);
  if (st[0]) {
    *status = CCL_B200_ERROR_INTEG;
    set_err("ccl_cls.c: ccl_angular_cov_limber(); integration error\n");
  }
}

// fftlog_transform_general (pyccl/ccl_fftlog.i:63-90), batched over mu
void ccl_b200_fftlog_transform_general(int nmu, const double* mu, int npk, const double* k_in, int n_in_k,
                                       const double* fk_in, double q, int spherical_bessel, double bessel_deriv,
                                       double plaw, double* output, int* status) {
  if (*status) return;
  if (!mu || !k_in || !fk_in || !output || nmu < 1 || npk < 1 || n_in_k < 2) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    return;
  }
  if (round(bessel_deriv) != bessel_deriv || bessel_deriv < 0) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("bessel_deriv must be a non-negative integer in a floating-point format.");
    return;
  }
  if (spherical_bessel != 0 && spherical_bessel != 1) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("0 and 1 are the only valid values of spherical_bessel.");
    return;
  }
  if (!use_device(status)) return;
  int device = 0;
  cudaGetDevice(&device);
  DeviceScratch& sc = scratch_for(device);
  std::lock_guard<std::mutex> lk2(sc.mu);
  if (!sc.call.reserve(fftlog_general_work_bytes(nmu, npk, n_in_k), status)) return;
  double* out_dev = nullptr;
  if (!cuda_ok(launch_fftlog_general(nmu, mu, npk, n_in_k, k_in, fk_in, q, spherical_bessel, bessel_deriv, plaw,
                                     sc.call.p, &out_dev, 0),
               "launch_fftlog_general", status))
    return;
  cudaMemcpyAsync(output, out_dev, (size_t)nmu * (npk + 1) * n_in_k * 8, cudaMemcpyDeviceToHost, 0);
  cuda_ok(cudaStreamSynchronize(0), "ccl_b200_fftlog_transform_general", status);
}

// ccl_correlation (src/ccl_correlation.c:410-428)
void ccl_b200_correlation(ccl_b200_cosmology* cosmo, int n_ell, const double* ell, const double* cls, int n_theta,
                          const double* theta, double* wtheta, int corr_type, int do_taper_cl,
                          const double* taper_cl_limits, int flag_method, int* status) {
  if (!cosmo || !ell || !cls || !theta || !wtheta || (do_taper_cl && !taper_cl_limits)) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    return;
  }
  if (*status) return;
  std::lock_guard<std::mutex> lk(cosmo->mu);
  const ccl_b200_params& par = cosmo->params;
  if (flag_method != CORR_FFTLOG && flag_method != CORR_LGNDRE && flag_method != CORR_BESSEL) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_correlation.c: ccl_correlation(): Unknown algorithm\n");
    return;
  }
  if (flag_method == CORR_LGNDRE && (corr_type == CORR_LM || corr_type == CORR_LP)) {
    *status = CCL_B200_ERROR_NOT_IMPLEMENTED;
    set_err("ccl_correlation.c: ccl_tracer_corr_legendre(): CCL does not support full-sky xi+- calcuations.\n");
    return;
  }
  if (flag_method == CORR_BESSEL && par.INTEGRATION_GAUSS_KRONROD_POINTS != 4) {
    *status = CCL_B200_ERROR_NOT_IMPLEMENTED;
    set_err("ccl_b200: only the 41-point Gauss-Kronrod rule (GSL_INTEG_GAUSS41 = 4) is built on the device");
    return;
  }
  // ccl_f1d_t_new(n_ell, ell, cls, cls[0], 0, extrap_const, extrap_logx_logy): gsl_spline_alloc(akima) needs 5
  // nodes; the power-law slope needs a same-sign last pair (src/ccl_f1d.c:97-102).  A missing spline is
  // CCL_ERROR_SPLINE for FFTLog (its status is tested first, :73-81) and CCL_ERROR_MEMORY otherwise (:190-197)
  const bool bad_nodes = n_ell < 5;
  const bool bad_slope = !bad_nodes && ((cls[n_ell - 1] * cls[n_ell - 2] <= 0) || (ell[n_ell - 1] * ell[n_ell - 2] <= 0));
  if (bad_nodes || bad_slope) {
    *status = (flag_method == CORR_FFTLOG && !bad_nodes) ? CCL_B200_ERROR_SPLINE : CCL_B200_ERROR_MEMORY;
    set_err("ccl_correlation.c: ccl_tracer_corr_*(): failed to create spline\n");
    return;
  }
  const int N = par.N_ELL_CORR;
  std::vector<double> l_arr;
  double L = 0.0;
  if (flag_method == CORR_FFTLOG) {
    // ccl_log_spacing (src/ccl_utils.c:109-148): running product with the end points overwritten
    if (N < 5 || !(par.ELL_MIN_CORR > 0 && par.ELL_MAX_CORR > 0)) {
      *status = CCL_B200_ERROR_LINSPACE;
      set_err("ccl_correlation.c: ccl_tracer_corr_fftlog(): ran out of memory\n");
      return;
    }
    l_arr.resize(N);
    const double xratio = exp((log(par.ELL_MAX_CORR) - log(par.ELL_MIN_CORR)) / (N - 1.));
    l_arr[0] = par.ELL_MIN_CORR;
    for (int i = 1; i < N - 1; ++i) l_arr[i] = l_arr[i - 1] * xratio;
    l_arr[N - 1] = par.ELL_MAX_CORR;
    L = log(l_arr[N - 1] / l_arr[0]) * N / (N - 1.);
  }
  if (n_theta == 0) return;
  const int device = cosmo->ts.device;
  if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice", status)) return;
  DeviceScratch& sc = scratch_for(device);
  std::lock_guard<std::mutex> lk2(sc.mu);
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t o_ell = 0;
  const size_t o_cls = o_ell + up((size_t)n_ell * 8);
  const size_t o_th = o_cls + up((size_t)n_ell * 8);
  const size_t o_la = o_th + up((size_t)n_theta * 8);
  const size_t o_out = o_la + up(l_arr.size() * 8);
  const size_t o_st = o_out + up((size_t)n_theta * 8);
  const size_t o_wk = o_st + 256;
  const size_t total = o_wk + correlation_work_bytes(flag_method, n_ell, n_theta, N, par.ELL_MAX_CORR, par.N_ITERATION);
  if (!sc.call.reserve(total, status)) return;
  char* base = sc.call.p;
  cudaMemcpyAsync(base + o_ell, ell, (size_t)n_ell * 8, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(base + o_cls, cls, (size_t)n_ell * 8, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(base + o_th, theta, (size_t)n_theta * 8, cudaMemcpyHostToDevice, 0);
  if (!l_arr.empty()) cudaMemcpyAsync(base + o_la, l_arr.data(), l_arr.size() * 8, cudaMemcpyHostToDevice, 0);
  cudaMemsetAsync(base + o_st, 0, 256, 0);
  CorrProblem P;
  memset(&P, 0, sizeof(P));
  P.ell = reinterpret_cast<const double*>(base + o_ell);
  P.cls = reinterpret_cast<const double*>(base + o_cls);
  P.theta = reinterpret_cast<const double*>(base + o_th);
  P.l_arr = reinterpret_cast<const double*>(base + o_la);
  P.n_ell = n_ell; P.n_theta = n_theta; P.n_ell_corr = N; P.corr_type = corr_type; P.method = flag_method;
  P.do_taper = do_taper_cl ? 1 : 0; P.n_iteration = par.N_ITERATION;
  if (do_taper_cl) for (int i = 0; i < 4; ++i) P.taper[i] = taper_cl_limits[i];
  P.ell_max_corr = par.ELL_MAX_CORR; P.epsrel = par.INTEGRATION_EPSREL; P.L = L;
  P.der_hi = log(cls[n_ell - 1] / cls[n_ell - 2]) / log(ell[n_ell - 1] / ell[n_ell - 2]);
  P.x_ini = ell[0]; P.x_end = ell[n_ell - 1]; P.y0 = cls[0]; P.y_end = cls[n_ell - 1];
  P.k0 = l_arr.empty() ? 0.0 : l_arr[0];
  P.wtheta = reinterpret_cast<double*>(base + o_out);
  P.status = reinterpret_cast<int*>(base + o_st);
  P.work = base + o_wk;
  if (!cuda_ok(launch_correlation(P, 0), "launch_correlation", status)) return;
  int dst = 0;
  cudaMemcpyAsync(wtheta, base + o_out, (size_t)n_theta * 8, cudaMemcpyDeviceToHost, 0);
  cudaMemcpyAsync(&dst, base + o_st, sizeof(int), cudaMemcpyDeviceToHost, 0);
  if (!cuda_ok(cudaStreamSynchronize(0), "ccl_b200_correlation", status)) return;
  if (dst) {  // the reference keeps the results and ORs the GSL codes into the status (src/ccl_correlation.c:243-251)
    *status = dst;
    set_err("ccl_correlation.c: ccl_tracer_corr_bessel(): GSL integration status %d\n", dst);
  }
}

// ---- tracers ----
ccl_b200_tracer* ccl_b200_cl_tracer_t_new(ccl_b200_cosmology* cosmo, int der_bessel, int der_angles,
                                          int n_w, const double* chi_w, const double* w_w, int na_ka,
                                          const double* a_ka, int nk_ka, const double* lk_ka,
                                          const double* fka_arr, const double* fk_arr,
                                          const double* fa_arr, int is_fka_log, int is_factorizable,
                                          int extrap_order_lok, int extrap_order_hik, int* status) {
  if (*status) return nullptr;
  if (!use_device(status)) return nullptr;
  auto* t = new ccl_b200_tracer();
  t->ts.device = current_device();
  t->ts.cosmos.resize(1);
  CosmoPlan& p = t->ts.cosmos[0];
  p.tracers.resize(1);
  const double cmax = cosmo ? cosmo->ts.cosmos[0].chi_max_nokernel : 1e15;
  if (!t->ts.plan_tracer(p.tracers[0], cmax, der_bessel, der_angles, n_w, chi_w, w_w, na_ka, a_ka, nk_ka,
                         lk_ka, fka_arr, fk_arr, fa_arr, is_fka_log, is_factorizable, extrap_order_lok,
                         extrap_order_hik, status) ||
      !t->ts.commit(0, status)) {
    delete t;
    return nullptr;
  }
  t->chi_min = p.tracers[0].chi_min;
  t->chi_max = p.tracers[0].chi_max;
  t->der_bessel = der_bessel;
  t->der_angles = der_angles;
  return t;
}

void ccl_b200_cl_tracer_t_free(ccl_b200_tracer* tr) {
  if (!tr) return;
  cudaSetDevice(tr->ts.device);
  delete tr;
}

double ccl_b200_cl_tracer_t_chi_min(const ccl_b200_tracer* tr) { return tr->chi_min; }
double ccl_b200_cl_tracer_t_chi_max(const ccl_b200_tracer* tr) { return tr->chi_max; }

void ccl_b200_cl_tracer_get_f_ell(const ccl_b200_tracer* tr, int n, const double* ell, double* out,
                                  int* status) {
  if (*status) return;
  // closed form, src/ccl_tracers.c:703-726 (host arithmetic of a per-ell constant, not table work)
  for (int i = 0; i < n; ++i) {
    const double l = ell[i];
    double f = 1;
    if (tr->der_angles == 1) f = l * (l + 1.);
    else if (tr->der_angles == 2) {
      if (l <= 1) f = 0;
      else if (l <= 10) f = sqrt((l + 2) * (l + 1) * l * (l - 1));
      else {
        const double lp1h = l + 0.5, lp1h2 = lp1h * lp1h;
        f = (l <= 1000) ? lp1h2 * (1 - 1.25 / lp1h2) : lp1h2;
      }
    }
    out[i] = f;
  }
}

void ccl_b200_cl_tracer_get_kernel(ccl_b200_tracer* tr, int n, const double* chi, double* out, int* status) {
  if (*status || !tr) return;
  const double* in[1] = {chi};
  const size_t nn[1] = {(size_t)n};
  run_vec(tr->ts.device, 1, in, nn, n, out, status, [&](const double* const* di, double* dout, int*) {
    launch_eval_tracer_kernel(tr->ts.d_tracers, n, di[0], dout, 0);
  });
}

void ccl_b200_cl_tracer_get_transfer(ccl_b200_tracer* tr, int nk, const double* lk, int na, const double* a,
                                     double* out, int* status) {
  if (*status || !tr) return;
  const size_t n = (size_t)nk * na;
  std::vector<double> lkv(n), av(n);
  for (int i = 0; i < nk; ++i)
    for (int j = 0; j < na; ++j) { lkv[(size_t)i * na + j] = lk[i]; av[(size_t)i * na + j] = a[j]; }
  const double* in[2] = {lkv.data(), av.data()};
  const size_t nn[2] = {n, n};
  run_vec(tr->ts.device, 2, in, nn, n, out, status, [&](const double* const* di, double* dout, int* dst) {
    launch_eval_tracer_transfer(tr->ts.d_tracers, (int)n, di[0], di[1], dout, dst, 0);
  });
}

// ---- collections ----
ccl_b200_collection* ccl_b200_cl_tracer_collection_t_new(int* status) {
  if (*status) return nullptr;
  return new ccl_b200_collection();
}
void ccl_b200_cl_tracer_collection_t_free(ccl_b200_collection* trc) { delete trc; }
void ccl_b200_add_cl_tracer_to_collection(ccl_b200_collection* trc, ccl_b200_tracer* tr, int* status) {
  if (*status) return;
  if (trc->ts.size() >= 100) {  // CCL_MAX_TRACERS_PER_COLLECTION, include/ccl_tracers.h:167
    *status = CCL_B200_ERROR_MEMORY;
    set_err("ccl_tracers.c: ccl_add_cl_tracer_to_collection(): already reached maximum number of tracers\n");
    return;
  }
  trc->ts.push_back(tr);
}

// ---- the hot path, single pair ----
void ccl_b200_angular_cls_limber(ccl_b200_cosmology* cosmo, ccl_b200_collection* trc1,
                                 ccl_b200_collection* trc2, ccl_b200_f2d* psp, int nl_out,
                                 const double* l_out, double* cl_out, int integration_method,
                                 int* status) {
  if (!cosmo || !trc1 || !trc2 || !psp) { *status = CCL_B200_ERROR_INCONSISTENT; return; }
  std::lock_guard<std::mutex> lk(cosmo->mu);
  if (!cosmo->ts.committed || !cosmo->ts.cosmos[0].achi.present) {
    *status = CCL_B200_ERROR_DISTANCES_INIT;
    set_err("ccl_cls.c: ccl_angular_cls_limber(): distance splines have not been precomputed!");
    return;
  }
  // The reference proceeds into the OpenMP region with local_status = *status and skips all work
  // if it is non-zero (src/ccl_cls.c:289-319).
  if (*status) return;
  if (integration_method != INTEG_QAG && integration_method != INTEG_SPLINE) {
    for (int i = 0; i < nl_out; ++i) cl_out[i] = NAN;
    *status = CCL_B200_ERROR_INTEG;  // CCL_ERROR_NOT_IMPLEMENTED is overwritten at :343-345
    set_err("ccl_cls.c: ccl_angular_cls_limber(); integration error\n");
    return;
  }
  if (integration_method == INTEG_QAG && !qag_rule_ok(cosmo->params, status)) {
    for (int i = 0; i < nl_out; ++i) cl_out[i] = NAN;
    return;
  }
  const int device = cosmo->ts.device;
  if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice", status)) return;
  DeviceScratch& sc = scratch_for(device);
  std::lock_guard<std::mutex> lk2(sc.mu);
  const int n1 = (int)trc1->ts.size(), n2 = (int)trc2->ts.size();
  const bool same = (trc1 == trc2);
  // layout of the call buffer: Tracer[n1+n2] | Cosmo | int2 colls[2] | int2 pair | ells | cl | status[2]
  std::vector<Tracer> trs;
  for (auto* t : trc1->ts) trs.push_back(t->ts.h_tracers[0]);
  for (auto* t : trc2->ts) trs.push_back(t->ts.h_tracers[0]);
  Cosmo cd = cosmo->ts.h_cosmos[0];
  const size_t o_tr = 0;
  const size_t o_cs = (o_tr + trs.size() * sizeof(Tracer) + 255) / 256 * 256;
  const size_t o_co = o_cs + 256 * ((sizeof(Cosmo) + 255) / 256);
  const size_t o_pr = o_co + 256;
  const size_t o_el = o_pr + 256;
  const size_t o_cl = o_el + ((size_t)nl_out * 8 + 255) / 256 * 256;
  const size_t o_st = o_cl + ((size_t)nl_out * 8 + 255) / 256 * 256;
  const size_t total = o_st + 256;
  if (!sc.call.reserve(total, status)) return;
  char* base = sc.call.p;
  cd.psp = psp->ts.d_psp;
  cd.tracers = reinterpret_cast<const Tracer*>(base + o_tr);
  cd.ntracers = n1 + n2;
  const int2 colls[2] = {make_int2(0, n1), make_int2(n1, n2)};
  const int2 pair = same ? make_int2(0, 0) : make_int2(0, 1);
  cudaMemcpyAsync(base + o_tr, trs.data(), trs.size() * sizeof(Tracer), cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(base + o_cs, &cd, sizeof(Cosmo), cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(base + o_co, colls, sizeof(colls), cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(base + o_pr, &pair, sizeof(pair), cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(base + o_el, l_out, (size_t)nl_out * 8, cudaMemcpyHostToDevice, 0);
  cudaMemsetAsync(base + o_st, 0, 256, 0);
  LimberProblem P;
  P.cosmos = reinterpret_cast<const Cosmo*>(base + o_cs);
  P.colls = reinterpret_cast<const int2*>(base + o_co);
  P.pairs = reinterpret_cast<const int2*>(base + o_pr);
  P.ells = reinterpret_cast<const double*>(base + o_el);
  P.ncosmo = 1; P.npair = 1; P.nl = nl_out; P.ncoll = 2;
  P.par = to_limber_params(cosmo->params);
  P.cl = reinterpret_cast<double*>(base + o_cl);
  P.status = reinterpret_cast<int*>(base + o_st);
  P.status_detail = P.status + 1;
  P.counters = nullptr;
  double* ws = nullptr;
  size_t wsb = 0;
  if (integration_method == INTEG_QAG) {
    wsb = limber_qag_workspace_bytes(P.par.N_ITERATION);
    if (!sc.qag.reserve(wsb, status)) return;
    ws = reinterpret_cast<double*>(sc.qag.p);
  }
  if (!cuda_ok(launch_limber(P, integration_method, ws, wsb, 0), "limber launch", status)) return;
  if (!cuda_ok(cudaMemcpy(cl_out, P.cl, (size_t)nl_out * 8, cudaMemcpyDeviceToHost), "limber kernel", status))
    return;
  int st[2] = {0, 0};
  cudaMemcpy(st, P.status, sizeof(st), cudaMemcpyDeviceToHost);
  if (st[0]) {
    *status = st[0];
    set_err("ccl_cls.c: ccl_angular_cls_limber(); integration error (first underlying code %d)\n", st[1]);
  }
}

// ---- batch ----
ccl_b200_batch* ccl_b200_batch_new(int ncosmo, const ccl_b200_params* params, int* status) {
  if (*status) return nullptr;
  if (!use_device(status)) return nullptr;
  auto* b = new ccl_b200_batch();
  if (params) b->params = *params; else ccl_b200_default_params(&b->params);
  b->ts.device = current_device();
  b->ts.cosmos.resize(ncosmo);
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (!cuda_ok(cudaStreamCreateWithPriority(&b->stream, cudaStreamNonBlocking, prio_lo), "cudaStreamCreate", status) ||
      !cuda_ok(cudaStreamCreateWithPriority(&b->prep_stream, cudaStreamNonBlocking, prio_hi), "cudaStreamCreate",
               status)) {
    ccl_b200_batch_free(b);
    return nullptr;
  }
  return b;
}

void ccl_b200_batch_free(ccl_b200_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ts.device);
  if (b->stream) {
    cudaStreamSynchronize(b->stream);
    cudaStreamDestroy(b->stream);
  }
  if (b->plan_event) cudaEventDestroy(b->plan_event);
  if (b->prep_stream) {
    cudaStreamSynchronize(b->prep_stream);
    cudaStreamDestroy(b->prep_stream);
  }
  delete b;
}

#define BATCH_SLOT(b, ic)                                                      \
  if (*status) return;                                                         \
  if (!(b) || (ic) < 0 || (ic) >= (int)(b)->ts.cosmos.size()) {               \
    *status = CCL_B200_ERROR_INCONSISTENT;                                     \
    set_err("ccl_b200: batch slot %d out of range", (ic));                    \
    return;                                                                    \
  }

void ccl_b200_batch_set_achi(ccl_b200_batch* b, int icosmo, int n, const double* chi, const double* a,
                             int* status) {
  BATCH_SLOT(b, icosmo);
  if (n < 5) { *status = CCL_B200_ERROR_MEMORY; set_err("ccl_b200: a(chi) spline needs >= 5 nodes"); return; }
  b->ts.plan_s1(b->ts.cosmos[icosmo].achi, 0, n, chi, a, status);
  b->ts.committed = false;
}

void ccl_b200_batch_set_growth(ccl_b200_batch* b, int icosmo, int n, const double* a, const double* growth,
                               int* status) {
  BATCH_SLOT(b, icosmo);
  if (n < 5) { *status = CCL_B200_ERROR_MEMORY; set_err("ccl_b200: growth spline needs >= 5 nodes"); return; }
  b->ts.plan_s1(b->ts.cosmos[icosmo].growth, 0, n, a, growth, status);
  b->ts.committed = false;
}

void ccl_b200_batch_set_chi_max_nokernel(ccl_b200_batch* b, int icosmo, double chi) {
  if (b && icosmo >= 0 && icosmo < (int)b->ts.cosmos.size()) b->ts.cosmos[icosmo].chi_max_nokernel = chi;
}

void ccl_b200_batch_set_pk2d(ccl_b200_batch* b, int icosmo, const double* lk_arr, int nk,
                             const double* a_arr, int na, const double* pk_arr, int order_lok,
                             int order_hik, int is_logp, int* status) {
  BATCH_SLOT(b, icosmo);
  CosmoPlan& p = b->ts.cosmos[icosmo];
  p.has_psp = b->ts.plan_f2d(p.psp, na, a_arr, nk, lk_arr, pk_arr, nullptr, nullptr, 0, order_lok, order_hik,
                             GROWTH_CCL, is_logp, 0, 2, status);
  b->ts.committed = false;
}

int ccl_b200_batch_add_tracer(ccl_b200_batch* b, int icosmo, int der_bessel, int der_angles, int n_w,
                              const double* chi_w, const double* w_w, int na_ka, const double* a_ka,
                              int nk_ka, const double* lk_ka, const double* fka_arr, const double* fk_arr,
                              const double* fa_arr, int is_fka_log, int is_factorizable,
                              int extrap_order_lok, int extrap_order_hik, int* status) {
  if (*status) return -1;
  if (!b || icosmo < 0 || icosmo >= (int)b->ts.cosmos.size()) { *status = CCL_B200_ERROR_INCONSISTENT; return -1; }
  CosmoPlan& p = b->ts.cosmos[icosmo];
  TracerPlan t;
  if (!b->ts.plan_tracer(t, p.chi_max_nokernel, der_bessel, der_angles, n_w, chi_w, w_w, na_ka, a_ka, nk_ka,
                         lk_ka, fka_arr, fk_arr, fa_arr, is_fka_log, is_factorizable, extrap_order_lok,
                         extrap_order_hik, status))
    return -1;
  p.tracers.push_back(t);
  b->ts.committed = false;
  return (int)p.tracers.size() - 1;
}

// ---- stacked setters: arrays carry a leading cosmology axis; slot = first + i ----
#define BATCH_RANGE(b, first, count)                                                        \
  if (*status) return;                                                                      \
  if (!(b) || (first) < 0 || (count) < 0 || (first) + (count) > (int)(b)->ts.cosmos.size()) { \
    *status = CCL_B200_ERROR_INCONSISTENT;                                                  \
    set_err("ccl_b200: batch slots [%d, %d) out of range", (first), (first) + (count));    \
    return;                                                                                 \
  }                                                                                         \
  /* staging threads start with no current device: pinned-memory queries and allocations of   \
     this call must bind to the batch's own GPU, not create a context on device 0 */         \
  cudaSetDevice((b)->ts.device);

void ccl_b200_batch_set_achi_stacked(ccl_b200_batch* b, int first, int count, const int* n_per, int stride,
                                     const double* chi, const double* a, int* status) {
  BATCH_RANGE(b, first, count);
  b->ts.raw.defer = true;
  const size_t tot = (size_t)count * stride;
  const size_t l_chi = b->ts.land_if_pinned(chi, tot), l_a = b->ts.land_if_pinned(a, tot);
  auto at = [](size_t base, size_t d) { return base == NO_LAND ? NO_LAND : base + d; };
  for (int i = 0; i < count && !*status; ++i) {
    const int n = n_per ? n_per[i] : stride;
    if (n < 5 || n > stride) { *status = CCL_B200_ERROR_MEMORY; set_err("ccl_b200: a(chi) spline needs 5 <= n <= stride"); break; }
    const size_t d = (size_t)i * stride;
    b->ts.plan_s1(b->ts.cosmos[first + i].achi, 0, n, chi + d, a + d, status, at(l_chi, d), at(l_a, d));
  }
  b->ts.raw.defer = false;
  b->ts.raw.flush();
  b->ts.run_checks();
  b->ts.committed = false;
}

void ccl_b200_batch_set_growth_stacked(ccl_b200_batch* b, int first, int count, int n, int stride_a,
                                       const double* a, const double* growth, int* status) {
  BATCH_RANGE(b, first, count);
  if (n < 5) { *status = CCL_B200_ERROR_MEMORY; set_err("ccl_b200: growth spline needs >= 5 nodes"); return; }
  b->ts.raw.defer = true;
  // stride_a == 0: one scale-factor grid shared by every slot
  const size_t l_a = stride_a ? b->ts.land_if_pinned(a, (size_t)count * stride_a) : b->ts.land_put_owned(a, n);
  const size_t l_g = b->ts.land_if_pinned(growth, (size_t)count * n);
  auto at = [](size_t base, size_t d) { return base == NO_LAND ? NO_LAND : base + d; };
  for (int i = 0; i < count && !*status; ++i)
    b->ts.plan_s1(b->ts.cosmos[first + i].growth, 0, n, a + (size_t)i * stride_a, growth + (size_t)i * n, status,
                  at(l_a, (size_t)i * stride_a), at(l_g, (size_t)i * n));
  b->ts.raw.defer = false;
  b->ts.raw.flush();
  b->ts.run_checks();
  b->ts.committed = false;
}

void ccl_b200_batch_set_chi_max_nokernel_stacked(ccl_b200_batch* b, int first, int count, const double* chi) {
  if (!b || first < 0 || count < 0 || first + count > (int)b->ts.cosmos.size()) return;
  for (int i = 0; i < count; ++i) b->ts.cosmos[first + i].chi_max_nokernel = chi[i];
}

void ccl_b200_batch_set_pk2d_stacked(ccl_b200_batch* b, int first, int count, const double* lk_arr, int nk,
                                     const double* a_arr, int na, const double* pk_arr, int order_lok,
                                     int order_hik, int is_logp, int* status) {
  BATCH_RANGE(b, first, count);
  b->ts.raw.defer = true;
  // the ln k and a grids are shared by every slot: landed once instead of once per cosmology
  Lands L;
  if (lk_arr && nk > 0) L.lk = b->ts.land_put_owned(lk_arr, nk);
  if (a_arr && na > 0) L.a = b->ts.land_put_owned(a_arr, na);
  const size_t l_pk = b->ts.land_if_pinned(pk_arr, (size_t)count * na * nk);
  for (int i = 0; i < count && !*status; ++i) {
    CosmoPlan& p = b->ts.cosmos[first + i];
    L.fka = l_pk == NO_LAND ? NO_LAND : l_pk + (size_t)i * na * nk;
    p.has_psp = b->ts.plan_f2d(p.psp, na, a_arr, nk, lk_arr, pk_arr + (size_t)i * na * nk, nullptr, nullptr, 0,
                               order_lok, order_hik, GROWTH_CCL, is_logp, 0, 2, status, L);
  }
  b->ts.raw.defer = false;
  b->ts.raw.flush();
  b->ts.run_checks();
  b->ts.committed = false;
}

int ccl_b200_batch_add_tracer_stacked(ccl_b200_batch* b, int first, int count, int der_bessel, int der_angles,
                                      int n_w, int stride_chi, const double* chi_w, const double* w_w,
                                      int na_ka, int stride_a, const double* a_ka, int nk_ka,
                                      const double* lk_ka, const double* fka_arr, const double* fk_arr,
                                      const double* fa_arr, int is_fka_log, int is_factorizable,
                                      int extrap_order_lok, int extrap_order_hik, int* status) {
  if (*status) return -1;
  if (!b || first < 0 || count < 0 || first + count > (int)b->ts.cosmos.size()) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    return -1;
  }
  cudaSetDevice(b->ts.device);  // see BATCH_RANGE
  int idx = -1;
  b->ts.raw.defer = true;
  const size_t cnt = (size_t)count;
  const bool has_kernel = n_w > 0 && chi_w && w_w;
  // stride 0: one grid shared by every slot (landed once)
  const size_t l_chi = !has_kernel ? NO_LAND
                       : stride_chi ? b->ts.land_if_pinned(chi_w, cnt * stride_chi)
                                    : b->ts.land_put_owned(chi_w, n_w);
  const size_t l_w = has_kernel ? b->ts.land_if_pinned(w_w, cnt * n_w) : NO_LAND;
  const size_t l_a = !(a_ka && na_ka > 0) ? NO_LAND
                     : stride_a ? b->ts.land_if_pinned(a_ka, cnt * stride_a)
                                : b->ts.land_put_owned(a_ka, na_ka);
  const size_t l_lk = (lk_ka && nk_ka > 0) ? b->ts.land_put_owned(lk_ka, nk_ka) : NO_LAND;
  const size_t l_fka = fka_arr ? b->ts.land_if_pinned(fka_arr, cnt * na_ka * nk_ka) : NO_LAND;
  const size_t l_fk = fk_arr ? b->ts.land_if_pinned(fk_arr, cnt * nk_ka) : NO_LAND;
  const size_t l_fa = fa_arr ? b->ts.land_if_pinned(fa_arr, cnt * na_ka) : NO_LAND;
  auto at = [](size_t base, size_t d) { return base == NO_LAND ? NO_LAND : base + d; };
  for (int i = 0; i < count && !*status; ++i) {
    CosmoPlan& p = b->ts.cosmos[first + i];
    if (p.tracers.capacity() < 16) p.tracers.reserve(16);
    TracerPlan t;
    const size_t si = (size_t)i;
    Lands L;
    L.chi = at(l_chi, si * stride_chi); L.w = at(l_w, si * n_w); L.a = at(l_a, si * stride_a); L.lk = l_lk;
    L.fka = at(l_fka, si * na_ka * nk_ka); L.fk = at(l_fk, si * nk_ka); L.fa = at(l_fa, si * na_ka);
    if (!b->ts.plan_tracer(t, p.chi_max_nokernel, der_bessel, der_angles, n_w,
                           chi_w ? chi_w + si * stride_chi : nullptr, w_w ? w_w + si * n_w : nullptr, na_ka,
                           a_ka ? a_ka + si * stride_a : nullptr, nk_ka, lk_ka,
                           fka_arr ? fka_arr + si * na_ka * nk_ka : nullptr, fk_arr ? fk_arr + si * nk_ka : nullptr,
                           fa_arr ? fa_arr + si * na_ka : nullptr, is_fka_log, is_factorizable, extrap_order_lok,
                           extrap_order_hik, status, L))
      break;
    p.tracers.push_back(t);
    idx = (int)p.tracers.size() - 1;
  }
  b->ts.raw.defer = false;
  b->ts.raw.flush();
  b->ts.run_checks();
  b->ts.committed = false;
  return *status ? -1 : idx;
}

void ccl_b200_batch_tracer_chi_range(const ccl_b200_batch* b, int icosmo, int itracer, double* chi_min,
                                     double* chi_max) {
  const TracerPlan& t = b->ts.cosmos[icosmo].tracers[itracer];
  *chi_min = t.chi_min;
  *chi_max = t.chi_max;
}

void ccl_b200_batch_reset(ccl_b200_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ts.device);
  b->ts.reset();
  b->table_deps.clear();
}

static void check_table_deps(ccl_b200_batch* b, int* status);

void ccl_b200_batch_commit(ccl_b200_batch* b, int* status) {
  if (*status || !b) return;
  if (b->ts.commit(b->prep_stream, status)) check_table_deps(b, status);
}

unsigned long long ccl_b200_batch_h2d_bytes(const ccl_b200_batch* b) { return b->ts.last_h2d; }
unsigned long long ccl_b200_batch_arena_bytes(const ccl_b200_batch* b) { return b->ts.dev.cap; }

void ccl_b200_batch_angular_cls_limber_dev(ccl_b200_batch* b, int ncoll, const int* coll_start,
                                           const int* coll_count, int npair, const int* pair1,
                                           const int* pair2, int nl, const double* ell,
                                           int integration_method, double* d_cl_out, int* d_status_out,
                                           void* stream_v, int* status) {
  if (*status || !b) return;
  if (!batch_check(b, ncoll, coll_start, coll_count, npair, pair1, pair2, status)) return;
  if (integration_method != INTEG_QAG && integration_method != INTEG_SPLINE) {
    *status = CCL_B200_ERROR_NOT_IMPLEMENTED;
    return;
  }
  if (integration_method == INTEG_QAG && !qag_rule_ok(b->params, status)) return;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!cuda_ok(cudaSetDevice(b->ts.device), "cudaSetDevice", status)) return;
  const int nc = (int)b->ts.cosmos.size();
  std::vector<int2> colls(ncoll), pairs(npair);
  for (int i = 0; i < ncoll; ++i) colls[i] = make_int2(coll_start[i], coll_count[i]);
  for (int i = 0; i < npair; ++i) pairs[i] = make_int2(pair1[i], pair2[i]);
  const size_t o_co = 0;
  const size_t o_pr = o_co + ((size_t)ncoll * 8 + 255) / 256 * 256;
  const size_t o_el = o_pr + ((size_t)npair * 8 + 255) / 256 * 256;
  const size_t o_ct = o_el + ((size_t)nl * 8 + 255) / 256 * 256;
  const size_t o_fp = o_ct + 256;
  // 'qag_quad': the grid-sharing panel-cache kernel (limber_qag_fast.cu); CCL_B200_NO_QAG_FAST=1 keeps every
  // C_ell on the general one-warp-per-C_ell kernel (A/B runs)
  static const bool qag_fast = getenv("CCL_B200_NO_QAG_FAST") == nullptr;
  const bool fast = fast_eligible(b, ncoll, npair) && (integration_method == INTEG_SPLINE || qag_fast);
  const unsigned long long redo_cap = 1ull << 20;
  const size_t total = o_fp;
  if (!b->call.reserve(total, status)) return;
  char* base = b->call.p;
  bool regroup = true;
  if (fast) {
    const char* before = b->plan.p;
    if (!b->plan.reserve(limber_fast_plan_bytes(nc, npair, redo_cap), status)) return;
    std::vector<int> key;
    key.reserve(2 * (size_t)(ncoll + npair) + 2);
    key.push_back(ncoll);
    for (int i = 0; i < ncoll; ++i) { key.push_back(coll_start[i]); key.push_back(coll_count[i]); }
    key.push_back(npair);
    for (int i = 0; i < npair; ++i) { key.push_back(pair1[i]); key.push_back(pair2[i]); }
    regroup = before != b->plan.p || b->plan_commit != b->ts.commit_gen || key != b->plan_key;
    if (regroup) {
      b->plan_key.swap(key);
      b->plan_commit = b->ts.commit_gen;
    } else if (b->plan_event) {
      cudaStreamWaitEvent(stream, b->plan_event, 0);  // the plan was written on the stream of an earlier call
    }
  }
  // small pageable copies: the driver stages them before returning
  cudaMemcpyAsync(base + o_co, colls.data(), (size_t)ncoll * 8, cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(base + o_pr, pairs.data(), (size_t)npair * 8, cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(base + o_el, ell, (size_t)nl * 8, cudaMemcpyHostToDevice, stream);
  cudaMemsetAsync(base + o_ct, 0, 256, stream);
  cudaMemsetAsync(d_status_out, 0, sizeof(int) * 2 * (size_t)nc * npair, stream);
  LimberProblem P;
  P.cosmos = b->ts.d_cosmos;
  P.colls = reinterpret_cast<const int2*>(base + o_co);
  P.pairs = reinterpret_cast<const int2*>(base + o_pr);
  P.ells = reinterpret_cast<const double*>(base + o_el);
  P.ncosmo = nc; P.npair = npair; P.nl = nl; P.ncoll = ncoll;
  P.par = to_limber_params(b->params);
  P.cl = d_cl_out;
  P.status = d_status_out;
  P.status_detail = d_status_out + (size_t)nc * npair;
  P.counters = reinterpret_cast<unsigned long long*>(base + o_ct);
  double* ws = nullptr;
  size_t wsb = 0;
  DeviceScratch& sc = scratch_for(b->ts.device);
  std::unique_lock<std::mutex> lk(sc.mu, std::defer_lock);
  if (integration_method == INTEG_QAG) {
    lk.lock();
    wsb = limber_qag_workspace_bytes(P.par.N_ITERATION) + (fast ? limber_qag_fast_workspace_bytes(P.par.N_ITERATION) : 0);
    if (!sc.qag.reserve(wsb, status)) return;
    ws = reinterpret_cast<double*>(sc.qag.p);
  }
  b->last_fast = fast;
  if (fast) {
    FastPlan plan = limber_fast_plan_carve(b->plan.p, nc, npair, redo_cap);
    plan.regroup = regroup ? 1 : 0;
    plan.ntr_max = 1;
    plan.has_rsd = 0;
    for (const CosmoPlan& c : b->ts.cosmos) {
      plan.ntr_max = std::max(plan.ntr_max, (int)c.tracers.size());
      for (const TracerPlan& t : c.tracers) plan.has_rsd |= (t.der_bessel > 0);
    }
    if (integration_method == INTEG_SPLINE) {
      if (!cuda_ok(launch_limber_fast(P, plan, stream), "limber fast launch", status)) return;
      if (regroup) {
        if (!b->plan_event) cudaEventCreateWithFlags(&b->plan_event, cudaEventDisableTiming);
        cudaEventRecord(b->plan_event, stream);
      }
      return;
    }
    // 'qag_quad': grid-sharing panels, then the general kernel over the redo list, then CQUAD
    if (!cuda_ok(launch_limber_qag_fast(P, plan, ws, wsb, stream), "limber qag fast launch", status)) return;
    if (regroup) {
      if (!b->plan_event) cudaEventCreateWithFlags(&b->plan_event, cudaEventDisableTiming);
      cudaEventRecord(b->plan_event, stream);
    }
  } else if (!cuda_ok(launch_limber(P, integration_method, ws, wsb, stream), "limber launch", status))
    return;
  if (integration_method == INTEG_QAG) {
    // the shared workspace must not be reused before this launch has drained
    cuda_ok(cudaStreamSynchronize(stream), "limber qag kernel", status);
    unsigned long long cnt[2] = {0, 0};
    cudaMemcpy(cnt, P.counters, sizeof(cnt), cudaMemcpyDeviceToHost);
    b->last_qag_evals = cnt[0];
    b->last_qag_nodes = fast ? cnt[1] : cnt[0];
  }
}

void ccl_b200_batch_angular_cls_limber_start(ccl_b200_batch* b, int ncoll, const int* coll_start,
                                             const int* coll_count, int npair, const int* pair1,
                                             const int* pair2, int nl, const double* ell, int integration_method,
                                             double* cl_out, int* status) {
  if (*status || !b) return;
  if (!cuda_ok(cudaSetDevice(b->ts.device), "cudaSetDevice", status)) return;
  const size_t nc = b->ts.cosmos.size();
  const size_t ncl = nc * npair * (size_t)nl;
  const size_t nst = nc * (size_t)npair;
  b->pending_status = 0;
  if (!b->out.reserve(ncl * 8 + nst * 2 * sizeof(int) + 512, status)) return;
  double* d_cl = reinterpret_cast<double*>(b->out.p);
  int* d_st = reinterpret_cast<int*>(b->out.p + (ncl * 8 + 255) / 256 * 256);
  ccl_b200_batch_angular_cls_limber_dev(b, ncoll, coll_start, coll_count, npair, pair1, pair2, nl, ell,
                                        integration_method, d_cl, d_st, b->stream, status);
  if (*status) return;
  // status words land in a page-locked buffer of the batch (2 ints per double slot)
  if (!b->host_status.pinned || b->host_status.cap * 2 < nst) {
    b->host_status.release();
    double* q = nullptr;
    const size_t nd = std::max<size_t>((nst + 1) / 2, 64);
    if (cudaHostAlloc((void**)&q, nd * sizeof(double), cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      *status = CCL_B200_ERROR_MEMORY;
      set_err("ccl_b200: cudaHostAlloc(status buffer) failed");
      return;
    }
    b->host_status.p = q; b->host_status.cap = nd; b->host_status.pinned = true;
  }
  if (ncl) cudaMemcpyAsync(cl_out, d_cl, ncl * 8, cudaMemcpyDeviceToHost, b->stream);
  if (nst) cudaMemcpyAsync(b->host_status.p, d_st, nst * sizeof(int), cudaMemcpyDeviceToHost, b->stream);
  b->pending_status = nst;
}

void ccl_b200_batch_angular_cls_limber_wait(ccl_b200_batch* b, int* status_out, int* status) {
  if (*status || !b) return;
  if (!cuda_ok(cudaSetDevice(b->ts.device), "cudaSetDevice", status)) return;
  if (!cuda_ok(cudaStreamSynchronize(b->stream), "limber kernel", status)) return;
  const int* st = reinterpret_cast<const int*>(b->host_status.p);
  bool any = false;
  for (size_t i = 0; i < b->pending_status; ++i) {
    if (status_out) status_out[i] = st[i];
    any = any || st[i];
  }
  b->pending_status = 0;
  if (any) {
    *status = CCL_B200_ERROR_INTEG;
    set_err("ccl_cls.c: ccl_angular_cls_limber(); integration error\n");
  }
}

void ccl_b200_batch_angular_cls_limber(ccl_b200_batch* b, int ncoll, const int* coll_start,
                                       const int* coll_count, int npair, const int* pair1,
                                       const int* pair2, int nl, const double* ell, int integration_method,
                                       double* cl_out, int* status_out, int* status) {
  ccl_b200_batch_angular_cls_limber_start(b, ncoll, coll_start, coll_count, npair, pair1, pair2, nl, ell,
                                          integration_method, cl_out, status);
  ccl_b200_batch_angular_cls_limber_wait(b, status_out, status);
}

unsigned long long ccl_b200_batch_spline_nodes(const ccl_b200_batch* b, int ncoll, const int* coll_start,
                                               const int* coll_count, int npair, const int* pair1,
                                               const int* pair2, int nl, const double* ell) {
  unsigned long long tot = 0;
  const ccl_b200_params& p = b->params;
  for (const CosmoPlan& c : b->ts.cosmos) {
    std::vector<double> cmin(ncoll, 1e15), cmax(ncoll, -1e15);
    for (int k = 0; k < ncoll; ++k)
      for (int t = coll_start[k]; t < coll_start[k] + coll_count[k]; ++t) {
        cmin[k] = std::min(cmin[k], c.tracers[t].chi_min);
        cmax[k] = std::max(cmax[k], c.tracers[t].chi_max);
      }
    for (int q = 0; q < npair; ++q) {
      const double chi_max = std::min(cmax[pair1[q]], cmax[pair2[q]]);
      const double chi_min0 = std::max(cmin[pair1[q]], cmin[pair2[q]]);
      for (int l = 0; l < nl; ++l) {
        double chi_min = chi_min0;
        if (chi_min <= 0) chi_min = 0.5 * (ell[l] + 0.5) / p.K_MAX;
        const double lkmax = log(std::min(p.K_MAX, 2 * (ell[l] + 0.5) / chi_min));
        const double lkmin = log(std::max(p.K_MIN, (ell[l] + 0.5) / chi_max));
        const int nk = (int)(std::max((lkmax - lkmin) / p.DLOGK_INTEGRATION + 0.5, 1.0)) + 1;
        if (nk >= 5) tot += nk;
      }
    }
  }
  return tot;
}

unsigned long long ccl_b200_batch_last_qag_evals(const ccl_b200_batch* b) { return b->last_qag_evals; }
unsigned long long ccl_b200_batch_last_qag_nodes_computed(const ccl_b200_batch* b) { return b->last_qag_nodes; }

int ccl_b200_batch_last_used_fast_path(const ccl_b200_batch* b) { return b->last_fast ? 1 : 0; }

// ---- device builders ----
void ccl_b200_build_background(int ncosmo, const double* params, int na, const double* a_grid, double dchi,
                               double root_epsrel, int root_niter, double a_growth_init, int stride, double* E,
                               double* chi, int* n_achi, double* achi_chi, double* achi_a, double* growth,
                               double* fgrowth, double* growth0, int* status) {
  if (*status) return;
  if (!use_device(status)) return;
  if (ncosmo <= 0 || na < 5 || stride < 5 || a_grid[na - 1] != 1.0) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_b200_build_background: need ncosmo > 0, na >= 5, stride >= 5 and an a-grid ending at 1");
    return;
  }
  const size_t nc = (size_t)ncosmo, sna = (size_t)na, sst = (size_t)stride;
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t o_par = 0, o_a = o_par + up(nc * CCL_B200_NPAR * 8), o_E = o_a + up(sna * 8),
               o_chi = o_E + up(nc * sna * 8), o_n = o_chi + up(nc * sna * 8), o_gx = o_n + up(nc * 4),
               o_ga = o_gx + up(nc * sst * 8), o_D = o_ga + up(nc * sst * 8), o_f = o_D + up(nc * sna * 8),
               o_g0 = o_f + up(nc * sna * 8), total = o_g0 + up(nc * 8);
  DevBuf buf;
  if (!buf.reserve(total, status)) return;
  char* d = buf.p;
  cudaMemcpy(d + o_par, params, nc * CCL_B200_NPAR * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_a, a_grid, sna * 8, cudaMemcpyHostToDevice);
  cudaMemset(d + o_gx, 0, 2 * up(nc * sst * 8));
  launch_build_background(ncosmo, na, (double*)(d + o_a), (double*)(d + o_par), dchi, root_epsrel, root_niter,
                          a_growth_init, stride, (double*)(d + o_E), (double*)(d + o_chi), (int*)(d + o_n),
                          (double*)(d + o_gx), (double*)(d + o_ga), (double*)(d + o_D), (double*)(d + o_f),
                          (double*)(d + o_g0), 0);
  if (!cuda_ok(cudaGetLastError(), "background builder launch", status)) return;
  if (!cuda_ok(cudaDeviceSynchronize(), "background builder", status)) return;
  cudaMemcpy(E, d + o_E, nc * sna * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(chi, d + o_chi, nc * sna * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(n_achi, d + o_n, nc * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(achi_chi, d + o_gx, nc * sst * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(achi_a, d + o_ga, nc * sst * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(growth, d + o_D, nc * sna * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(fgrowth, d + o_f, nc * sna * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(growth0, d + o_g0, nc * 8, cudaMemcpyDeviceToHost);
  for (int i = 0; i < ncosmo; ++i)
    if (n_achi[i] < 5) {
      *status = CCL_B200_ERROR_MEMORY;
      set_err("ccl_b200_build_background: cosmology %d needs %d a(chi) nodes, stride is %d", i, -n_achi[i], stride);
      return;
    }
}

void ccl_b200_build_linear_power(int ncosmo, const double* params, int model, int nk, const double* k_grid,
                                 int na_pk, const double* a_pk, int na_bg, const double* a_bg,
                                 const double* growth, double k_min, double k_max, double* logp,
                                 double* sigma8_unnorm, int* status) {
  if (*status) return;
  if (!use_device(status)) return;
  if (ncosmo <= 0 || nk < 4 || na_pk < 1 || na_bg < 5 || model < 0 || model > 2) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_b200_build_linear_power: inconsistent arguments");
    return;
  }
  const size_t nc = (size_t)ncosmo;
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t o_par = 0, o_k = o_par + up(nc * CCL_B200_NPAR * 8), o_apk = o_k + up((size_t)nk * 8),
               o_abg = o_apk + up((size_t)na_pk * 8), o_D = o_abg + up((size_t)na_bg * 8),
               o_lp = o_D + up(nc * na_bg * 8), o_s8 = o_lp + up(nc * na_pk * nk * 8), total = o_s8 + up(nc * 8);
  DevBuf buf;
  if (!buf.reserve(total, status)) return;
  char* d = buf.p;
  cudaMemcpy(d + o_par, params, nc * CCL_B200_NPAR * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_k, k_grid, (size_t)nk * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_apk, a_pk, (size_t)na_pk * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_abg, a_bg, (size_t)na_bg * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_D, growth, nc * na_bg * 8, cudaMemcpyHostToDevice);
  launch_build_linpower(ncosmo, nk, (double*)(d + o_k), na_pk, (double*)(d + o_apk), na_bg, (double*)(d + o_abg),
                        (double*)(d + o_D), (double*)(d + o_par), model, log10(k_min), log10(k_max),
                        (double*)(d + o_lp), (double*)(d + o_s8), 0);
  if (!cuda_ok(cudaGetLastError(), "linear power builder launch", status)) return;
  if (!cuda_ok(cudaDeviceSynchronize(), "linear power builder", status)) return;
  cudaMemcpy(logp, d + o_lp, nc * na_pk * nk * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(sigma8_unnorm, d + o_s8, nc * 8, cudaMemcpyDeviceToHost);
}

void ccl_b200_build_tracer_kernels(int ncosmo, const double* params, int na, const double* a_bg,
                                   const double* chi_bg, const int* n_achi, int stride, const double* achi_chi,
                                   const double* achi_a, int nz, const double* z, int ntr, const double* nz_arr,
                                   const double* sz_arr, const double* inv_norm, int nchi, double* chi_z,
                                   double* w_nc, double* chi_l, double* w_l, int* status) {
  if (*status) return;
  if (!use_device(status)) return;
  if (ncosmo <= 0 || ntr <= 0 || nz < 5 || na < 5 || nchi < 1 || nchi > 256 || ntr > 65535) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_b200_build_tracer_kernels: need nz >= 5, 1 <= nchi <= 256, ntr <= 65535");
    return;
  }
  const size_t nc = (size_t)ncosmo, snz = (size_t)nz, st = (size_t)stride, nt = (size_t)ntr;
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += up(bytes); return o; };
  const size_t o_par = take(nc * CCL_B200_NPAR * 8), o_a = take((size_t)na * 8), o_chi = take(nc * na * 8),
               o_n = take(nc * 4), o_gx = take(nc * st * 8), o_ga = take(nc * st * 8), o_z = take(snz * 8),
               o_nz = take(nt * snz * 8), o_sz = take(nt * snz * 8), o_in = take(nt * 8), o_cz = take(nc * snz * 8),
               o_wn = take(nc * nt * snz * 8), o_cl = take(nc * nchi * 8), o_wl = take(nc * nt * nchi * 8);
  DevBuf buf;
  if (!buf.reserve(off, status)) return;
  char* d = buf.p;
  cudaMemcpy(d + o_par, params, nc * CCL_B200_NPAR * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_a, a_bg, (size_t)na * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_chi, chi_bg, nc * na * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_n, n_achi, nc * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_gx, achi_chi, nc * st * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_ga, achi_a, nc * st * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_z, z, snz * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_nz, nz_arr, nt * snz * 8, cudaMemcpyHostToDevice);
  if (sz_arr) cudaMemcpy(d + o_sz, sz_arr, nt * snz * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_in, inv_norm, nt * 8, cudaMemcpyHostToDevice);
  launch_build_tracer_kernels(ncosmo, na, (double*)(d + o_a), (double*)(d + o_chi), (int*)(d + o_n), stride,
                              (double*)(d + o_gx), (double*)(d + o_ga), (double*)(d + o_par), nz, (double*)(d + o_z),
                              ntr, (double*)(d + o_nz), sz_arr ? (double*)(d + o_sz) : nullptr, (double*)(d + o_in),
                              nullptr, nchi, (double*)(d + o_cz), (double*)(d + o_wn), (double*)(d + o_cl), (double*)(d + o_wl), 0);
  if (!cuda_ok(cudaGetLastError(), "tracer kernel builder launch", status)) return;
  if (!cuda_ok(cudaDeviceSynchronize(), "tracer kernel builder", status)) return;
  cudaMemcpy(chi_z, d + o_cz, nc * snz * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(w_nc, d + o_wn, nc * nt * snz * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(chi_l, d + o_cl, nc * nchi * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(w_l, d + o_wl, nc * nt * nchi * 8, cudaMemcpyDeviceToHost);
}

void ccl_b200_build_halofit(int ncosmo, const double* params, int nk, const double* lk, int na, const double* a,
                            const double* logplin, double* logpnl, double* rsigma, int* status) {
  if (*status) return;
  if (!use_device(status)) return;
  if (ncosmo <= 0 || nk < 4 || na < 1) { *status = CCL_B200_ERROR_INCONSISTENT; return; }
  const size_t nc = (size_t)ncosmo, tab = nc * na * nk * 8;
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += up(bytes); return o; };
  const size_t o_par = take(nc * CCL_B200_NPAR * 8), o_lk = take((size_t)nk * 8), o_a = take((size_t)na * 8),
               o_in = take(tab), o_out = take(tab), o_rs = take(nc * na * 8), o_f = take(4);
  DevBuf buf;
  if (!buf.reserve(off, status)) return;
  char* d = buf.p;
  cudaMemcpy(d + o_par, params, nc * CCL_B200_NPAR * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_lk, lk, (size_t)nk * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_a, a, (size_t)na * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_in, logplin, tab, cudaMemcpyHostToDevice);
  cudaMemset(d + o_f, 0, 4);
  launch_build_halofit(ncosmo, nk, (double*)(d + o_lk), na, (double*)(d + o_a), (double*)(d + o_in),
                       (double*)(d + o_par), (double*)(d + o_out), (double*)(d + o_rs), (int*)(d + o_f), 0);
  if (!cuda_ok(cudaGetLastError(), "halofit builder launch", status)) return;
  if (!cuda_ok(cudaDeviceSynchronize(), "halofit builder", status)) return;
  cudaMemcpy(logpnl, d + o_out, tab, cudaMemcpyDeviceToHost);
  cudaMemcpy(rsigma, d + o_rs, nc * na * 8, cudaMemcpyDeviceToHost);
  int failed = 0;
  cudaMemcpy(&failed, d + o_f, 4, cudaMemcpyDeviceToHost);
  if (failed) {
    *status = CCL_B200_ERROR_INTEG;
    set_err("ccl_halofit.c: could not solve for non-linear scale for halofit");
  }
}

// ---- device-resident tables ----
ccl_b200_tables* ccl_b200_tables_build(int ncosmo, const double* params, int na, const double* a_grid, double dchi,
                                       double root_epsrel, int root_niter, double a_growth_init, int achi_stride,
                                       double a_spline_min, int model, int halofit, int nk, const double* k_grid,
                                       const double* lk_grid, int na_pk, const double* a_pk, double k_min,
                                       double k_max, int nz,
                                       const double* z, int ntr, const double* nz_arr, const double* sz_arr,
                                       const double* inv_norm, const double* lens_scale, int nchi, int* status) {
  if (*status) return nullptr;
  if (!use_device(status)) return nullptr;
  if (ncosmo <= 0 || na < 5 || achi_stride < 5 || a_grid[na - 1] != 1.0 || nk < 4 || na_pk < 4 || model < 0 ||
      model > 2 || (ntr > 0 && (nz < 5 || nchi < 5 || nchi > 256)) || ntr < 0 || ntr > 65535) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_b200_tables_build: inconsistent arguments");
    return nullptr;
  }
  auto* t = new ccl_b200_tables();
  t->device = current_device();
  t->ncosmo = ncosmo; t->na = na; t->stride = achi_stride; t->nk = nk; t->na_pk = na_pk; t->nz = nz; t->ntr = ntr;
  t->nchi = nchi; t->has_nl = halofit != 0;
  const size_t nc = (size_t)ncosmo, sna = (size_t)na, sst = (size_t)achi_stride, tab = nc * na_pk * nk, nt = (size_t)ntr;
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += up(bytes); return o; };
  const size_t o_par = take(nc * CCL_B200_NPAR * 8), o_a = take(sna * 8), o_E = take(nc * sna * 8),
               o_chi = take(nc * sna * 8), o_n = take(nc * 4), o_gx = take(nc * sst * 8), o_ga = take(nc * sst * 8),
               o_D = take(nc * sna * 8), o_f = take(nc * sna * 8), o_g0 = take(nc * 8), o_k = take((size_t)nk * 8),
               o_lk = take((size_t)nk * 8), o_apk = take((size_t)na_pk * 8), o_lin = take(tab * 8),
               o_nl = take(halofit ? tab * 8 : 8), o_s8 = take(nc * 8), o_rs = take(nc * na_pk * 8), o_fail = take(4),
               o_z = take((size_t)nz * 8), o_nz = take(nt * nz * 8), o_sz = take(nt * nz * 8), o_in = take(nt * 8),
               o_cz = take(nc * nz * 8), o_wn = take(nc * nt * nz * 8), o_cl = take(nc * nchi * 8),
               o_wl = take(nc * nt * nchi * 8), o_rng = take(nc * (nt ? nt : 1) * 4 * 8 * 2), o_cm = take(nc * 8);
  if (!t->buf.reserve(off, status)) { delete t; return nullptr; }
  char* d = t->buf.p;
  auto D = [&](size_t o) { return reinterpret_cast<double*>(d + o); };
  t->d_par = D(o_par); t->d_a = D(o_a); t->d_E = D(o_E); t->d_chi = D(o_chi); t->d_nachi = (int*)(d + o_n);
  t->d_gx = D(o_gx); t->d_ga = D(o_ga); t->d_D = D(o_D); t->d_f = D(o_f); t->d_g0 = D(o_g0); t->d_k = D(o_k);
  t->d_lk = D(o_lk); t->d_apk = D(o_apk); t->d_lin = D(o_lin); t->d_nl = D(o_nl); t->d_s8 = D(o_s8); t->d_rs = D(o_rs);
  t->d_fail = (int*)(d + o_fail); t->d_z = D(o_z); t->d_nz = D(o_nz); t->d_sz = D(o_sz); t->d_in = D(o_in);
  t->d_chiz = D(o_cz); t->d_wnc = D(o_wn); t->d_chil = D(o_cl); t->d_wl = D(o_wl); t->d_rng = D(o_rng); t->d_cmax = D(o_cm);
  cudaStream_t st = 0;
  cudaMemcpyAsync(t->d_par, params, nc * CCL_B200_NPAR * 8, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(t->d_a, a_grid, sna * 8, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(t->d_k, k_grid, (size_t)nk * 8, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(t->d_apk, a_pk, (size_t)na_pk * 8, cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(t->d_fail, 0, 4, st);
  cudaMemcpyAsync(t->d_lk, lk_grid, (size_t)nk * 8, cudaMemcpyHostToDevice, st);  // the caller's ln k nodes, verbatim
  launch_build_background(ncosmo, na, t->d_a, t->d_par, dchi, root_epsrel, root_niter, a_growth_init, achi_stride,
                          t->d_E, t->d_chi, t->d_nachi, t->d_gx, t->d_ga, t->d_D, t->d_f, t->d_g0, st);
  launch_chi_at(ncosmo, na, t->d_a, t->d_chi, a_spline_min, t->d_cmax, st);
  launch_build_linpower(ncosmo, nk, t->d_k, na_pk, t->d_apk, na, t->d_a, t->d_D, t->d_par, model, log10(k_min),
                        log10(k_max), t->d_lin, t->d_s8, st);
  if (ntr > 0) {
    cudaMemcpyAsync(t->d_z, z, (size_t)nz * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(t->d_nz, nz_arr, nt * nz * 8, cudaMemcpyHostToDevice, st);
    if (sz_arr) cudaMemcpyAsync(t->d_sz, sz_arr, nt * nz * 8, cudaMemcpyHostToDevice, st);
    std::vector<double> eff(nt);
    for (size_t i = 0; i < nt; ++i) eff[i] = inv_norm[i];
    cudaMemcpy(t->d_in, eff.data(), nt * 8, cudaMemcpyHostToDevice);
    // per-distribution factor of the lensing kernels (e.g. -2 for magnification); 0 = kernel not needed,
    // its integrals are skipped
    double* d_sc = nullptr;
    if (lens_scale) {
      d_sc = t->d_rng;  // scratch until the ranges are computed
      cudaMemcpy(d_sc, lens_scale, nt * 8, cudaMemcpyHostToDevice);
    }
    launch_build_tracer_kernels(ncosmo, na, t->d_a, t->d_chi, t->d_nachi, achi_stride, t->d_gx, t->d_ga, t->d_par, nz,
                                t->d_z, ntr, t->d_nz, sz_arr ? t->d_sz : nullptr, t->d_in, d_sc, nchi, t->d_chiz,
                                t->d_wnc, t->d_chil, t->d_wl, st);
    if (lens_scale) launch_scale_rows(ncosmo, ntr, nchi, t->d_wl, d_sc, st);
    launch_tracer_ranges(ncosmo, ntr, nz, t->d_chiz, 0, t->d_wnc, t->d_rng, st);
    launch_tracer_ranges(ncosmo, ntr, nchi, t->d_chil, 0, t->d_wl, t->d_rng + nc * nt * 4, st);
  }
  if (halofit) {
    cudaEvent_t ev_lin = nullptr;
    cudaStreamCreateWithFlags(&t->s_nl, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ev_lin, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&t->ev_nl, cudaEventDisableTiming);
    // launched after the tracer kernels so that their CTAs are dispatched first: the host waits for those (and
    // the ranges) only, and plans the batch while halofit runs
    cudaEventRecord(ev_lin, st);
    cudaStreamWaitEvent(t->s_nl, ev_lin, 0);
    launch_build_halofit(ncosmo, nk, t->d_lk, na_pk, t->d_apk, t->d_lin, t->d_par, t->d_nl, t->d_rs, t->d_fail, t->s_nl);
    cudaEventRecord(t->ev_nl, t->s_nl);
    cudaEventDestroy(ev_lin);
  }
  if (!cuda_ok(cudaGetLastError(), "table builder launch", status)) { delete t; return nullptr; }
  if (!cuda_ok(cudaStreamSynchronize(st), "table builders", status)) { delete t; return nullptr; }
  // host metadata (small): sizes and end nodes
  t->h_a.assign(a_grid, a_grid + na);
  t->h_apk.assign(a_pk, a_pk + na_pk);
  t->h_lk.assign(lk_grid, lk_grid + nk);
  t->h_nachi.resize(nc);
  cudaMemcpy(t->h_nachi.data(), t->d_nachi, nc * 4, cudaMemcpyDeviceToHost);
  t->h_chi_amin.resize(nc);
  cudaMemcpy2D(t->h_chi_amin.data(), 8, t->d_chi, sna * 8, 8, nc, cudaMemcpyDeviceToHost);  // chi(a_grid[0])
  t->h_cmax.resize(nc);
  cudaMemcpy(t->h_cmax.data(), t->d_cmax, nc * 8, cudaMemcpyDeviceToHost);
  if (ntr > 0) {
    t->h_rng_nc.resize(nc * nt * 4);
    t->h_rng_l.resize(nc * nt * 4);
    cudaMemcpy(t->h_rng_nc.data(), t->d_rng, nc * nt * 4 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(t->h_rng_l.data(), t->d_rng + nc * nt * 4, nc * nt * 4 * 8, cudaMemcpyDeviceToHost);
  }
  for (size_t i = 0; i < nc; ++i)
    if (t->h_nachi[i] < 5) {
      *status = CCL_B200_ERROR_MEMORY;
      set_err("ccl_b200_tables_build: cosmology %zu needs %d a(chi) nodes, stride is %d", i, -t->h_nachi[i], achi_stride);
      delete t;
      return nullptr;
    }
  return t;
}

// the halofit pass of a table set has finished (its consumers were ordered behind ev_nl): report its failure flag
static void check_table_deps(ccl_b200_batch* b, int* status) {
  for (const ccl_b200_tables* t : b->table_deps) {
    if (!t->ev_nl) continue;
    cudaEventSynchronize(t->ev_nl);
    int failed = 0;
    cudaMemcpy(&failed, t->d_fail, 4, cudaMemcpyDeviceToHost);
    if (failed && !*status) {
      *status = CCL_B200_ERROR_INTEG;
      set_err("ccl_halofit.c: could not solve for non-linear scale for halofit");
    }
  }
  b->table_deps.clear();
}

void ccl_b200_tables_free(ccl_b200_tables* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  delete t;
}

void ccl_b200_batch_set_from_tables(ccl_b200_batch* b, int first, const ccl_b200_tables* t, int order_lok,
                                    int order_hik, int use_linear, int* status) {
  if (*status) return;
  if (!b || !t || first < 0 || first + t->ncosmo > (int)b->ts.cosmos.size() || b->ts.device != t->device) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_b200_batch_set_from_tables: slot range or device mismatch");
    return;
  }
  const double* pk = (t->has_nl && !use_linear) ? t->d_nl : t->d_lin;
  if (t->ev_nl && pk == t->d_nl) {  // the halofit table may still be in the making: order this batch's streams behind it
    cudaStreamWaitEvent(b->prep_stream, t->ev_nl, 0);
    cudaStreamWaitEvent(b->stream, t->ev_nl, 0);
    b->table_deps.push_back(t);
  }
  const int gu = TableSet::quasi_uniform(t->h_a.data(), t->na);
  for (int i = 0; i < t->ncosmo; ++i) {
    CosmoPlan& p = b->ts.cosmos[first + i];
    const size_t si = (size_t)i;
    // a(chi): uniform grid from chi(a = 1) = 0 to chi(a_min)
    b->ts.plan_s1_dev(p.achi, 0, t->h_nachi[i], 0.0, t->h_chi_amin[i], 1, t->d_gx + si * t->stride, t->d_ga + si * t->stride);
    b->ts.plan_s1_dev(p.growth, 0, t->na, t->h_a[0], t->h_a[t->na - 1], gu, t->d_a, t->d_D + si * t->na);
    p.chi_max_nokernel = t->h_cmax[i];
    b->ts.plan_psp_dev(p.psp, t->nk, t->h_lk.data(), t->d_lk, t->na_pk, t->h_apk.data(), t->d_apk,
                       pk + si * t->na_pk * t->nk, order_lok, order_hik, 1);
    p.has_psp = true;
  }
  b->ts.committed = false;
}

int ccl_b200_batch_add_tracer_from_tables(ccl_b200_batch* b, int first, const ccl_b200_tables* t, int kind, int itr,
                                          int der_bessel, int der_angles, int na_ka, const double* a_ka,
                                          const double* fa_arr, int* status) {
  if (*status) return -1;
  if (!b || !t || first < 0 || first + t->ncosmo > (int)b->ts.cosmos.size() || itr < 0 || itr >= t->ntr ||
      (kind != 0 && kind != 1) || der_angles < 0 || der_angles > 2 || der_bessel < -1 || der_bessel > 2) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_b200_batch_add_tracer_from_tables: inconsistent arguments");
    return -1;
  }
  const int n = kind ? t->nchi : t->nz;
  const std::vector<double>& rng = kind ? t->h_rng_l : t->h_rng_nc;
  int idx = -1;
  b->ts.raw.defer = true;
  // the a-only transfer is shared by every slot: landed once, its constancy checked once
  Lands L;
  bool same = false;
  if (fa_arr != nullptr && a_ka != nullptr && na_ka > 0) {
    L.a = b->ts.land_put_owned(a_ka, na_ka);
    L.fa = b->ts.land_put_owned(fa_arr, na_ka);
    same = true;
    for (int j = 1; j < na_ka && same; ++j) same = (fa_arr[j] == fa_arr[0]);
  }
  for (int i = 0; i < t->ncosmo && !*status; ++i) {
    CosmoPlan& p = b->ts.cosmos[first + i];
    if (p.tracers.capacity() < 16) p.tracers.reserve(16);
    const size_t si = (size_t)i, ti = si * t->ntr + itr;
    TracerPlan tp;
    tp.der_bessel = der_bessel;
    tp.der_angles = der_angles;
    tp.chi_min = rng[ti * 4 + 0];
    tp.chi_max = rng[ti * 4 + 1];
    const double* d_x = kind ? t->d_chil + si * n : t->d_chiz + si * n;
    const double* d_w = (kind ? t->d_wl : t->d_wnc) + ti * n;
    b->ts.plan_s1_dev(tp.kernel, 0, n, rng[ti * 4 + 2], rng[ti * 4 + 3], 0, d_x, d_w);
    if (fa_arr != nullptr && a_ka != nullptr) {
      tp.has_transfer = true;
      if (!b->ts.plan_f2d(tp.transfer, na_ka, a_ka, 0, nullptr, nullptr, nullptr, fa_arr, 1, 0, 2, GROWTH_CONST, 0, 1.0,
                          0, status, L))
        break;
      tp.const_transfer = same;
      tp.const_transfer_value = fa_arr[0];
    }
    p.tracers.push_back(tp);
    idx = (int)p.tracers.size() - 1;
  }
  b->ts.raw.defer = false;
  b->ts.raw.flush();
  b->ts.run_checks();
  b->ts.committed = false;
  return *status ? -1 : idx;
}

int ccl_b200_batch_add_tracer_from_tables_bg(ccl_b200_batch* b, int first, ccl_b200_tables* t, int kind, int itr,
                                             int der_bessel, int der_angles, int bg_kind, int na_ka,
                                             const double* a_ka, const double* amp, int* status) {
  if (*status) return -1;
  if (!b || !t || first < 0 || first + t->ncosmo > (int)b->ts.cosmos.size() || itr < 0 || itr >= t->ntr ||
      (kind != 0 && kind != 1) || der_angles < 0 || der_angles > 2 || der_bessel < -1 || der_bessel > 2 ||
      (bg_kind != 1 && bg_kind != 2) || na_ka < 3 || !a_ka || (bg_kind == 2 && !amp)) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_b200_batch_add_tracer_from_tables_bg: inconsistent arguments");
    return -1;
  }
  for (int j = 1; j < na_ka; ++j)
    if (!(a_ka[j] > a_ka[j - 1])) {
      *status = CCL_B200_ERROR_SPLINE;
      set_err("ccl_b200: spline nodes must be strictly increasing");
      return -1;
    }
  if (a_ka[0] < t->h_a.front() || a_ka[na_ka - 1] > t->h_a.back()) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_b200: transfer nodes outside the background a-spline");
    return -1;
  }
  if (!cuda_ok(cudaSetDevice(t->device), "cudaSetDevice", status)) return -1;
  const size_t nc = (size_t)t->ncosmo, n = (size_t)na_ka;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  t->extras.emplace_back();
  DevBuf& buf = t->extras.back();
  const size_t o_a = 0, o_amp = up(n * 8), o_fa = o_amp + up(n * 8);
  if (!buf.reserve(o_fa + up(nc * n * 8), status)) return -1;
  double* d_a = reinterpret_cast<double*>(buf.p + o_a);
  double* d_amp = reinterpret_cast<double*>(buf.p + o_amp);
  double* d_fa = reinterpret_cast<double*>(buf.p + o_fa);
  cudaMemcpy(d_a, a_ka, n * 8, cudaMemcpyHostToDevice);
  if (amp) cudaMemcpy(d_amp, amp, n * 8, cudaMemcpyHostToDevice);
  launch_build_bg_transfer(t->ncosmo, t->na, t->d_a, t->d_D, t->d_f, t->d_par, bg_kind, na_ka, d_a, d_amp, d_fa, 0);
  if (!cuda_ok(cudaGetLastError(), "transfer builder launch", status)) return -1;
  if (!cuda_ok(cudaDeviceSynchronize(), "transfer builder", status)) return -1;
  const int nk = kind ? t->nchi : t->nz;
  const std::vector<double>& rng = kind ? t->h_rng_l : t->h_rng_nc;
  const int a_uniform = TableSet::quasi_uniform(a_ka, na_ka);
  int idx = -1;
  for (size_t i = 0; i < nc; ++i) {
    CosmoPlan& p = b->ts.cosmos[first + i];
    if (p.tracers.capacity() < 16) p.tracers.reserve(16);
    const size_t ti = i * t->ntr + itr;
    TracerPlan tp;
    tp.der_bessel = der_bessel;
    tp.der_angles = der_angles;
    tp.chi_min = rng[ti * 4 + 0];
    tp.chi_max = rng[ti * 4 + 1];
    const double* d_x = kind ? t->d_chil + i * nk : t->d_chiz + i * nk;
    const double* d_w = (kind ? t->d_wl : t->d_wnc) + ti * nk;
    b->ts.plan_s1_dev(tp.kernel, 0, nk, rng[ti * 4 + 2], rng[ti * 4 + 3], 0, d_x, d_w);
    // the flags ccl_f2d_t_new gives a factorised a-only transfer (ccl_cl_tracer_t_new, src/ccl_tracers.c:670-684)
    tp.has_transfer = true;
    F2DPlan& f = tp.transfer;
    memset(&f.flags, 0, sizeof(F2D));
    F2D& g = f.flags;
    g.is_factorizable = 1; g.is_k_constant = 1; g.is_a_constant = 0;
    g.extrap_order_lok = 0; g.extrap_order_hik = 2; g.extrap_linear_growth = GROWTH_CONST;
    g.is_log = 0; g.growth_factor_0 = 1.0; g.growth_exponent = 0;
    g.amin = a_ka[0]; g.amax = a_ka[na_ka - 1];
    g.has_fa = 1;
    b->ts.plan_s1_dev(f.fa, 1, na_ka, a_ka[0], a_ka[na_ka - 1], a_uniform, d_a, d_fa + i * n);
    p.tracers.push_back(tp);
    idx = (int)p.tracers.size() - 1;
  }
  b->ts.committed = false;
  return idx;
}

int ccl_b200_batch_add_kappa_tracer_from_tables(ccl_b200_batch* b, int first, ccl_b200_tables* t, double z_source,
                                                int n_samples, int* status) {
  if (*status) return -1;
  if (!b || !t || first < 0 || first + t->ncosmo > (int)b->ts.cosmos.size() || n_samples < 5 || !(z_source > 0)) {
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_b200_batch_add_kappa_tracer_from_tables: inconsistent arguments (n_samples >= 5, z_source > 0)");
    return -1;
  }
  if (1. / (1. + z_source) < t->h_a.front()) {  // pyccl: _check_background_spline_compatibility
    *status = CCL_B200_ERROR_INCONSISTENT;
    set_err("ccl_b200: tracer defined over wider redshift range than the background a-spline");
    return -1;
  }
  if (!cuda_ok(cudaSetDevice(t->device), "cudaSetDevice", status)) return -1;
  const size_t nc = (size_t)t->ncosmo, n = (size_t)n_samples;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  t->extras.emplace_back();
  DevBuf& buf = t->extras.back();
  const size_t o_chi = 0, o_w = up(nc * n * 8), o_rng = o_w + up(nc * n * 8);
  if (!buf.reserve(o_rng + up(nc * 4 * 8), status)) return -1;
  double* d_chi = reinterpret_cast<double*>(buf.p + o_chi);
  double* d_w = reinterpret_cast<double*>(buf.p + o_w);
  double* d_rng = reinterpret_cast<double*>(buf.p + o_rng);
  launch_build_kappa_kernel(t->ncosmo, t->na, t->d_a, t->d_chi, t->d_nachi, t->stride, t->d_gx, t->d_ga, t->d_par,
                            z_source, n_samples, d_chi, d_w, 0);
  launch_tracer_ranges(t->ncosmo, 1, n_samples, d_chi, 1, d_w, d_rng, 0);
  if (!cuda_ok(cudaGetLastError(), "kappa kernel builder launch", status)) return -1;
  std::vector<double> rng(nc * 4);
  if (!cuda_ok(cudaMemcpy(rng.data(), d_rng, nc * 4 * 8, cudaMemcpyDeviceToHost), "kappa kernel builder", status))
    return -1;
  int idx = -1;
  for (size_t i = 0; i < nc; ++i) {
    CosmoPlan& p = b->ts.cosmos[first + i];
    if (p.tracers.capacity() < 16) p.tracers.reserve(16);
    TracerPlan tp;
    tp.der_bessel = -1;   // pyccl/tracers.py:1009-1013
    tp.der_angles = 1;
    tp.chi_min = rng[i * 4 + 0];
    tp.chi_max = rng[i * 4 + 1];
    b->ts.plan_s1_dev(tp.kernel, 0, n_samples, rng[i * 4 + 2], rng[i * 4 + 3], 1, d_chi + i * n, d_w + i * n);
    p.tracers.push_back(tp);
    idx = (int)p.tracers.size() - 1;
  }
  b->ts.committed = false;
  return idx;
}

}  // extern "C"
