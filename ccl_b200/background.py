"""Host-side background tables: E(a), chi(a), a(chi), D(a), f(a).

These are the table *builders* of src/ccl_background.c: they run once per
cosmology and produce the flat arrays the device path reads.  The quadratures are done with
fixed high-order Gauss-Legendre panels instead of GSL's adaptive CQUAD/RKCK, i.e. they are
more accurate than the reference's tolerances (CQUAD epsrel 1e-6, RKCK epsrel 1e-6); the
a(chi) Newton chain keeps the reference's stopping rule (ROOT_EPSREL) because that rule, not
the quadrature, decides the node values (src/ccl_background.c:370-410, 549-556).
Massive neutrinos are out of scope (configs use m_nu = 0; src/ccl_background.c:23-31).
"""
import numpy as np

from .parameters import physical_constants as const
from .synthetic import linlog_spacing

_GLX, _GLW = np.polynomial.legendre.leggauss(12)


def h_over_h0(a, p):
    """h_over_h0 (src/ccl_background.c:18-48) for m_nu = 0."""
    a = np.asarray(a, dtype=float)
    return np.sqrt((p["Omega_c"] + p["Omega_b"]
                    + p["Omega_l"] * a ** (-3 * (p["w0"] + p["wa"])) * np.exp(3 * p["wa"] * (a - 1))
                    + p["Omega_k"] * a + (p["Omega_g"] + p["Omega_nu_rel"]) / a) / (a * a * a))


def omega_x(a, p, label):
    """ccl_omega_x (src/ccl_background.c:63-103) for the species the path needs."""
    a = np.asarray(a, dtype=float)
    hn = h_over_h0(a, p)
    if label == "matter":
        return (p["Omega_c"] + p["Omega_b"]) / (a * a * a) / hn / hn
    if label == "dark_energy":
        return p["Omega_l"] * a ** (-3 * (1 + p["w0"] + p["wa"])) * np.exp(3 * p["wa"] * (a - 1)) / hn / hn
    if label == "radiation":
        return p["Omega_g"] / a ** 4 / hn / hn
    if label == "curvature":
        return p["Omega_k"] / (a * a) / hn / hn
    if label == "neutrinos_rel":
        return p["Omega_nu_rel"] / a ** 4 / hn / hn
    if label == "neutrinos_massive":
        return 0.0 * a
    if label == "critical":
        return 1.0 + 0.0 * a
    raise ValueError("unknown species %s" % label)


def _chi_integrand(a, p):
    # chi_integrand / h (src/ccl_background.c:141-147, 313-315)
    return const.CLIGHT_HMPC / (a * a * h_over_h0(a, p)) / p["h"]


def _panel(lo, hi, p):
    """Gauss-Legendre integral of the chi integrand over [lo, hi] (arrays allowed)."""
    lo = np.asarray(lo, dtype=float)
    hi = np.asarray(hi, dtype=float)
    mid, half = 0.5 * (hi + lo), 0.5 * (hi - lo)
    x = mid[..., None] + half[..., None] * _GLX
    return half * np.sum(_GLW * _chi_integrand(x, p), axis=-1)


class Background:
    """Tables of ccl_cosmology_compute_distances / _compute_growth for one cosmology."""

    def __init__(self, p, spline_params, gsl_params):
        self.p = p
        self.sp = spline_params
        self.gp = gsl_params
        self.a = linlog_spacing(self.sp["A_SPLINE_MINLOG"], self.sp["A_SPLINE_MIN"], self.sp["A_SPLINE_MAX"],
                                self.sp["A_SPLINE_NLOG"], self.sp["A_SPLINE_NA"])
        self.E = None
        self.chi = None
        self.achi_chi = None
        self.achi_a = None
        self.growth = None
        self.fgrowth = None
        self.growth0 = None

    # ---- distances (src/ccl_background.c:417-589) ----
    def compute_distances(self):
        a = self.a
        self.E = h_over_h0(a, self.p)
        seg = _panel(a[:-1], a[1:], self.p)
        chi = np.zeros_like(a)
        chi[:-1] = np.cumsum(seg[::-1])[::-1]
        self.chi = chi
        self._fine_a, self._fine_chi = a, chi
        # a(chi) on a uniform chi grid, Newton chain seeded by the previous root
        chi0, chif = chi[-1], chi[0]
        a0, af = a[-1], a[0]
        na = int((chif - chi0) / 5.0)
        from .synthetic import CLIGHT_HMPC  # noqa: F401
        dx = (chif - chi0) / (na - 1.)
        chi_grid = chi0 + dx * np.arange(na)
        chi_grid[0], chi_grid[-1] = chi0, chif
        a_grid = np.empty(na)
        a_grid[0], a_grid[-1] = a0, af
        eps = self.gp["ROOT_EPSREL"]
        nmax = self.gp["ROOT_N_ITERATION"]
        cur = a0
        for i in range(1, na - 1):
            target = chi_grid[i]
            f = target - self._chi_scalar(cur)
            df = float(_chi_integrand(cur, self.p))
            it = 0
            while True:
                it += 1
                prev = cur
                cur = cur - f / df
                f = target - self._chi_scalar(cur)
                df = float(_chi_integrand(cur, self.p))
                if abs(cur - prev) < eps * abs(cur) or cur == prev or it > nmax:
                    break
            a_grid[i] = cur
        self.achi_chi, self.achi_a = chi_grid, a_grid

    def _chi_scalar(self, a):
        """chi(a) = chi(node above) + integral from a to that node (exact to quadrature precision)."""
        fa = self._fine_a
        if a >= 1.0:
            return float(-_panel(1.0, a, self.p)) if a > 1.0 else 0.0
        j = int(np.searchsorted(fa, a, side="right"))  # fa[j-1] <= a < fa[j]
        j = min(max(j, 1), fa.size - 1)
        return float(self._fine_chi[j] + _panel(a, fa[j], self.p))

    def chi_of_a_exact(self, a):
        return np.array([self._chi_scalar(float(x)) for x in np.atleast_1d(a)])

    # ---- growth (src/ccl_background.c:153-165, 209-284, 763-1002) ----
    def compute_growth(self):
        from scipy.integrate import solve_ivp
        p = self.p
        ai = self.gp["EPS_SCALEFAC_GROWTH"]

        def rhs(lna, y):
            a = np.exp(lna)
            hn = float(h_over_h0(a, p))
            om = (p["Omega_c"] + p["Omega_b"]) / (a * a * a) / hn / hn
            # d/dlna = a d/da of (src/ccl_background.c:153-165)
            return [a * y[1] / (a * a * a * hn), a * 1.5 * hn * a * om * y[0]]

        y0 = [ai, ai ** 3 * float(h_over_h0(ai, p))]
        an = self.a[self.a >= ai]
        nodes = an if an[-1] == 1.0 else np.concatenate([an, [1.0]])
        appended = nodes.size - an.size
        sol = solve_ivp(rhs, (np.log(ai), 0.0), y0, method="DOP853", t_eval=np.log(nodes), rtol=1e-11,
                        atol=1e-30)
        g = sol.y[0]
        yy = sol.y[1]
        growth0 = g[-1]
        D = np.empty_like(self.a)
        f = np.empty_like(self.a)
        lo = self.a < ai
        D[lo] = self.a[lo]
        f[lo] = 1.0
        ng = nodes.size - appended
        D[~lo] = g[:ng]
        f[~lo] = yy[:ng] / (an * an * h_over_h0(an, p) * g[:ng])
        self.growth = D / growth0
        self.fgrowth = f
        self.growth0 = growth0
