"""Host-side (numpy) GSL-compatible 1-D splines used while *constructing* tables.

The reference builds its background / tracer tables in C with GSL Akima splines
(cosmo->spline_params.A_SPLINE_TYPE = gsl_interp_akima, src/ccl_core.c:249) and natural cubic
splines; the same interpolants are needed on the host to turn n(z), chi(a), D(a) ... into the
flat arrays handed to the C ABI.  Nothing here is on the Limber hot path (that runs in
libccl_b200.so); these are vectorised restatements for table set-up only.
"""
import numpy as np


class Akima:
    """gsl_interp_akima (non-periodic).  y may be [..., n]; splines along the last axis."""

    def __init__(self, x, y):
        x = np.asarray(x, dtype=float)
        y = np.asarray(y, dtype=float)
        n = x.size
        if n < 5:
            raise ValueError("Akima spline needs at least 5 points")
        if not np.all(np.diff(x) > 0):
            raise ValueError("spline nodes must be strictly increasing")
        self.x, self.y, self.n = x, y, n
        h = np.diff(x)
        m = np.empty(y.shape[:-1] + (n + 3,))           # m[..., j+2] = slope j, j in [-2, n]
        m[..., 2:n + 1] = np.diff(y, axis=-1) / h
        m[..., 1] = 2.0 * m[..., 2] - m[..., 3]
        m[..., 0] = 3.0 * m[..., 2] - 2.0 * m[..., 3]
        m[..., n + 1] = 2.0 * m[..., n] - m[..., n - 1]
        m[..., n + 2] = 3.0 * m[..., n] - 2.0 * m[..., n - 1]
        i = np.arange(n - 1)
        mm2, mm1, m0, mp1, mp2 = (m[..., i], m[..., i + 1], m[..., i + 2], m[..., i + 3], m[..., i + 4])
        NE = np.abs(mp1 - m0) + np.abs(mm1 - mm2)
        NE_next = np.abs(mp2 - mp1) + np.abs(m0 - mm1)
        with np.errstate(divide="ignore", invalid="ignore"):
            alpha_i = np.abs(mm1 - mm2) / NE
            alpha_ip1 = np.abs(m0 - mm1) / NE_next
            tL = np.where(NE_next == 0.0, m0, (1.0 - alpha_ip1) * m0 + alpha_ip1 * mp1)
            b = (1.0 - alpha_i) * mm1 + alpha_i * m0
            c = (3.0 * m0 - 2.0 * b - tL) / h
            d = (b + tL - 2.0 * m0) / (h * h)
        tie = NE == 0.0
        self.b = np.where(tie, m0, b)
        self.c = np.where(tie, 0.0, c)
        self.d = np.where(tie, 0.0, d)

    def _index(self, v):
        return np.clip(np.searchsorted(self.x, v, side="right") - 1, 0, self.n - 2)

    def __call__(self, v):
        """gsl_spline_eval for a 1-D spline (NaN outside the node range, like GSL's EDOM)."""
        v = np.asarray(v, dtype=float)
        i = self._index(v)
        dx = v - self.x[i]
        out = self.y[..., i] + dx * (self.b[..., i] + dx * (self.c[..., i] + self.d[..., i] * dx))
        return np.where((v < self.x[0]) | (v > self.x[-1]), np.nan, out)

    def interval_integrals(self):
        """Integral of every whole interval [x_i, x_{i+1}] (akima_eval_integ, interior formula)."""
        h = np.diff(self.x)
        return h * (self.y[..., :-1] + h * (0.5 * self.b + h * (self.c / 3.0 + 0.25 * self.d * h)))

    def integrate_nodes(self, i0=0, i1=None):
        """Integral between nodes i0 and i1 (default: last node)."""
        seg = self.interval_integrals()
        i1 = self.n - 1 if i1 is None else i1
        return seg[..., i0:i1].sum(axis=-1)

    def integrate(self, a, b):
        """gsl_spline_eval_integ for scalar limits inside the node range."""
        if a == b:
            return 0.0 * self.y[..., 0]
        ia, ib = int(self._index(a)), int(self._index(b))
        tot = 0.0
        for i in range(ia, ib + 1):
            x1 = a if i == ia else self.x[i]
            x2 = b if i == ib else self.x[i + 1]
            r1, r2 = x1 - self.x[i], x2 - self.x[i]
            r12 = r1 + r2
            tot = tot + (x2 - x1) * (self.y[..., i] + 0.5 * self.b[..., i] * r12
                                     + (1. / 3.) * self.c[..., i] * (r1 * r1 + r2 * r2 + r1 * r2)
                                     + 0.25 * self.d[..., i] * r12 * (r1 * r1 + r2 * r2))
        return tot


class F1D:
    """ccl_f1d_t with constant extrapolation (src/ccl_f1d.c:118-191): the end nodes themselves
    return the extrapolation value."""

    def __init__(self, x, y, y0=0.0, yf=0.0):
        self.spl = Akima(x, y)
        self.y0, self.yf = y0, yf

    def __call__(self, v):
        v = np.asarray(v, dtype=float)
        x = self.spl.x
        inner = self.spl(np.clip(v, x[0], x[-1]))
        return np.where(v <= x[0], self.y0, np.where(v >= x[-1], self.yf, inner))


class CSpline:
    """gsl_interp_cspline (natural).  y may be [..., n]."""

    def __init__(self, x, y):
        from scipy.linalg import solve_banded
        x = np.asarray(x, dtype=float)
        y = np.asarray(y, dtype=float)
        n = x.size
        if n < 3:
            raise ValueError("cubic spline needs at least 3 points")
        self.x, self.y, self.n = x, y, n
        h = np.diff(x)
        dy = np.diff(y, axis=-1)
        c = np.zeros(y.shape)
        g = 3.0 * (dy[..., 1:] / h[1:] - dy[..., :-1] / h[:-1])
        ab = np.zeros((3, n - 2))
        ab[1] = 2.0 * (h[1:] + h[:-1])
        ab[0, 1:] = h[1:-1]
        ab[2, :-1] = h[1:-1]
        sol = solve_banded((1, 1), ab, np.moveaxis(g, -1, 0).reshape(n - 2, -1))
        c[..., 1:-1] = np.moveaxis(sol.reshape((n - 2,) + y.shape[:-1]), 0, -1)
        self.c = c
        self.b = dy / h - h * (c[..., 1:] + 2.0 * c[..., :-1]) / 3.0
        self.d = (c[..., 1:] - c[..., :-1]) / (3.0 * h)

    def _index(self, v):
        return np.clip(np.searchsorted(self.x, v, side="right") - 1, 0, self.n - 2)

    def __call__(self, v, nu=0):
        v = np.asarray(v, dtype=float)
        i = self._index(v)
        dx = v - self.x[i]
        b, c, d = self.b[..., i], self.c[..., i], self.d[..., i]
        if nu == 0:
            return self.y[..., i] + dx * (b + dx * (c + dx * d))
        if nu == 1:
            return b + dx * (2.0 * c + 3.0 * d * dx)
        return 2.0 * c + 6.0 * d * dx


def gauss_legendre(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return x, w
