"""Synthetic flat tables with the shapes of the reference's defaults, built
with plain numpy from closed-form ingredients.  They are *inputs* (what pyccl's constructors
would hand to the C layer), used by the low-level tests, smoke() and the bench to exercise the
C ABI without going through the Cosmology/Tracer classes.  No physics accuracy is claimed.
"""
import numpy as np

CLIGHT_HMPC = 2997.92458


def linlog_spacing(xminlog, xmin, xmax, nlog, nlin):
    """ccl_linlog_spacing (src/ccl_utils.c:47-103)."""
    x = np.empty(nlin + nlog - 1)
    dx = (xmax - xmin) / (nlin - 1.)
    log_xchange = np.log(xmin)
    log_xmin = np.log(xminlog)
    dlog_x = (log_xchange - log_xmin) / (nlog - 1.)
    i = np.arange(nlin + nlog - 1)
    x[:nlog] = np.exp(log_xmin + dlog_x * i[:nlog])
    x[nlog:] = xmin + dx * (i[nlog:] - nlog + 1)
    x[0] = xminlog
    x[nlog - 1] = xmin
    x[-1] = xmax
    return x


def log_spacing(xmin, xmax, n):
    """ccl_log_spacing (src/ccl_utils.c:111-148): multiplicative recurrence."""
    dlog = (np.log(xmax) - np.log(xmin)) / (n - 1.)
    ratio = np.exp(dlog)
    x = np.empty(n)
    x[0] = xmin
    for i in range(1, n - 1):
        x[i] = x[i - 1] * ratio
    x[-1] = xmax
    return x


def background(Om=0.3, h=0.7, w0=-1.0, wa=0.0, dchi=5.0):
    """a(chi) on a uniform chi grid from 0 to chi(a=1e-4) for a flat w0-wa model (numpy quadrature)."""
    a = np.concatenate([np.geomspace(1e-4, 0.1, 4000, endpoint=False), np.linspace(0.1, 1.0, 20000)])
    OL = 1 - Om - 8.5e-5
    E = np.sqrt((Om + OL * a ** (-3 * (w0 + wa)) * np.exp(3 * wa * (a - 1)) + 8.5e-5 / a) / a ** 3)
    f = CLIGHT_HMPC / h / (a * a * E)
    chi = np.zeros_like(a)
    chi[:-1] = np.cumsum((0.5 * (f[1:] + f[:-1]) * np.diff(a))[::-1])[::-1]
    chi_max = chi[0]
    n = int(chi_max / dchi)
    chi_grid = np.linspace(0.0, chi_max, n)
    a_grid = np.interp(chi_grid, chi[::-1], a[::-1])
    a_grid[0] = 1.0
    a_grid[-1] = 1e-4
    # growth (Carroll-Press-Turner-like fit is enough for a synthetic D(a))
    a_g = linlog_spacing(1e-4, 0.1, 1.0, 250, 250)
    Eg = np.sqrt((Om + OL * a_g ** (-3 * (w0 + wa)) * np.exp(3 * wa * (a_g - 1)) + 8.5e-5 / a_g) / a_g ** 3)
    Oma = Om / a_g ** 3 / Eg ** 2
    OLa = 1 - Oma
    g = 2.5 * Oma / (Oma ** (4. / 7) - OLa + (1 + Oma / 2) * (1 + OLa / 70))
    D = g * a_g
    D /= D[-1]
    return dict(chi=chi_grid, a=a_grid, a_growth=a_g, growth=D, chi_of_a=(a, chi), h=h, Om=Om)


def power(bg, ns=0.96, sigma_like=1.0, nk=1220, nonlinear=True):
    """log P(k,a) on the default grids: BBKS-shaped linear power x D^2 (+ a smooth boost)."""
    k = log_spacing(5e-5, 1e3, nk)
    lk = np.log(k)
    a_pk = linlog_spacing(0.01, 0.1, 1.0, 11, 40)
    q = k / (bg["Om"] * bg["h"] ** 2)
    T = np.log(1 + 2.34 * q) / (2.34 * q) * (1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
    plin = 2.2e6 * sigma_like * k ** ns * T ** 2
    D = np.interp(a_pk, bg["a_growth"], bg["growth"])
    logp = np.log(plin)[None, :] + 2 * np.log(D)[:, None]
    if nonlinear:
        d2 = np.exp(logp) * k[None, :] ** 3 / (2 * np.pi ** 2)
        logp = logp + np.log1p(0.6 * d2 ** 0.9 / (1 + 0.02 * d2 ** 0.55))
    return dict(lk=lk, a=a_pk, logp=logp)


def _chi_of_z(bg, z):
    a, chi = bg["chi_of_a"]
    return np.interp(1. / (1 + z), a, chi)


def _E_of_z(bg, z, w0=-1.0, wa=0.0):
    a = 1. / (1 + z)
    Om = bg["Om"]
    OL = 1 - Om - 8.5e-5
    return np.sqrt((Om + OL * a ** (-3 * (w0 + wa)) * np.exp(3 * wa * (a - 1)) + 8.5e-5 / a) / a ** 3)


def smail_bins(z, z0, alpha, edges=None, nbins=None, sigma_z=0.05):
    """Smail parent n(z) cut into photo-z bins with error-function edges."""
    from math import erf
    parent = z ** 2 * np.exp(-(z / z0) ** alpha)
    if edges is None:
        cdf = np.cumsum(parent)
        cdf /= cdf[-1]
        edges = np.interp(np.linspace(0, 1, nbins + 1), cdf, z)
        edges[0], edges[-1] = 0.0, z[-1]
    verf = np.vectorize(erf)
    out = []
    for lo, hi in zip(edges[:-1], edges[1:]):
        s = sigma_z * (1 + z) * np.sqrt(2.)
        out.append(parent * 0.5 * (verf((hi - z) / s) - verf((lo - z) / s)))
    return out


def lsst_like_tracers(bg, n_lens=5, n_src=10, nz=500, n_samples=256, with_ia=False, with_rsd=False,
                      with_mag=False, w0=-1.0, wa=0.0):
    """Tracer tables of an LSST-Y1-like 3x2pt set: n_lens NumberCounts + n_src WeakLensing.

    Returns (subtracers, collections): `subtracers` is a list of dicts holding exactly the
    arguments of ccl_cl_tracer_t_new (src/ccl_tracers.c:569); `collections` is a list of
    (start, count) ranges, one per pyccl-level Tracer.
    """
    z = np.linspace(0.0, 3.5, nz)
    h, Om = bg["h"], bg["Om"]
    chi_z = _chi_of_z(bg, z)
    Hz = h * _E_of_z(bg, z, w0, wa) / CLIGHT_HMPC
    lens_nz = smail_bins(z, 0.26, 0.94, edges=np.arange(0.2, 0.2 * (n_lens + 1) + 1e-9, 0.2), sigma_z=0.03)
    src_nz = smail_bins(z, 0.13, 0.78, nbins=n_src, sigma_z=0.05)
    subs, colls = [], []

    def lensing_kernel(nzv):
        zk = z[-1] / n_samples * np.arange(n_samples) + 1e-15
        chik = _chi_of_z(bg, zk)
        norm = np.trapezoid(nzv, z)
        w = np.empty(n_samples)
        for i in range(n_samples):
            m = chi_z >= chik[i]
            integ = np.where(m, nzv * (chi_z - chik[i]) / np.where(chi_z > 0, chi_z, 1.0), 0.0)
            w[i] = np.trapezoid(integ, z)
        pref = 1.5 * (h / CLIGHT_HMPC) ** 2 * Om
        return chik, w / norm * pref * chik * (1 + zk)

    a_rev = 1. / (1 + z[::-1])
    for i, nzv in enumerate(lens_nz):
        start = len(subs)
        kern = Hz * nzv / np.trapezoid(nzv, z)
        bias = (1.0 + 0.2 * i) * np.ones(nz)
        subs.append(dict(der_bessel=0, der_angles=0, chi=chi_z, w=kern, a=a_rev, fa=bias[::-1]))
        if with_rsd:
            fz = (Om * (1 + z) ** 3 / _E_of_z(bg, z, w0, wa) ** 2) ** 0.55
            subs.append(dict(der_bessel=2, der_angles=0, chi=chi_z, w=kern, a=a_rev, fa=-fz[::-1]))
        if with_mag:
            ck, wk = lensing_kernel(nzv)
            subs.append(dict(der_bessel=-1, der_angles=1, chi=ck, w=-2 * wk * (1 - 2.5 * 0.4)))
        colls.append((start, len(subs) - start))
    for i, nzv in enumerate(src_nz):
        start = len(subs)
        ck, wk = lensing_kernel(nzv)
        subs.append(dict(der_bessel=-1, der_angles=2, chi=ck, w=wk))
        if with_ia:
            kern = Hz * nzv / np.trapezoid(nzv, z)
            D = np.interp(1. / (1 + z), bg["a_growth"], bg["growth"])
            aia = -0.5 * ((1 + z) / 1.62) ** 0.5 * 5e-14 * 2.775e11 * Om / D
            subs.append(dict(der_bessel=-1, der_angles=2, chi=chi_z, w=kern, a=a_rev, fa=aia[::-1]))
        colls.append((start, len(subs) - start))
    return subs, colls


def all_pairs(ncoll):
    return [(i, j) for i in range(ncoll) for j in range(i, ncoll)]
