"""Host-side P(k,a) table builders: BBKS / Eisenstein-Hu linear power, sigma8 normalisation and
Takahashi+12 halofit (src/ccl_power.c:28-232, src/ccl_bbks.c, src/ccl_eh.c, src/ccl_halofit.c).

Set-up code of SURVEY.md 3.3 / 8(f)-2: it produces the log P(k,a) grid (50 x 1220 by default)
that the device path interpolates.  Quadratures use fixed Gauss-Legendre panels that are tighter
than the reference's CQUAD epsrel (1e-5); root finding (halofit's r_sigma) is done to 1e-12.
"""
import numpy as np

from ._hostspline import CSpline
from .background import omega_x
from .synthetic import linlog_spacing, log_spacing

_GLX, _GLW = np.polynomial.legendre.leggauss(16)
# quadrature panels per 25 e-folds of k and root tolerance of the halofit moments: the constants of
# csrc/builders.cu (HF_PANELS, HF_ROOT_TOL), see the convergence table there
HF_PANELS, HF_ROOT_TOL = 100, 1e-10


def _gl_nodes(lo, hi, npanel):
    """Composite 16-point Gauss-Legendre nodes/weights on [lo, hi]."""
    edges = np.linspace(lo, hi, npanel + 1)
    mid = 0.5 * (edges[1:] + edges[:-1])
    half = 0.5 * (edges[1:] - edges[:-1])
    x = (mid[:, None] + half[:, None] * _GLX).ravel()
    w = (half[:, None] * _GLW).ravel()
    return x, w


def bbks_power(p, k):
    """ccl_bbks_power (src/ccl_bbks.c:10-26)."""
    tfac = p["T_CMB"] / 2.7
    q = tfac * tfac * k / (p["Omega_m"] * p["h"] * p["h"] *
                           np.exp(-p["Omega_b"] * (1.0 + np.sqrt(2. * p["h"]) / p["Omega_m"])))
    t2 = ((np.log(1. + 2.34 * q) / (2.34 * q)) ** 2.0 /
          np.sqrt(1. + 3.89 * q + (16.1 * q) ** 2.0 + (5.46 * q) ** 3.0 + (6.71 * q) ** 4.0))
    return k ** p["n_s"] * t2


def eh_power(p, k, wiggled=True):
    """ccl_eh_struct_new + ccl_eh_power (src/ccl_eh.c:9-204); k in 1/Mpc."""
    h = p["h"]
    OMh2 = p["Omega_m"] * h * h
    OBh2 = p["Omega_b"] * h * h
    th2p7 = p["T_CMB"] / 2.7
    zeq = 2.5E4 * OMh2 / th2p7 ** 4
    keq = 0.0746 * OMh2 / (h * th2p7 * th2p7)
    b1 = 0.313 * OMh2 ** -0.419 * (1 + 0.607 * OMh2 ** 0.674)
    b2 = 0.238 * OMh2 ** 0.223
    zdrag = 1291 * OMh2 ** 0.251 * (1 + b1 * OBh2 ** b2) / (1 + 0.659 * OMh2 ** 0.828)
    Req = 31.5 * OBh2 * 1000. / (zeq * th2p7 ** 4)
    Rd = 31.5 * OBh2 * 1000. / ((1 + zdrag) * th2p7 ** 4)
    rsound = 2 / (3 * keq) * np.sqrt(6 / Req) * np.log((np.sqrt(1 + Rd) + np.sqrt(Rd + Req)) / (1 + np.sqrt(Req)))
    kSilk = 1.6 * OBh2 ** 0.52 * OMh2 ** 0.73 * (1 + (10.4 * OMh2) ** -0.95) / h
    a1 = (46.9 * OMh2) ** 0.670 * (1 + (32.1 * OMh2) ** -0.532)
    a2 = (12.0 * OMh2) ** 0.424 * (1 + (45.0 * OMh2) ** -0.582)
    b_frac = OBh2 / OMh2
    alphac = a1 ** -b_frac * a2 ** (-b_frac ** 3)
    bb1 = 0.944 / (1 + (458 * OMh2) ** -0.708)
    bb2 = (0.395 * OMh2) ** -0.0266
    betac = 1 / (1 + bb1 * ((1 - b_frac) ** bb2 - 1))
    y = zeq / (1 + zdrag)
    sqy = np.sqrt(1 + y)
    gy = y * (-6 * sqy + (2 + 3 * y) * np.log((sqy + 1) / (sqy - 1)))
    alphab = 2.07 * keq * rsound * (1 + Rd) ** -0.75 * gy
    betab = 0.5 + b_frac + (3 - 2 * b_frac) * np.sqrt((17.2 * OMh2) ** 2 + 1)
    bnode = 8.41 * OMh2 ** 0.435
    rsound_approx = h * 44.5 * np.log(9.83 / OMh2) / np.sqrt(1 + 10 * OBh2 ** 0.75)

    kh = np.asarray(k, dtype=float) / h  # h/Mpc

    def tk0(kk, a, b):
        q = kk / (13.41 * keq)
        c = 14.2 / a + 386. / (1 + 69.9 * q ** 1.08)
        l = np.log(np.e + 1.8 * b * q)
        return l / (l + c * q * q)

    if wiggled:
        f = 1 / (1 + (kh * rsound / 5.4) ** 4)
        tc = f * tk0(kh, 1, betac) + (1 - f) * tk0(kh, alphac, betac)
        x = kh * rsound
        x_bessel = x * (1 + bnode ** 3 / (x * x * x)) ** (-1. / 3.)
        part1 = tk0(kh, 1, 1) / (1 + (x / 5.2) ** 2)
        part2 = alphab / (1 + (betab / x) ** 3) * np.exp(-(kh / kSilk) ** 1.4)
        ax2 = x_bessel * x_bessel
        jl = np.where(ax2 < 1e-4, 1 - ax2 * (1 - ax2 / 20.) / 6., np.sin(x_bessel) / np.where(x_bessel == 0, 1, x_bessel))
        tb = jl * (part1 + part2)
        tk = b_frac * tb + (1 - b_frac) * tc
    else:
        alpha_gamma = 1 - 0.328 * np.log(431 * OMh2) * b_frac + 0.38 * np.log(22.3 * OMh2) * b_frac * b_frac
        gamma_eff = p["Omega_m"] * h * (alpha_gamma + (1 - alpha_gamma) / (1 + (0.43 * kh * rsound_approx) ** 4))
        q = kh * th2p7 * th2p7 / gamma_eff
        l0 = np.log(2 * np.e + 1.8 * q)
        c0 = 14.2 + 731 / (1 + 62.5 * q)
        tk = l0 / (l0 + c0 * q * q)
    return np.asarray(k, dtype=float) ** p["n_s"] * tk * tk


def pk_grids(sp):
    """k and a grids of every internally built P(k) (src/ccl_power.c:33-43, 57, 74-76)."""
    ndecades = np.log10(sp["K_MAX"]) - np.log10(sp["K_MIN"])
    nk = int(np.ceil(ndecades * sp["N_K"]))
    k = log_spacing(sp["K_MIN"], sp["K_MAX"], nk)
    a = linlog_spacing(sp["A_SPLINE_MINLOG_PK"], sp["A_SPLINE_MIN_PK"], sp["A_SPLINE_MAX"],
                       sp["A_SPLINE_NLOG_PK"], sp["A_SPLINE_NA_PK"])
    return k, a


def _w_tophat(kR):
    """w_tophat (src/ccl_power.c:390-404)."""
    kR2 = kR * kR
    small = 1. + kR2 * (-1.0 / 10.0 + kR2 * (1.0 / 280.0 + kR2 * (-1.0 / 15120.0 + kR2 * (
        1.0 / 1330560.0 + kR2 * (-1.0 / 172972800.0)))))
    kRs = np.where(kR < 0.1, 1.0, kR)
    big = 3. * (np.sin(kRs) - kRs * np.cos(kRs)) / (kRs * kRs * kRs)
    return np.where(kR < 0.1, small, big)


def sigmaR_from_table(lk, logp_row, R, sp):
    """ccl_sigmaR (src/ccl_power.c:420-454) on the a-row of a log P table: the integrand reads the
    table through its natural cubic spline in ln k (what the bicubic reduces to on a node row)."""
    spl = CSpline(lk, logp_row)
    x, w = _gl_nodes(np.log10(sp["K_MIN"]), np.log10(sp["K_MAX"]), 600)
    k = 10. ** x
    lkq = np.clip(x * np.log(10.), lk[0], lk[-1])
    pk = np.exp(spl(lkq))
    wth = _w_tophat(k * R)
    integ = np.sum(w * pk * k * k * k * wth * wth)
    return np.sqrt(integ * np.log(10.) / (2 * np.pi * np.pi))


def linear_power_table(p, sp, growth_of_a, model):
    """ccl_compute_linpower_analytic (src/ccl_power.c:28-146).  Returns (a, lk, logP[na, nk], sigma8)."""
    if p.get("sigma8") is None or np.isnan(p["sigma8"]):
        raise ValueError("sigma8 not set, required for analytic power spectra")
    k, a = pk_grids(sp)
    if model == "bbks":
        pk = bbks_power(p, k)
    elif model == "eisenstein_hu":
        pk = eh_power(p, k, True)
    elif model == "eisenstein_hu_nowiggles":
        pk = eh_power(p, k, False)
    else:
        raise ValueError("unknown analytic transfer function %s" % model)
    y = np.log(pk)
    lk = np.log(k)
    g2 = 2. * np.log(growth_of_a(a))
    y2d = y[None, :] + g2[:, None]
    sigma8 = sigmaR_from_table(lk, y2d[-1], 8. / p["h"], sp)
    y2d = y2d + 2 * (np.log(p["sigma8"]) - np.log(sigma8))
    return a, lk, y2d, sigma8


# ---------------------------------------------------------------------------------------------
# halofit (Takahashi et al. 2012 + Bird et al. 2012), src/ccl_halofit.c
# ---------------------------------------------------------------------------------------------
def _plin_row_eval(spl, lk_tab, lnk, order_hik=2):
    """ccl_f2d_t_eval of the linear P(k) at a table scale factor: natural cspline inside the
    table, polynomial extrapolation in ln k above it (src/ccl_f2d.c:288-342)."""
    inside = np.clip(lnk, lk_tab[0], lk_tab[-1])
    val = spl(inside)
    dlk = lnk - inside
    hi = dlk > 0
    if np.any(hi) and order_hik > 0:
        d1 = spl(lk_tab[-1], 1)
        val = val + np.where(hi, d1 * dlk, 0.0)
        if order_hik > 1:
            d2 = spl(lk_tab[-1], 2)
            val = val + np.where(hi, 0.5 * d2 * dlk * dlk, 0.0)
    lo = dlk < 0
    if np.any(lo):
        d1 = spl(lk_tab[0], 1)
        val = val + np.where(lo, d1 * dlk, 0.0)
    return np.exp(val)


def halofit_table(p, a_arr, lk_arr, logplin):
    """ccl_apply_halofit (src/ccl_power.c:171-232): log P_NL on the grid of the linear table."""
    from scipy.optimize import brentq
    from ._hostspline import Akima
    na = a_arr.size
    k = np.exp(lk_arr)
    rsigma = np.empty(na)
    sigma2 = np.empty(na)
    neff = np.empty(na)
    Cc = np.empty(na)
    lnkmin, lnkmax_t = lk_arr[0], lk_arr[-1]
    for i in range(na):
        spl = CSpline(lk_arr, logplin[i])

        def moments(R, which):
            lnkmax = max(lnkmax_t, np.log(30. / R))
            x, w = _gl_nodes(lnkmin, lnkmax, int(HF_PANELS * max(1.0, (lnkmax - lnkmin) / 25.0)))
            kk = np.exp(x)
            k2 = kk * kk
            base = _plin_row_eval(spl, lk_arr, x) * kk * k2 / 2.0 / np.pi / np.pi * np.exp(-k2 * R * R)
            if which == 0:
                return np.sum(w * base)
            if which == 1:
                return np.sum(w * base * (-k2 * 2.0 * R))
            return np.sum(w * base * (-2.0 * k2 + 4.0 * k2 * k2 * R * R))

        f = lambda lnr: moments(np.exp(lnr), 0) - 1.0  # noqa: E731
        lo, hi = np.log(1e-64), np.log(1e16)
        # bracket more tightly first: sigma2(R) is monotonically decreasing in R
        lo, hi = np.log(1e-6), np.log(1e4)
        flo = f(lo)
        while flo < 0 and lo > np.log(1e-64):   # the reference brackets with [1e-64, 1e16]
            lo = max(lo - np.log(1e4), np.log(1e-64))
            flo = f(lo)
        if flo * f(hi) > 0:
            raise RuntimeError("could not solve for non-linear scale for halofit at scale factor %f" % a_arr[i])
        lnr = brentq(f, lo, hi, xtol=HF_ROOT_TOL, rtol=HF_ROOT_TOL)
        R = np.exp(lnr)
        rsigma[i] = R
        sigma2[i] = moments(R, 0)
        d1 = moments(R, 1)
        neff[i] = -R / sigma2[i] * d1 - 3.0
        d2 = moments(R, 2)
        ds = (neff[i] + 3.0) / (-R / sigma2[i])
        Cc[i] = -1.0 * (d2 * R * R / sigma2[i] + ds * R / sigma2[i] - ds * ds * R * R / sigma2[i] / sigma2[i])
    # the reference evaluates Akima splines of these at the same a nodes -> node values
    omegaMz = omega_x(a_arr, p, "matter")
    omegaDEwz = omega_x(a_arr, p, "dark_energy")
    weffa = p["w0"]
    out = np.empty_like(logplin)
    for i in range(na):
        ksigma = 1.0 / rsigma[i]
        n1 = neff[i]
        n2, n3, n4 = n1 * n1, n1 ** 3, n1 ** 4
        C = Cc[i]
        delta2_norm = k * k * k / 2.0 / np.pi / np.pi
        fnu = 0.0
        an = 10.0 ** (1.5222 + 2.8553 * n1 + 2.3706 * n2 + 0.9903 * n3 + 0.2250 * n4 - 0.6038 * C
                      + 0.1749 * omegaDEwz[i] * (1.0 + weffa))
        bn = 10.0 ** (-0.5642 + 0.5864 * n1 + 0.5716 * n2 - 1.5474 * C + 0.2279 * omegaDEwz[i] * (1.0 + weffa))
        cn = 10.0 ** (0.3698 + 2.0404 * n1 + 0.8161 * n2 + 0.5869 * C)
        gamman = 0.1971 - 0.0843 * n1 + 0.8460 * C
        alphan = abs(6.0835 + 1.3373 * n1 - 0.1959 * n2 - 5.5274 * C)
        betan = 2.0379 - 0.7354 * n1 + 0.3157 * n2 + 1.2490 * n3 + 0.3980 * n4 - 0.1682 * C
        mun = 0.0
        nun = 10.0 ** (5.2105 + 3.6902 * n1)
        om = omegaMz[i]
        if abs(1.0 - om) > 0.01:
            f1a, f2a, f3a = om ** -0.0732, om ** -0.1423, om ** 0.0725
            f1b, f2b, f3b = om ** -0.0307, om ** -0.0585, om ** 0.0743
            fb = omegaDEwz[i] / (1.0 - om)
            f1 = fb * f1b + (1.0 - fb) * f1a
            f2 = fb * f2b + (1.0 - fb) * f2a
            f3 = fb * f3b + (1.0 - fb) * f3a
        else:
            f1 = f2 = f3 = 1.0
        betan += fnu * (1.081 + 0.395 * n2)
        PkL = np.exp(logplin[i])
        y = k / ksigma
        y2 = y * y
        fy = y / 4.0 + y2 / 8.0
        DeltakL = PkL * delta2_norm
        kh = k / p["h"]
        kh2 = kh * kh
        DeltakL_tilde = DeltakL * (1.0 + fnu * (47.48 * kh2) / (1.0 + 1.5 * kh2))
        DeltakQ = DeltakL * (1.0 + DeltakL_tilde) ** betan / (1.0 + alphan * DeltakL_tilde) * np.exp(-fy)
        DeltakHprime = an * y ** (3.0 * f1) / (1.0 + bn * y ** f2 + (cn * f3 * y) ** (3.0 - gamman))
        DeltakH = DeltakHprime / (1.0 + mun / y + nun / y2)
        Qnu = fnu * (0.977 - 18.015 * (p["Omega_m"] - 0.3))
        DeltakH = DeltakH * (1.0 + Qnu)
        out[i] = np.log((DeltakQ + DeltakH) / delta2_norm)
    return out, dict(rsigma=rsigma, sigma2=sigma2, n_eff=neff, C=Cc)
