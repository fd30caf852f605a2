#!/usr/bin/env python
"""Extract the reference's golden vectors for the Limber path into tests/golden/.

Source (read-only, only available in the build container): /root/reference/benchmarks/data/
  * run_{b1b1,b1b2,b2b1,b2b2}{analytic,histo}_log_cl_{dd,dl,di,dc,ll,li,lc,ii}.txt, run_log_cl_cc.txt:
    C_ell of 23 tracer-pair types from an independent code, asserted by the reference at
    0.1 sigma (qag_quad) / 0.2 sigma (spline) of cosmic variance (benchmarks/test_cells.py:125-223)
  * bin{1,2}_histo.txt, ia_amp_{analytic,histo}_{1,2}.txt: the n(z) and IA amplitude inputs
Only the 541 multipoles the reference test reads are kept (benchmarks/test_cells.py:22-26).
"""
import os
import sys

import numpy as np

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/benchmarks/data/"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)

nls = 541
ells = np.zeros(nls)
ells[:50] = np.arange(50) + 2
ells[50:] = ells[49] + 6 * (np.arange(nls - 50) + 1)
out = {"ells": ells}
for nztyp in ("analytic", "histo"):
    for b in ("b1b1", "b1b2", "b2b1", "b2b2"):
        for kind in ("dd", "dl", "di", "dc", "ll", "li", "lc", "ii"):
            f = os.path.join(SRC, "run_%s%s_log_cl_%s.txt" % (b, nztyp, kind))
            if os.path.exists(f):
                out["%s_%s_%s" % (nztyp, b, kind)] = np.loadtxt(f, unpack=True)[1][ells.astype(int)]
    for i in (1, 2):
        out["ia_amp_%s_%d" % (nztyp, i)] = np.loadtxt(os.path.join(SRC, "ia_amp_%s_%d.txt" % (nztyp, i)))
for i in (1, 2):
    out["bin%d_histo" % i] = np.loadtxt(os.path.join(SRC, "bin%d_histo.txt" % i))
out["cc"] = np.loadtxt(os.path.join(SRC, "run_log_cl_cc.txt"), unpack=True)[1][ells.astype(int)]
np.savez_compressed(os.path.join(OUT, "cells_benchmark.npz"), **out)
print("wrote %d arrays to %s" % (len(out), os.path.join(OUT, "cells_benchmark.npz")))

# Upstream tables the path reads (pins of the host table builders, benchmarks/test_bbks.py 1e-5,
# test_distances.py 1e-4, test_growth.py 1e-4): independent-code values shipped with the reference.
up = {}
for m in (1, 2, 3):
    up["bbks_model%d_pk" % m] = np.loadtxt(os.path.join(SRC, "model%d_pk.txt" % m))
up["chi_model1_5"] = np.genfromtxt(os.path.join(SRC, "chi_model1-5.txt"))
up["growth_model1_5"] = np.genfromtxt(os.path.join(SRC, "growth_model1-5.txt"))
up["eh_model1_pk"] = np.loadtxt(os.path.join(SRC, "model1_pk_eh.txt"))
np.savez_compressed(os.path.join(OUT, "upstream_tables.npz"), **up)
print("wrote %d arrays to %s" % (len(up), os.path.join(OUT, "upstream_tables.npz")))


# ---- halofit pin: the reference's 3-D correlation benchmark of BBKS + halofit P(k) ----
# benchmarks/data/model{1,2,3}_xi.txt: xi(r) at z = 0..5 from an independent code, asserted by the
# reference at |r^2 dxi| < 3e-2 (benchmarks/test_correlation_3d.py:6-54, matter_power_spectrum='halofit').
xi = {}
for m in (1, 2, 3):
    f = os.path.join(SRC, "model%d_xi.txt" % m)
    if os.path.exists(f):
        xi["model%d" % m] = np.loadtxt(f)
if xi:
    np.savez_compressed(os.path.join(OUT, "halofit_xi.npz"), **xi)
    print("halofit_xi.npz:", {k: v.shape for k, v in xi.items()})


# ---- C_ell -> xi(theta) pin: the reference's angular correlation benchmark ----
# benchmarks/data/run_{b1b1,..}{analytic,histo}_log_wt_{dd,dl,di,ll_pp,ll_mm,li_pp,li_mm,ii_pp,ii_mm}.txt:
# xi(theta) of 28 tracer-pair cases at 15 angles from an independent code, asserted by the reference at 0.2
# (fftlog) / 0.1 (bessel) of the LSST error bars in sigma_{clustering,ggl,xi+,xi-}_Nbin5
# (benchmarks/test_correlation.py:12-231).
wt = {}
for nztyp in ("analytic", "histo"):
    for b in ("b1b1", "b1b2", "b2b1", "b2b2"):
        for kind in ("dd", "dl", "di", "ll_pp", "ll_mm", "li_pp", "li_mm", "ii_pp", "ii_mm"):
            f = os.path.join(SRC, "run_%s%s_log_wt_%s.txt" % (b, nztyp, kind))
            if os.path.exists(f):
                th, xi_ = np.loadtxt(f, unpack=True)
                wt["%s_%s_%s" % (nztyp, b, kind)] = xi_
                wt["theta"] = th
for name in ("clustering", "ggl", "xi+", "xi-"):
    f = os.path.join(SRC, "sigma_%s_Nbin5" % name)
    if os.path.exists(f):
        wt["sigma_" + name] = np.loadtxt(f, unpack=True)
if wt:
    np.savez_compressed(os.path.join(OUT, "correlation_benchmark.npz"), **wt)
    print("correlation_benchmark.npz: %d arrays" % len(wt))


# ---- FKEM pin: the reference's non-Limber (N5K) benchmark ----
# benchmarks/data/nonlimber/: truth C_ell of 10 clustering x 5 shear bins from brute-force non-Limber codes
# (benchmarks_nl_full_cl{gg,gs,ss}.npz), asserted by the reference at chi^2 < 0.3 per multipole against LSST-like
# error bars (benchmarks/test_nonlimber.py:180-204), with the N5K inputs they were computed from: n(z)
# (dNdzs_fullwidth.npz), P(k, z) linear and non-linear (pk.npz) and the background (background.npz).
nl_dir = os.path.join(SRC, "nonlimber")
if os.path.isdir(nl_dir):
    nl = {}
    for name, keys in (("background", ("z", "chi", "Ez")), ("pk", ("k", "z", "pk_nl", "pk_lin")),
                       ("dNdzs_fullwidth", ("z_sh", "dNdz_sh", "z_cl", "dNdz_cl"))):
        d = np.load(os.path.join(nl_dir, name + ".npz"))
        for key in keys:
            nl[name.split("_")[0] + "_" + key] = d[key]
    for kind in ("gg", "gs", "ss"):
        d = np.load(os.path.join(nl_dir, "benchmarks_nl_full_cl%s.npz" % kind))
        nl["cls_" + kind] = d["cls"]
        nl["ls"] = d["ls"]
    np.savez_compressed(os.path.join(OUT, "nonlimber_n5k.npz"), **nl)
    print("nonlimber_n5k.npz:", {k: v.shape for k, v in nl.items()})
