#!/usr/bin/env python
"""Latency of single pyccl-style angular_cl calls (BASELINE configs[0], the README example) through the
public API, next to the oracle on the host cores.  usage (GPU box): python tools/bench_api_call.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ccl_b200 as ccl  # noqa: E402

t0 = time.perf_counter()
cosmo = ccl.Cosmology(Omega_c=0.27, Omega_b=0.045, h=0.67, sigma8=0.8, n_s=0.96, transfer_function='bbks')
z = np.linspace(0., 1., 500)
n = np.ones_like(z)
lens1 = ccl.WeakLensingTracer(cosmo, dndz=(z, n))
lens2 = ccl.WeakLensingTracer(cosmo, dndz=(z, n))
ell = np.arange(2, 2001).astype(float)
t1 = time.perf_counter()
cl = ccl.angular_cl(cosmo, lens1, lens2, ell, limber_integration_method='spline')
t2 = time.perf_counter()
print("first setup (library load + CUDA context + cosmology tables + tracers) %.1f ms; first angular_cl call "
      "(upload + kernels) %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
t0 = time.perf_counter()
cosmo_b = ccl.Cosmology(Omega_c=0.26, Omega_b=0.045, h=0.68, sigma8=0.81, n_s=0.96, transfer_function='bbks')
lens_b = ccl.WeakLensingTracer(cosmo_b, dndz=(z, n))
cl_b = ccl.angular_cl(cosmo_b, lens_b, lens_b, ell, limber_integration_method='spline')
print("a second cosmology, parameters -> tables (device builders) -> tracer -> %d C_ell: %.1f ms" %
      (ell.size, (time.perf_counter() - t0) * 1e3))
for method in ("spline", "qag_quad"):
    ts = []
    for _ in range(20):
        ta = time.perf_counter()
        cl = ccl.angular_cl(cosmo, lens1, lens2, ell, limber_integration_method=method)
        ts.append(time.perf_counter() - ta)
    print("%-9s repeat call, %d ells: median %.3f ms (min %.3f)  -> %.2e C_ell/s" %
          (method, ell.size, np.median(ts) * 1e3, np.min(ts) * 1e3, ell.size / np.median(ts)))
import oracle  # noqa: E402
from golden_setup import oracle_objects  # noqa: E402
oc, op, otr = oracle_objects(cosmo, {"a": lens1, "b": lens2})
oracle.set_num_threads(len(os.sched_getaffinity(0)))
for method, m in (("spline", oracle.INTEG_SPLINE), ("qag_quad", oracle.INTEG_QAG)):
    ts = []
    for _ in range(5):
        ta = time.perf_counter()
        ref = oracle.angular_cl(oc, otr["a"], otr["b"], op, ell, m)[0]
        ts.append(time.perf_counter() - ta)
    print("oracle %-9s on %d host threads: median %.2f ms" % (method, oracle.num_threads(), np.median(ts) * 1e3))

# ---- the consumers either side of angular_cl: C_ell -> xi(theta) and the FKEM non-Limber branch ----
theta = np.geomspace(0.01, 5.0, 30)
for method in ("fftlog", "legendre", "bessel"):
    ts, to = [], []
    for _ in range(6):
        ta = time.perf_counter()
        xi = ccl.correlation(cosmo, ell=ell, C_ell=cl, theta=theta, type='NN', method=method)
        ts.append(time.perf_counter() - ta)
    for _ in range(3):
        ta = time.perf_counter()
        ref, st = oracle.correlation(ell, cl, theta, 'NN', method)
        to.append(time.perf_counter() - ta)
    print("correlation %-8s %d angles: device median %.3f ms, oracle on %d host threads %.2f ms, max |dev - oracle| / max|xi| %.1e"
          % (method, theta.size, np.median(ts) * 1e3, oracle.num_threads(), np.median(to) * 1e3,
             np.max(np.abs(xi - ref)) / np.max(np.abs(ref))))
zz = np.linspace(0.01, 2.0, 200)
nc = ccl.NumberCountsTracer(cosmo, has_rsd=False, dndz=(zz, np.exp(-((zz - 1.0) ** 2) / 0.1)), bias=(zz, np.ones_like(zz)))
ls = np.unique(np.geomspace(2, 2000, 128).astype(int)).astype(float)
for kw, name in ((dict(l_limber=-1), "Limber only"), (dict(l_limber=200, fkem_Nchi=400, fkem_chi_min=1e-4), "FKEM to l=200"),
                 (dict(l_limber='auto', fkem_Nchi=400, fkem_chi_min=1e-4), "FKEM auto")):
    ts = []
    for _ in range(4):
        nc._fkem_fft = None  # no reuse of cached transforms between repetitions
        ta = time.perf_counter()
        c, meta = ccl.angular_cl(cosmo, nc, nc, ls, return_meta=True, **kw)
        ts.append(time.perf_counter() - ta)
    print("angular_cl %-14s %d ells: median %.2f ms (l_limber %s)" % (name, ls.size, np.median(ts) * 1e3, meta['l_limber']))
