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
print("setup (cosmology tables + tracers, host) %.1f ms; first angular_cl call (upload + kernels) %.1f ms" %
      ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
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
