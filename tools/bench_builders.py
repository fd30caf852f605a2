#!/usr/bin/env python
"""Time the device table builders for a batch of w0-wa cosmologies (host in/out C-ABI entries)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccl_b200 as ccl  # noqa: E402
from ccl_b200 import device_builders as db, synthetic as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(0)
cos = [dict(ccl.Cosmology(Omega_c=rng.uniform(0.2, 0.35), Omega_b=rng.uniform(0.04, 0.055), h=rng.uniform(0.6, 0.8),
                          n_s=rng.uniform(0.92, 1.0), sigma8=rng.uniform(0.7, 0.9), w0=rng.uniform(-1.3, -0.7),
                          wa=rng.uniform(-0.5, 0.5), transfer_function='eisenstein_hu')._params) for _ in range(n)]
z = np.linspace(0.0, 3.5, 500)
nzs = S.smail_bins(z, 0.26, 0.94, edges=np.arange(0.2, 1.2 + 1e-9, 0.2), sigma_z=0.03) + \
    S.smail_bins(z, 0.13, 0.78, nbins=10, sigma_z=0.05)
for rep in range(2):
    t0 = time.perf_counter(); bgd = db.build_background(cos)
    t1 = time.perf_counter(); lin = db.build_linear_power(cos, "eisenstein_hu", bgd)
    t2 = time.perf_counter(); nl = db.build_halofit(cos, lin)
    t3 = time.perf_counter(); trk = db.build_tracer_kernels(cos, bgd, z, nzs)
    t4 = time.perf_counter()
    print("n=%d background %.1f ms, linear power %.1f ms, halofit %.1f ms, tracer kernels %.1f ms (incl. H2D/D2H of the "
          "host in/out entries)" % (n, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
