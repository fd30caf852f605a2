#!/usr/bin/env python
"""Where the host time of the stacked setters goes (per table kind), pinned vs pageable inputs.
usage (GPU box): python tools/stage_profile.py [ncosmo]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import ccl_b200 as cb  # noqa: E402
from ccl_b200 import _lib  # noqa: E402
from ccl_b200.lowlevel import darr, iarr  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
tables = bench.build_host_tables(2)
stk0 = bench.stack_host_tables(tables, n, 0)
L = _lib.load()
for label, stk in (("pageable", stk0), ("pinned", cb.pin_stacked(stk0))):
    b = cb.LibBatch(n)
    for rep in range(3):
        t = {}
        t0 = time.perf_counter()
        b.reset()
        t["reset"] = time.perf_counter() - t0
        st = C.c_int(0)
        n_per, chi, a = stk["achi"]
        kn, pn = iarr(n_per); kc, pc = darr(chi); ka, pa = darr(a)
        t0 = time.perf_counter()
        L.ccl_b200_batch_set_achi_stacked(b.h, 0, n, pn, kc.shape[1], pc, pa, C.byref(st))
        t["achi"] = time.perf_counter() - t0
        ag, D = stk["growth"]
        kg, pg = darr(ag); kd, pd_ = darr(D)
        t0 = time.perf_counter()
        L.ccl_b200_batch_set_growth_stacked(b.h, 0, n, kd.shape[1], 0 if kg.ndim == 1 else kg.shape[1], pg, pd_, C.byref(st))
        t["growth"] = time.perf_counter() - t0
        lk, apk, pk, olo, ohi, islog = stk["pk"]
        kl, pl = darr(lk); kap, pap = darr(apk); kp, pp = darr(pk)
        t0 = time.perf_counter()
        L.ccl_b200_batch_set_pk2d_stacked(b.h, 0, n, pl, len(kl), pap, len(kap), pp, 1, 2, 1, C.byref(st))
        t["pk"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        b2 = cb.LibBatch(n)
        t["(new batch)"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        b.reset()
        b.set_stacked(0, stk)
        t["all incl 15 tracers"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        b.commit()
        t["commit"] = time.perf_counter() - t0
        assert st.value == 0
    print(label, n, " ".join("%s %.2f ms;" % (k, v * 1e3) for k, v in t.items()))
