#!/usr/bin/env python
"""bench.py -- Limber C_ell throughput of the B200 path (driver contract).

Workload (BASELINE.json configs[3], the configuration the metric is quoted on):
  4096 DISTINCT w0-wa cosmologies (Latin hypercube, seed 0) x LSST-Y1-like 3x2pt
  (5 NumberCounts + 10 WeakLensing -> 120 pairs) x 100 log-spaced ells 2..3000, 'spline' Limber
  integration, fp64.  Every table (E, chi(a), a(chi), growth, Eisenstein-Hu P(k) + sigma8, halofit, the 15
  tracer kernels) is built by the device builders from the 7 parameters of each cosmology.  The batch is
  sharded by cosmology over the N ranks (strong scaling of the 4096-cosmology job) and the C_ell blocks
  are gathered to rank 0 with one NCCL gather inside the timed region.

One "step" = one pass of the hot path (ccl_angular_cls_limber for every cosmology, pair, ell)
over the resident batch.  `value` is C_ell/s with all tables already in HBM; `e2e` times the
host-buffer C-ABI path per step: the caller's stacked host tables (pinned) -> H2D copies, the device
table preparation (spline coefficients / bicubic planes), the Limber kernels and the D2H copy of every
C_ell; `from_parameters` times parameters -> device-built tables -> C_ell with no table on the host;
`qag` is the same resident pass with limber_integration_method='qag_quad' (pyccl's default) on a bounded
number of cosmologies.

`--impl reference` times the CPU restatement of the reference (oracle/, OpenMP over ell like
src/ccl_cls.c:282-356, pairs serial like benchmarks/test_timing.py:31-34) on a bounded sample
of the same workload; the GSL-based reference itself cannot be built in this image.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NCOSMO_TOTAL = 4096
N_DISTINCT = 8          # distinct synthetic cosmologies; slots cycle through them with a P(k) offset
NL = 100
REF_SAMPLE = 8          # cosmologies per step of the CPU reference arm (their host-built tables take ~5 s each)
QAG_SAMPLE = 512        # cosmologies of the 'qag_quad' leg of the default run
CPU_SAMPLE_S = 10.0     # seconds of CPU work of the cpu_baseline leg of the default run
# algorithmic flops per integrand node: base + per-sub-tracer + Akima integral
F_BASE, F_AKIMA = 81.0, 33.0
F_SUB = {"wl": 10.0, "nc": 26.0}


def build_host_tables(n_distinct, seed=0):
    """Flat host tables (what pyccl's constructors hand to the C layer) for n_distinct cosmologies."""
    from ccl_b200 import synthetic as S
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_distinct):
        Om = 0.25 + 0.1 * rng.random()
        h = 0.6 + 0.2 * rng.random()
        w0 = -1.3 + 0.6 * rng.random()
        wa = -0.5 + rng.random()
        bg = S.background(Om=Om, h=h, w0=w0, wa=wa)
        pk = S.power(bg, ns=0.92 + 0.08 * rng.random(), sigma_like=0.8 + 0.4 * rng.random())
        subs, colls = S.lsst_like_tracers(bg, w0=w0, wa=wa)
        a, chi = bg["chi_of_a"]
        out.append(dict(bg=bg, pk=pk, subs=subs, colls=colls, cmax=float(np.interp(0.1, a, chi))))
    return out


def fill_batch(batch, tables, nslots, offset0=0):
    """Stage nslots cosmologies (cycling through the distinct tables, perturbing log P)."""
    for ic in range(nslots):
        t = tables[(ic + offset0) % len(tables)]
        bg, pk = t["bg"], t["pk"]
        batch.set_cosmology(ic, bg["chi"], bg["a"], bg["a_growth"], bg["growth"], t["cmax"])
        batch.set_pk2d(ic, pk["lk"], pk["a"], pk["logp"] + 1e-3 * ((ic + offset0) // len(tables)))
        for s in t["subs"]:
            batch.add_tracer(ic, **s)


def stack_host_tables(tables, nslots, offset0=0):
    """The same per-slot tables as fill_batch, as stacked host arrays with a leading cosmology axis
    (what a likelihood code holds for a batch of cosmologies); consumed by LibBatch.set_stacked."""
    idx = [(ic + offset0) % len(tables) for ic in range(nslots)]
    shift = np.array([1e-3 * ((ic + offset0) // len(tables)) for ic in range(nslots)])
    nmax = max(t["bg"]["chi"].size for t in tables)
    n_per = np.array([tables[i]["bg"]["chi"].size for i in idx], dtype=np.int32)
    chi = np.zeros((nslots, nmax))
    a = np.zeros((nslots, nmax))
    for ic, i in enumerate(idx):
        bg = tables[i]["bg"]
        chi[ic, :n_per[ic]] = bg["chi"]
        a[ic, :n_per[ic]] = bg["a"]
    t0 = tables[0]
    stk = dict(count=nslots, achi=(n_per, chi, a),
               growth=(np.stack([tables[i]["bg"]["a_growth"] for i in idx]),
                       np.stack([tables[i]["bg"]["growth"] for i in idx])),
               chi_max_nokernel=np.array([tables[i]["cmax"] for i in idx]),
               pk=(t0["pk"]["lk"], t0["pk"]["a"],
                   np.stack([tables[i]["pk"]["logp"] for i in idx]) + shift[:, None, None], 1, 2, True),
               tracers=[])
    for it, s0 in enumerate(t0["subs"]):
        tr = dict(der_bessel=s0.get("der_bessel", 0), der_angles=s0.get("der_angles", 0),
                  chi=np.stack([tables[i]["subs"][it]["chi"] for i in idx]),
                  w=np.stack([tables[i]["subs"][it]["w"] for i in idx]))
        if s0.get("fa") is not None:
            tr["a"] = s0["a"]                                   # a = 1/(1+z) grid shared by all cosmologies
            tr["fa"] = np.stack([tables[i]["subs"][it]["fa"] for i in idx])
        stk["tracers"].append(tr)
    return stk


def algorithmic_flops(batch, colls, pairs, ells):
    """Sum over (cosmology, pair, ell) of nk * (F_base + sum F_tracer + F_akima), exact: the node
    counts come from ccl_b200_batch_spline_nodes (get_k_interval + the nk rule on every cosmology's own chi
    ranges).  Collections 0..4 are NumberCounts (constant bias counted as the a-only cspline it is in the
    reference, although the kernel folds it into the ell prefactor), 5..14 WeakLensing.
    Returns (nodes, flops, mean flop per node without the Akima term)."""
    kind = lambda c: "nc" if c < 5 else "wl"   # noqa: E731
    nodes_tot, flops_tot, fl_noak = 0, 0.0, 0.0
    for k1 in ("nc", "wl"):
        for k2 in ("nc", "wl"):
            sub = [p for p in pairs if (kind(p[0]), kind(p[1])) == (k1, k2)]
            if not sub:
                continue
            n = int(batch.spline_nodes(colls, sub, ells))
            nodes_tot += n
            flops_tot += n * (F_BASE + F_SUB[k1] + F_SUB[k2] + F_AKIMA)
            fl_noak += n * (F_BASE + F_SUB[k1] + F_SUB[k2])
    return nodes_tot, flops_tot, fl_noak / max(nodes_tot, 1)


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.02)
        except Exception as e:  # NVML unavailable: report that, never fake numbers
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def geometry():
    from ccl_b200 import synthetic as S
    colls = [(i, 1) for i in range(15)]
    pairs = S.all_pairs(15)
    ells = np.geomspace(2, 3000, NL)
    return colls, pairs, ells


def lhs_cosmologies(n, seed=0):
    """n derived-parameter dicts on a Latin hypercube (seed 0) over the SURVEY 8d config-4 ranges:
    Omega_c [0.2, 0.35], Omega_b [0.04, 0.055], h [0.6, 0.8], n_s [0.92, 1.0], sigma8 [0.7, 0.9],
    w0 [-1.3, -0.7], wa [-0.5, 0.5]."""
    from ccl_b200.cosmology import Cosmology
    rng = np.random.default_rng(seed)
    lo = np.array([0.2, 0.04, 0.6, 0.92, 0.7, -1.3, -0.5])
    hi = np.array([0.35, 0.055, 0.8, 1.0, 0.9, -0.7, 0.5])
    u = np.stack([(rng.permutation(n) + rng.random(n)) / n for _ in range(7)], axis=1)
    v = lo + u * (hi - lo)
    return [Cosmology._build_parameters(Omega_c=r[0], Omega_b=r[1], h=r[2], n_s=r[3], sigma8=r[4], A_s=None,
                                        Omega_k=0., Omega_g=None, Neff=3.044, m_nu=0., mass_split='normal',
                                        w0=r[5], wa=r[6], T_CMB=2.7255, T_ncdm=0.71611) for r in v]


LENS_BIAS = 1.0 + 0.2 * np.arange(5)


def survey_nz():
    """z grid and the 5 lens / 10 source n(z) of the LSST-Y1-like set."""
    from ccl_b200 import synthetic as S
    zz = np.linspace(0.0, 3.5, 500)
    lens = S.smail_bins(zz, 0.26, 0.94, edges=np.arange(0.2, 1.2 + 1e-9, 0.2), sigma_z=0.03)
    src = S.smail_bins(zz, 0.13, 0.78, nbins=10, sigma_z=0.05)
    return zz, lens, src


def stacked_row(stk, ic):
    """Cosmology ic of a stacked-table dict as the flat per-cosmology tables the oracle helpers take:
    (bg, pk, subs, chi_max_nokernel)."""
    n_per, chi, a = stk["achi"]
    n = int(n_per[ic])
    ag, D = stk["growth"]
    bg = dict(chi=np.ascontiguousarray(chi[ic, :n]), a=np.ascontiguousarray(a[ic, :n]),
              a_growth=np.asarray(ag if np.ndim(ag) == 1 else ag[ic]), growth=np.ascontiguousarray(D[ic]))
    lk, apk, pk, olo, ohi, islog = stk["pk"]
    subs = []
    for t in stk["tracers"]:
        s = dict(der_bessel=t.get("der_bessel", 0), der_angles=t.get("der_angles", 0),
                 chi=np.ascontiguousarray(t["chi"] if np.ndim(t["chi"]) == 1 else t["chi"][ic]),
                 w=np.ascontiguousarray(t["w"][ic]))
        if t.get("fa") is not None:
            s["a"] = np.asarray(t["a"] if np.ndim(t["a"]) == 1 else t["a"][ic])
            s["fa"] = np.ascontiguousarray(t["fa"][ic])
        subs.append(s)
    return bg, dict(lk=lk, a=apk, logp=pk[ic]), subs, float(stk["chi_max_nokernel"][ic])


def host_tables_from_parameters(p):
    """The same flat tables for one parameter set through the pyccl-compatible host classes (numpy twins of
    the device builders): the inputs of the CPU reference arm, which must not touch the GPU."""
    import ccl_b200 as ccl
    ccl.set_table_builder("host")      # numpy twins: this arm is pure CPU
    zz, lens, src = survey_nz()
    cosmo = ccl.Cosmology(Omega_c=p["Omega_c"], Omega_b=p["Omega_b"], h=p["h"], n_s=p["n_s"], sigma8=p["sigma8"],
                          w0=p["w0"], wa=p["wa"], transfer_function='eisenstein_hu', matter_power_spectrum='halofit')
    trc = [ccl.NumberCountsTracer(cosmo, has_rsd=False, dndz=(zz, n), bias=(zz, LENS_BIAS[i] * np.ones_like(zz)))
           for i, n in enumerate(lens)]
    trc += [ccl.WeakLensingTracer(cosmo, dndz=(zz, n)) for n in src]
    cosmo.compute_growth()
    b = cosmo._bg
    pk = cosmo.get_nonlin_power()
    bg = dict(chi=b.achi_chi, a=b.achi_a, a_growth=b.a, growth=b.growth)
    subs = [t._trc[0] for t in trc]
    return bg, dict(lk=pk._lk, a=pk._a, logp=pk._tab), subs, cosmo.chi_max_nokernel()


def workload_config(ncosmo, method):
    """The `config` object of the JSON line: identical in the B200 arm and the reference arm."""
    return {"workload": "LSST-Y1-like 3x2pt likelihood batch (BASELINE configs[3]): %d cosmologies x 120 pairs "
                        "(5 NumberCounts + 10 WeakLensing) x 100 log-spaced ells 2..3000" % ncosmo,
            "method": method, "cosmologies": ncosmo, "distinct_cosmologies": ncosmo,
            "parameters": "Latin hypercube (seed 0): Omega_c 0.2-0.35, Omega_b 0.04-0.055, h 0.6-0.8, n_s 0.92-1.0, "
                          "sigma8 0.7-0.9, w0 -1.3--0.7, wa -0.5-0.5 (SURVEY 8d config 4)",
            "tables": "per cosmology: E, chi(a), a(chi), growth, Eisenstein-Hu P(k) + sigma8, halofit, "
                      "5 number-counts + 10 lensing kernels",
            "sharding": "by cosmology, contiguous blocks, one gather",
            "l2": "inputs larger than L2 (3.4 MB of tables per cosmology, 14 GB per 4096 cosmologies)"}


def run_reference(args, rank, json_out=sys.stdout):
    """CPU restatement of the reference on a bounded sample: REF_SAMPLE cosmologies x 120 pairs x 100 ells.
    Pure host code: the sample's tables come from the numpy twins of the device builders."""
    if rank != 0:
        return
    import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import oracle_cosmo, oracle_pk, oracle_tracer
    colls, pairs, ells = geometry()
    nsample = min(REF_SAMPLE, args.ncosmo)
    plist = lhs_cosmologies(args.ncosmo)[:nsample]      # the first cosmologies of the same hypercube
    integ = oracle.INTEG_SPLINE if args.method == "spline" else oracle.INTEG_QAG
    objs = []
    for p in plist:
        bg, pk, subs, cmax = host_tables_from_parameters(p)
        objs.append((oracle_cosmo(bg), oracle_pk(pk), [oracle_tracer(s_, cmax) for s_ in subs]))
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1)
    oracle.set_num_threads(len(os.sched_getaffinity(0)))
    cores = oracle.num_threads()

    def one_pass():
        for oc, op, otr in objs:
            for (i, j) in pairs:
                oracle.angular_cl(oc, [otr[i]], [otr[j]], op, ells, integ)

    for _ in range(max(args.warmup, 1)):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_pass()
    dt = (time.perf_counter() - t0) / args.steps
    n_cl = nsample * len(pairs) * len(ells)
    val = n_cl / dt
    sample = ("%d of %d cosmologies x 120 pairs x 100 ells per step ('%s'), pairs serial, OpenMP over ell"
              % (nsample, args.ncosmo, args.method))
    print(json.dumps({
        "impl": "reference", "metric": "3x2pt C_ell/s (pair x ell), Limber '%s'" % args.method, "value": val,
        "unit": "C_ell/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.ncosmo, args.method),
        "run": {"reference_impl": "oracle/ CPU restatement of src/ccl_cls.c + GSL, gcc -O3 -fomit-frame-pointer like "
                                  "CMakeLists.txt:63 (the GSL-based reference is not buildable in this image: "
                                  "DESIGN.md section 3)", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "C_ell/s", "cores": cores, "kind": "port", "sample": sample,
                         "cosmologies_per_s": nsample / dt},
        "e2e": {"value": val, "unit": "C_ell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), file=json_out)


def main():
    # Libraries (NCCL's version banner, torchrun) may write to stdout: keep fd 1 for the ONE JSON line
    # of the contract and send everything else to stderr.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")
    try:
        _main(json_out)
    finally:
        json_out.flush()


def _main(json_out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ncosmo", type=int, default=NCOSMO_TOTAL, help="total cosmologies in the job")
    ap.add_argument("--method", default="spline", choices=["spline", "qag_quad"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-y10", action="store_true", help="skip the LSST-Y10 single-cosmology leg (configs[4])")
    ap.add_argument("--from-parameters", type=int, default=None, metavar="N",
                    help="cosmologies of the parameters -> device-built tables -> C_ell leg (default: this rank's "
                         "share of the batch; 0 skips it)")
    ap.add_argument("--qag-sample", type=int, default=QAG_SAMPLE, metavar="N",
                    help="cosmologies of the 'qag_quad' leg of a 'spline' run (0 skips it)")
    ap.add_argument("--e2e-trace", action="store_true", help="print the per-chunk stage timeline of one e2e step to stderr")
    ap.add_argument("--e2e-pageable", action="store_true",
                    help="keep the e2e input tables in pageable host memory (staged through pinned arenas by host threads)")
    ap.add_argument("--e2e-chunks", type=int, default=None, help="cosmology chunks of the pipelined host-buffer entry")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, json_out)
        return

    import torch
    import torch.distributed as dist
    import ccl_b200 as cb
    from ccl_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    if world > 1:
        # one core set per rank (like numactl in a launcher): the staging / planning threads of the ranks of one
        # box must not fight over the same cores
        cores = sorted(os.sched_getaffinity(0))
        lw = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        per = max(1, len(cores) // lw)
        os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per] or cores)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import ctypes as C
    st = C.c_int(0)
    _lib.load().ccl_b200_set_device(local_rank, C.byref(st))
    assert st.value == 0, _lib.last_error()

    method = _lib.INTEGRATION_SPLINE if args.method == "spline" else _lib.INTEGRATION_QAG_QUAD
    colls, pairs, ells = geometry()
    npair, nl = len(pairs), len(ells)
    # shard by cosmology: contiguous blocks, remainder to the first ranks
    base, rem = divmod(args.ncosmo, world)
    nloc = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)

    # ---- the stated workload: nloc DISTINCT Latin-hypercube cosmologies, every table device-built ----
    from ccl_b200 import device_builders as db
    from ccl_b200.pipeline import slice_stacked
    plist = lhs_cosmologies(args.ncosmo)
    mine = plist[off:off + nloc]
    zz, lens_nz, src_nz = survey_nz()
    t_b = time.perf_counter()
    stk, _ = db.build_3x2pt_stacked(mine, zz, lens_nz, src_nz, LENS_BIAS)   # device builders -> the caller's host tables
    build_s = time.perf_counter() - t_b
    if not args.e2e_pageable:
        stk = cb.pin_stacked(stk)       # page-locked, as the contract asks: the library DMAs straight from these
    batch = cb.LibBatch(nloc)
    batch.set_stacked(0, stk)
    batch.commit()
    nodes, flops, f_node_noakima = algorithmic_flops(batch, colls, pairs, ells)

    d_cl = torch.empty((nloc, npair, nl), dtype=torch.float64, device="cuda")
    d_st = torch.zeros((2, nloc, npair), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    from ccl_b200 import sharding
    sizes = sharding.shard_sizes(args.ncosmo, world)
    assert sizes[rank] == nloc and sharding.shard_range(args.ncosmo, world, rank)[0] == off

    def step():
        # this rank's block of cosmologies, then ONE gather of the C_ell blocks to rank 0 (product code)
        return sharding.angular_cl_by_cosmology(batch, colls, pairs, ells, sizes, method, out=d_cl, d_status=d_st, dst=0)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record()
    for i in range(args.steps):
        kev[i][0].record()
        batch.angular_cl_dev(colls, pairs, ells, d_cl.data_ptr(), d_st.data_ptr(), method, stream)
        kev[i][1].record()
        if world > 1:
            sharding.gather_blocks(d_cl, sizes, dst=0)
    ev1.record()
    sync_all()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    if world > 1:
        tmax = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    ms_per_step = elapsed_ms / args.steps
    n_cl_total = args.ncosmo * npair * nl
    value = n_cl_total / (ms_per_step * 1e-3)
    n_fail = int((d_st[0] != 0).sum().item())

    fast_path = args.method == "spline" and batch.last_used_fast_path
    # ---- fp64 roofline of the dominant kernel (this rank's launch) ----
    st = C.c_int(0)
    peak = _lib.load().ccl_b200_measure_fp64_peak(C.byref(st))
    if args.method == "qag_quad":
        # unit = one GK41 evaluation: the kernel counts them; same per-node costs, no Akima term
        nodes = int(batch.last_qag_evals)
        flops = nodes * f_node_noakima
    achieved = flops / (kern_ms * 1e-3) / 1e12
    traffic = executed = None
    tfile = os.path.join(ROOT, "profiles", "traffic.json")   # one ncu capture of this launch (tools/ncu_traffic.sh)
    if os.path.exists(tfile):
        tj = json.load(open(tfile))
        if tj.get("ncosmo") == nloc and tj.get("method") == args.method:
            traffic = tj.get("dram_bytes_per_launch")
            executed = tj.get("executed_fp64_flops_per_launch")
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak and peak > 0 else None, "traffic": traffic,
                "kernel": ("limber_spline_fast_kernel" if batch.last_used_fast_path else "limber_spline_kernel")
                if args.method == "spline" else "limber_qag_fast_kernel",
                "kernel_ms": kern_ms, "algorithmic_flops_per_launch": flops, "integrand_nodes_per_launch": nodes,
                "executed_fp64_flops_per_launch": executed,
                "flop_model": "SURVEY 8d per node: 81 (k, chi, a(chi), bicubic P, product) + per sub-tracer (WL 10, "
                              "NC 26 incl. its constant-bias a-only cspline, which the kernel folds into the ell "
                              "prefactor) + 33 Akima; executed < algorithmic because a(chi), P and every "
                              "collection's transfer are evaluated once per shared node grid",
                "peak_source": "measured in-run: DFMA microbenchmark ccl_b200_measure_fp64_peak "
                               "(MEASURED_PEAKS.json holds no fp64 figure)"}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        cl_host = None
        e2e_steps = max(1, min(args.steps, 3))

        parts = {"stage_host_tables_ms": 0.0, "h2d_and_table_prep_ms": 0.0, "limber_and_d2h_ms": 0.0}
        cl_pinned = torch.empty((nloc, npair, nl), dtype=torch.float64, pin_memory=True).numpy()

        def serial_step():
            """The three stages back to back on the resident batch: gives the per-stage breakdown."""
            t_a = time.perf_counter()
            batch.reset()
            batch.set_stacked(0, stk)
            t_b = time.perf_counter()
            batch.commit()
            t_c = time.perf_counter()
            out = batch.angular_cl(colls, pairs, ells, method, out=cl_pinned)
            t_d = time.perf_counter()
            parts["stage_host_tables_ms"] += (t_b - t_a) * 1e3
            parts["h2d_and_table_prep_ms"] += (t_c - t_b) * 1e3
            parts["limber_and_d2h_ms"] += (t_d - t_c) * 1e3
            return out

        serial_step()
        sync_all()
        for k_ in parts:
            parts[k_] = 0.0
        serial_step()
        # the public host-buffer entry: the same three stages pipelined over chunks of cosmologies
        pipe = cb.PipelinedBatch(nloc, nchunks=args.e2e_chunks)

        def e2e_step():
            return pipe.angular_cl(stk, colls, pairs, ells, method, out=cl_pinned)

        e2e_step()
        sync_all()
        if args.e2e_trace and rank == 0:
            pipe.trace = []
            t_ref = time.perf_counter()
            e2e_step()
            for name, c, ta, tb in sorted(pipe.trace, key=lambda r: r[2]):
                print("e2e-trace %-8s chunk %2d  %8.2f -> %8.2f ms (%6.2f)" % (name, c, (ta - t_ref) * 1e3, (tb - t_ref) * 1e3,
                                                                         (tb - ta) * 1e3), file=sys.stderr)
            pipe.trace = None
            sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            cl_host, stat_host, _ = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        h2d = int(pipe.h2d_bytes) + pipe.nchunks * (nl * 8 + (len(colls) + npair) * 8)
        d2h = int(cl_host.nbytes + stat_host.nbytes)
        e2e = {"value": n_cl_total / dt, "unit": "C_ell/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": dt * 1e3,
               "chunks": pipe.nchunks, "host_inputs": "pageable" if args.e2e_pageable else "pinned",
               "serial_breakdown_rank0": dict(parts),
               "identical_to_resident_path": bool(np.array_equal(cl_host, d_cl.cpu().numpy())),
               "note": "per step through the host-buffer entry (PipelinedBatch over the C ABI): stacked host tables "
                       "(pinned: one DMA per stacked array straight from the caller's buffers; pageable: staged "
                       "into pinned arenas by host threads) -> H2D copies, device table prep, Limber kernels, D2H "
                       "of all C_ell into pinned memory; the three stages overlap across chunks of cosmologies"}

    # ---- the whole likelihood step from cosmological parameters (device-built tables, nothing on the host) ----
    from_parameters = None
    nfp = nloc if args.from_parameters is None else min(nloc, args.from_parameters)
    if nfp and rank == 0:
        fp_list = mine[:nfp]
        fp_batch, fp_colls, fp_tabs = db.build_3x2pt_batch(fp_list, zz, lens_nz, src_nz, LENS_BIAS)   # warm-up
        d_cl2 = torch.empty((nfp, npair, nl), dtype=torch.float64, device="cuda")
        d_st2 = torch.zeros((2, nfp, npair), dtype=torch.int32, device="cuda")
        reps = []
        import gc
        gc.collect()
        gc.disable()         # a generation-2 collection over the harness's own objects (the 4096 Cosmology objects, the
                             # e2e leg's chunk lists) landing inside a pass would add tens of ms to it
        for _ in range(5):   # the median of five passes
            del fp_tabs
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fp_batch, fp_colls, fp_tabs = db.build_3x2pt_batch(fp_list, zz, lens_nz, src_nz, LENS_BIAS, batch=fp_batch)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            fp_batch.angular_cl_dev(fp_colls, pairs, ells, d_cl2.data_ptr(), d_st2.data_ptr(), method, stream)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            reps.append((t2 - t0, t0, t1, t2))
        gc.enable()
        _, t0, t1, t2 = sorted(reps)[2]
        same = float(torch.max(torch.abs(d_cl2 - d_cl[:nfp]) / torch.clamp(torch.abs(d_cl[:nfp]), min=1e-300)).item())
        from_parameters = {"cosmologies": nfp, "cosmologies_per_s": nfp / (t2 - t0), "C_ell_per_s": nfp * npair * nl / (t2 - t0),
                           "build_tables_ms": (t1 - t0) * 1e3, "limber_ms": (t2 - t1) * 1e3,
                           "failed_cl_entries": int((d_st2[0] != 0).sum().item()), "fast_path": bool(fp_batch.last_used_fast_path),
                           "max_rel_diff_vs_host_buffer_tables": same,
                           "passes_ms": [round(r[0] * 1e3, 2) for r in reps],
                           "passes_build_ms": [round((r[2] - r[1]) * 1e3, 2) for r in reps],
                           "note": "median of 5 passes; the batch's own 7-parameter cosmologies -> E, chi(a), a(chi), growth, Eisenstein-Hu P(k) + "
                                   "sigma8, halofit, 5 number-counts + 10 lensing kernels -> 120 x 100 C_ell, every table "
                                   "built and kept on the device"}
        del fp_batch, fp_tabs, d_cl2, d_st2

    # ---- 'qag_quad' (pyccl's default method) on the first cosmologies of the same batch ----
    qag = None
    nq = min(nloc, args.qag_sample)
    if nq and rank == 0 and world == 1 and args.method == "spline":
        qb = cb.LibBatch(nq)
        qb.set_stacked(0, slice_stacked(stk, 0, nq))
        qb.commit()
        d_clq = torch.empty((nq, npair, nl), dtype=torch.float64, device="cuda")
        d_stq = torch.zeros((2, nq, npair), dtype=torch.int32, device="cuda")
        qm = _lib.INTEGRATION_QAG_QUAD
        for _ in range(2):
            qb.angular_cl_dev(colls, pairs, ells, d_clq.data_ptr(), d_stq.data_ptr(), qm, stream)
        torch.cuda.synchronize()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nrep = 2
        q0.record()
        for _ in range(nrep):
            qb.angular_cl_dev(colls, pairs, ells, d_clq.data_ptr(), d_stq.data_ptr(), qm, stream)
        q1.record()
        torch.cuda.synchronize()
        q_ms = q0.elapsed_time(q1) / nrep
        q_evals = int(qb.last_qag_evals)
        q_comp = int(qb.last_qag_nodes_computed)
        q_flops = q_evals * f_node_noakima
        q_ach = q_flops / (q_ms * 1e-3) / 1e12
        qag = {"value": nq * npair * nl / (q_ms * 1e-3), "unit": "C_ell/s", "cosmologies": nq, "ms_per_pass": q_ms,
               "integrand_evaluations": q_evals, "evaluations_per_cl": q_evals / (nq * npair * nl),
               "abscissae_computed": q_comp,
               "achieved": q_ach, "frac": q_ach / peak if peak and peak > 0 else None,
               "failed_cl_entries": int((d_stq[0] != 0).sum().item()),
               "max_rel_err": None}
        q_host = d_clq[:2].cpu().numpy()
        del qb, d_stq

    # ---- BASELINE configs[4]: LSST-Y10 shear + IA, ONE cosmology, 210 pairs x 1000 ells, sharded by ell tile ----
    y10 = None
    if not args.no_y10:
        from ccl_b200 import synthetic as S
        y_stk = None
        y_bg = S.background()
        y_subs, y_colls = S.lsst_like_tracers(y_bg, n_lens=0, n_src=20, with_ia=True)
        y_pairs = S.all_pairs(len(y_colls))
        y_ells = np.geomspace(2, 5000, 1000)
        if rank == 0:      # the table pack lives on rank 0 and is broadcast once
            a_, chi_ = y_bg["chi_of_a"]
            y_tab = dict(bg=y_bg, pk=S.power(y_bg), subs=y_subs, colls=y_colls, cmax=float(np.interp(0.1, a_, chi_)))
            y_stk = stack_host_tables([y_tab], 1)
        tsl = sharding.TileShardedLimber(y_stk, y_colls, y_pairs, y_ells, src=0)
        for _ in range(3):
            y_out = tsl.angular_cl(method)
        sync_all()
        y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nrep = 10
        y0.record()
        for _ in range(nrep):
            y_out = tsl.angular_cl(method)
        y1.record()
        sync_all()
        y_ms = y0.elapsed_time(y1) / nrep
        if world > 1:
            tt = torch.tensor([y_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            y_ms = float(tt.item())
        if rank == 0:
            loads = nodes_loads = None
            w_ell = sharding.nodes_per_ell(tsl.batch, y_colls, y_pairs, y_ells)
            loads = [float(w_ell[p].sum()) for p in tsl.plan]
            same = None
            if world > 1:   # the split reproduces the unsplit single-GPU result bit for bit
                full = torch.empty((1, len(y_pairs), len(y_ells)), dtype=torch.float64, device="cuda")
                stf = torch.zeros((2, 1, len(y_pairs)), dtype=torch.int32, device="cuda")
                tsl.batch.angular_cl_dev(y_colls, y_pairs, y_ells, full.data_ptr(), stf.data_ptr(), method, stream)
                torch.cuda.synchronize()
                same = bool(torch.equal(full, y_out))
            y10 = {"value": len(y_pairs) * len(y_ells) / (y_ms * 1e-3), "unit": "C_ell/s", "ms_per_pass": y_ms,
                   "n_gpus": world, "pairs": len(y_pairs), "ells": len(y_ells), "finite": bool(torch.isfinite(y_out).all().item()),
                   "nodes_per_rank": loads, "identical_to_unsplit": same,
                   "note": "one cosmology, 20 IA-augmented source bins; table pack broadcast once, ell tiles (all pairs x the "
                           "ells of the rank) balanced on integrand nodes, one launch per rank, one gather (ccl_b200.sharding."
                           "TileShardedLimber)"}
        del tsl

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from common import oracle_cosmo, oracle_pk, oracle_tracer
        oracle.set_num_threads(len(os.sched_getaffinity(0)))
        # bounded sample: cosmologies of the batch, in slot order, until ~CPU_SAMPLE_S seconds of CPU work; every one
        # of them is also a parity check of the GPU result (same tables on both sides)
        integ = oracle.INTEG_SPLINE if args.method == "spline" else oracle.INTEG_QAG
        ndone, err, t_cpu = 0, 0.0, 0.0
        err_q = 0.0
        while ndone < nloc and t_cpu < CPU_SAMPLE_S:
            bg_i, pk_i, subs_i, cmax_i = stacked_row(stk, ndone)
            oc, op = oracle_cosmo(bg_i), oracle_pk(pk_i)
            otr = [oracle_tracer(s_, cmax_i) for s_ in subs_i]
            t0 = time.perf_counter()
            ref = np.empty((npair, nl))
            for q, (i, j) in enumerate(pairs):
                ref[q], _ = oracle.angular_cl(oc, [otr[i]], [otr[j]], op, ells, integ)
            t_cpu += time.perf_counter() - t0
            got = d_cl[ndone].cpu().numpy()
            scale = np.maximum(np.abs(ref), 1e-6 * np.max(np.abs(ref), axis=1, keepdims=True))
            err = max(err, float(np.max(np.abs(got - ref) / np.maximum(scale, 1e-300))))
            if qag is not None and ndone < 2:      # 'qag_quad' leg: GPU vs oracle on the first two cosmologies
                for q, (i, j) in enumerate(pairs):
                    rq, _ = oracle.angular_cl(oc, [otr[i]], [otr[j]], op, ells, oracle.INTEG_QAG)
                    sc = np.maximum(np.abs(rq), 1e-6 * np.max(np.abs(rq)))
                    err_q = max(err_q, float(np.max(np.abs(q_host[ndone, q] - rq) / np.maximum(sc, 1e-300))))
            ndone += 1
        if qag is not None:
            qag["max_rel_err"] = err_q
        cpu_baseline = {"value": ndone * npair * nl / t_cpu, "unit": "C_ell/s", "cores": oracle.num_threads(),
                        "kind": "port", "cosmologies_per_s": ndone / t_cpu,
                        "sample": "%d of %d cosmologies x 120 pairs x 100 ells (%.1f s of CPU work), '%s', pairs "
                                  "serial, OpenMP over ell (oracle restatement of src/ccl_cls.c)"
                                  % (ndone, args.ncosmo, t_cpu, args.method),
                        "max_rel_err_gpu_vs_cpu_on_sample": err}

    if rank == 0:
        print(json.dumps({
            "metric": "3x2pt C_ell/s (pair x ell), Limber '%s'" % args.method, "value": value, "unit": "C_ell/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args.ncosmo, args.method),
            "run": {"cosmologies_per_s": value / (npair * nl), "failed_cl_entries": n_fail,
                    "tables_gb_per_rank": batch.arena_bytes / 1e9, "gather": "one NCCL gather to rank 0" if world > 1 else None,
                    "device_table_build_s_rank0": build_s,
                    "fast_path": bool(fast_path),
                    "qag_abscissae_computed": int(batch.last_qag_nodes_computed) if args.method == "qag_quad" else None},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "from_parameters": from_parameters,
            "qag": qag, "y10": y10,
            # fast 'spline' path: grouping + integrator + redo kernels per step; otherwise one kernel
            "gpu_launches": args.steps * (3 if fast_path else 1), "clocks": clocks,
        }), file=json_out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
