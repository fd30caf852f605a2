"""Device table builders against the host builders of this package (which are pinned against the
reference's benchmark files in tests/test_upstream_tables_cpu.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cosmologies():
    """Cosmologies whose own tables come from the numpy twins: the host side of every comparison below."""
    import ccl_b200 as ccl
    prev = ccl.get_table_builder()
    ccl.set_table_builder("host")
    try:
        return _make_cosmologies(ccl)
    finally:
        ccl.set_table_builder(prev)


def _make_cosmologies(ccl):
    rng = np.random.default_rng(11)
    out = [ccl.Cosmology(Omega_c=0.27, Omega_b=0.045, h=0.67, sigma8=0.8, n_s=0.96, transfer_function='bbks')]
    for _ in range(5):
        out.append(ccl.Cosmology(Omega_c=rng.uniform(0.2, 0.35), Omega_b=rng.uniform(0.04, 0.055),
                                 h=rng.uniform(0.6, 0.8), n_s=rng.uniform(0.92, 1.0), sigma8=rng.uniform(0.7, 0.9),
                                 w0=rng.uniform(-1.3, -0.7), wa=rng.uniform(-0.5, 0.5),
                                 transfer_function='eisenstein_hu'))
    out.append(ccl.Cosmology(Omega_c=0.25, Omega_b=0.05, h=0.7, sigma8=0.8, n_s=0.96, Omega_k=0.05, Neff=0,
                             Omega_g=0, transfer_function='bbks'))
    return out


def test_background_matches_host_builder():
    from ccl_b200 import device_builders as db
    cos = _cosmologies()
    dev = db.build_background(cos)
    for i, c in enumerate(cos):
        c.compute_distances()
        c.compute_growth()
        bg = c._bg
        assert np.array_equal(dev["a"], bg.a)
        assert np.max(np.abs(dev["E"][i] / bg.E - 1)) < 1e-14
        assert np.max(np.abs(dev["chi"][i][:-1] / bg.chi[:-1] - 1)) < 1e-13 and dev["chi"][i][-1] == 0
        n = dev["n_achi"][i]
        assert n == bg.achi_chi.size
        assert np.max(np.abs(dev["achi_chi"][i, :n] - bg.achi_chi)) < 1e-9
        # same Newton chain and stopping rule: the node values agree far below the rule's own 1e-4
        assert np.max(np.abs(dev["achi_a"][i, :n] / bg.achi_a - 1)) < 1e-11
        # RK4 at dlna = 2e-3 vs DOP853 at rtol 1e-11 (the reference's RKCK runs at 1e-6)
        assert np.max(np.abs(dev["growth"][i] / bg.growth - 1)) < 2e-9
        assert np.max(np.abs(dev["fgrowth"][i] / bg.fgrowth - 1)) < 2e-9
        assert abs(dev["growth0"][i] / bg.growth0 - 1) < 2e-9


@pytest.mark.parametrize("model", ["bbks", "eisenstein_hu", "eisenstein_hu_nowiggles"])
def test_linear_power_matches_host_builder(model):
    import ccl_b200 as ccl
    from ccl_b200 import device_builders as db
    cos = _cosmologies()[:5]
    bgd = db.build_background(cos)
    dev = db.build_linear_power(cos, model, bgd)
    for i, c in enumerate(cos):
        ref = ccl.Pk2D.from_model(c, model)
        a, lk, tab = ref.get_spline_arrays_raw()
        assert np.array_equal(a, dev["a"]) and np.max(np.abs(lk - dev["lk"])) < 1e-14
        # log P agrees to the growth tables' agreement (2e-9 relative in D -> 4e-9 absolute in log P)
        assert np.max(np.abs(dev["logp"][i] - tab)) < 2e-8, (model, i, np.max(np.abs(dev["logp"][i] - tab)))


def test_tracer_kernels_match_host_builder():
    import ccl_b200 as ccl
    from ccl_b200 import device_builders as db
    from ccl_b200 import synthetic as S
    cos = _cosmologies()
    z = np.linspace(0.0, 3.5, 500)
    nzs = S.smail_bins(z, 0.26, 0.94, edges=np.arange(0.2, 1.2 + 1e-9, 0.2), sigma_z=0.03)[:2] + \
        S.smail_bins(z, 0.13, 0.78, nbins=10, sigma_z=0.05)[::4]
    sz = [0.4 + 0.1 * z for _ in nzs]
    bgd = db.build_background(cos)
    for mag in (None, sz):
        dev = db.build_tracer_kernels(cos, bgd, z, nzs, mag)
        for i, c in enumerate(cos):
            for t, n in enumerate(nzs):
                chi, w = ccl.get_density_kernel(c, dndz=(z, n))
                assert np.max(np.abs(dev["chi_z"][i] - chi)) < 1e-7 * chi.max()
                assert np.max(np.abs(dev["w_nc"][i, t] - w)) < 1e-12 * np.max(np.abs(w))
                ck, wk = ccl.get_lensing_kernel(c, dndz=(z, n), mag_bias=None if mag is None else (z, mag[t]),
                                                n_chi=256)
                assert np.max(np.abs(dev["chi_l"][i] - ck)) < 1e-7 * ck.max()
                # same Akima integral and edge term on tables that agree to ~1e-11
                assert np.max(np.abs(dev["w_l"][i, t] - wk)) < 1e-8 * np.max(np.abs(wk)), (i, t)


def test_halofit_matches_host_builder():
    import ccl_b200 as ccl
    from ccl_b200 import device_builders as db
    from ccl_b200.power import halofit_table
    cos = _cosmologies()[:4]
    bgd = db.build_background(cos)
    lin = db.build_linear_power(cos, "eisenstein_hu", bgd)
    dev = db.build_halofit(cos, lin)
    for i, c in enumerate(cos):
        ref, diag = halofit_table(c._params, lin["a"], lin["lk"], lin["logp"][i])
        assert np.max(np.abs(dev["rsigma"][i] / diag["rsigma"] - 1)) < 1e-9
        # both integrate the same spline with the same nodes; the roots agree to the solvers' 1e-13
        assert np.max(np.abs(dev["logp"][i] - ref)) < 1e-8, np.max(np.abs(dev["logp"][i] - ref))


def test_3x2pt_from_parameters_matches_api_path():
    """C_ell of a 3x2pt set computed from device-built tables == C_ell through the pyccl-compatible
    classes (host-built tables), for several w0-wa cosmologies in one batch."""
    import ccl_b200 as ccl
    from ccl_b200 import device_builders as db
    from ccl_b200 import synthetic as S
    cos = _cosmologies()[1:5]
    z = np.linspace(0.0, 3.5, 500)
    lens = S.smail_bins(z, 0.26, 0.94, edges=np.arange(0.2, 0.8 + 1e-9, 0.2), sigma_z=0.03)
    src = S.smail_bins(z, 0.13, 0.78, nbins=4, sigma_z=0.05)
    bias = [1.2, 1.4, 1.6]
    stk, colls = db.build_3x2pt_stacked(cos, z, lens, src, bias, model="eisenstein_hu", halofit=True)
    batch = ccl.LibBatch(len(cos))
    batch.set_stacked(0, stk)
    batch.commit()
    pairs = [(i, j) for i in range(len(colls)) for j in range(i, len(colls))]
    ells = np.unique(np.geomspace(2, 3000, 16).astype(int)).astype(float)
    cl, stat, st = batch.angular_cl(colls, pairs, ells)
    assert st == 0 and batch.last_used_fast_path
    for ic, c in enumerate(cos):
        c2 = ccl.Cosmology(**{**c._params_init_kwargs, "transfer_function": "eisenstein_hu",
                              "matter_power_spectrum": "halofit"})
        trs = [ccl.NumberCountsTracer(c2, dndz=(z, n), bias=(z, b * np.ones_like(z)), has_rsd=False)
               for n, b in zip(lens, bias)] + [ccl.WeakLensingTracer(c2, dndz=(z, n)) for n in src]
        for q, (i, j) in enumerate(pairs):
            ref = ccl.angular_cl(c2, trs[i], trs[j], ells, limber_integration_method='spline')
            scale = np.max(np.abs(ref))
            if scale == 0:
                assert np.all(cl[ic, q] == 0)
                continue
            err = np.max(np.abs(cl[ic, q] - ref) / np.maximum(np.abs(ref), 1e-6 * scale))
            # tables agree to ~1e-8 (growth ODE, quadratures); the integrator is identical
            assert err < 5e-7, (ic, (i, j), err)


def test_kappa_tracer_from_device_tables():
    """CMB-lensing kappa kernels built on the device (ccl_get_kappa_kernel) from device-built background
    tables: kappa x kappa, kappa x shear and kappa x counts agree with the pyccl-compatible classes; kappa
    reaches a < a_min of P(k,a), so these C_ell also take the growth-extrapolation (redo) route."""
    import ccl_b200 as ccl
    from ccl_b200 import device_builders as db
    from ccl_b200 import synthetic as S
    cos = _cosmologies()[1:4]
    z = np.linspace(0.0, 3.5, 500)
    lens = S.smail_bins(z, 0.26, 0.94, edges=np.arange(0.2, 0.6 + 1e-9, 0.2), sigma_z=0.03)
    src = S.smail_bins(z, 0.13, 0.78, nbins=2, sigma_z=0.05)
    batch, colls, tabs = db.build_3x2pt_batch(cos, z, lens, src, [1.3, 1.5], model="eisenstein_hu", halofit=True)
    ik = tabs.add_kappa_tracer(batch, 1100.0)
    colls = list(colls) + [(ik, 1)]
    batch.commit()
    nk = len(colls) - 1
    pairs = [(nk, nk), (0, nk), (3, nk), (1, 2)]
    ells = np.unique(np.geomspace(2, 2000, 12).astype(int)).astype(float)
    cl, stat, st = batch.angular_cl(colls, pairs, ells)
    assert st == 0 and not stat.any()
    for ic, c in enumerate(cos):
        c2 = ccl.Cosmology(**{**c._params_init_kwargs, "transfer_function": "eisenstein_hu",
                              "matter_power_spectrum": "halofit"})
        trs = [ccl.NumberCountsTracer(c2, dndz=(z, n), bias=(z, b * np.ones_like(z)), has_rsd=False)
               for n, b in zip(lens, [1.3, 1.5])] + [ccl.WeakLensingTracer(c2, dndz=(z, n)) for n in src]
        trs.append(ccl.CMBLensingTracer(c2, z_source=1100.0))
        for q, (i, j) in enumerate(pairs):
            ref = ccl.angular_cl(c2, trs[i], trs[j], ells, limber_integration_method='spline')
            err = np.max(np.abs(cl[ic, q] / ref - 1))
            assert err < 5e-7, (ic, (i, j), err)


def test_all_tracer_flavours_from_device_tables():
    """Number counts with bias + RSD + magnification and shear with intrinsic alignments, every kernel and
    every cosmology-dependent transfer (-f(a), -A_IA rho_m / D(a)) built on the device: C_ell agree with
    the pyccl-compatible classes on host-built tables."""
    import ccl_b200 as ccl
    from ccl_b200 import device_builders as db
    from ccl_b200 import synthetic as S
    cos = _cosmologies()[2:5]
    z = np.linspace(0.0, 3.5, 500)
    lens = S.smail_bins(z, 0.26, 0.94, edges=np.arange(0.2, 0.6 + 1e-9, 0.2), sigma_z=0.03)
    src = S.smail_bins(z, 0.13, 0.78, nbins=2, sigma_z=0.05)
    nl, ns = len(lens), len(src)
    bias = [1.3, 1.5]
    s_mag = 0.2 + 0.1 * z                      # magnification bias s(z)
    a_ia = 0.8 * ((1 + z) / 1.62) ** 0.5       # IA amplitude A(z)
    tabs = db.DeviceTables(cos, z, list(lens) + list(src), mag_bias_list=[s_mag] * (nl + ns),
                           lens_scale=[-2.0] * nl + [1.0] * ns)
    batch = ccl.LibBatch(len(cos))
    tabs.fill(batch)
    a_rev = np.ascontiguousarray(1. / (1 + z[::-1]))
    colls, first = [], 0
    for i in range(nl):
        tabs.add_tracer(batch, "nc", i, 0, 0, transfer_a=(a_rev, np.full(z.size, bias[i])))
        tabs.add_tracer_bg(batch, "nc", i, 2, 0, "rsd", a_rev)
        last = tabs.add_tracer(batch, "lensing", i, -1, 1)
        colls.append((first, last - first + 1))
        first = last + 1
    rho_crit = ccl.physical_constants.RHO_CRITICAL
    # the shear kernels must not carry the magnification factor q(z) = 1 - 2.5 s(z): they come from a
    # second table object built without magnification bias
    tabs_s = db.DeviceTables(cos, z, list(src), lens_scale=[1.0] * ns)
    for j in range(ns):
        tabs_s.add_tracer(batch, "lensing", j, -1, 2)
        last = tabs_s.add_tracer_bg(batch, "nc", j, -1, 2, "ia", a_rev, amp=a_ia[::-1] * 5e-14 * rho_crit)
        colls.append((first, last - first + 1))
        first = last + 1
    batch.commit()
    pairs = [(i, j) for i in range(len(colls)) for j in range(i, len(colls))]
    ells = np.unique(np.geomspace(2, 2000, 10).astype(int)).astype(float)
    cl, stat, st = batch.angular_cl(colls, pairs, ells)
    assert st == 0 and not stat.any()
    for ic, c in enumerate(cos):
        c2 = ccl.Cosmology(**{**c._params_init_kwargs, "transfer_function": "eisenstein_hu",
                              "matter_power_spectrum": "halofit"})
        trs = [ccl.NumberCountsTracer(c2, dndz=(z, n), bias=(z, b * np.ones_like(z)), mag_bias=(z, s_mag), has_rsd=True)
               for n, b in zip(lens, bias)] + [ccl.WeakLensingTracer(c2, dndz=(z, n), ia_bias=(z, a_ia)) for n in src]
        for q, (i, j) in enumerate(pairs):
            ref = ccl.angular_cl(c2, trs[i], trs[j], ells, limber_integration_method='spline')
            scale = np.max(np.abs(ref))
            err = np.max(np.abs(cl[ic, q] - ref) / np.maximum(np.abs(ref), 1e-6 * scale))
            assert err < 2e-6, (ic, (i, j), err)


def test_device_resident_pipeline_equals_round_trip():
    """Tables handed from the builders to the batch inside HBM give the same C_ell, bit for bit, as the
    same tables taken through the host (build_3x2pt_stacked + set_stacked)."""
    import ccl_b200 as ccl
    from ccl_b200 import device_builders as db
    from ccl_b200 import synthetic as S
    cos = _cosmologies()[1:6]
    z = np.linspace(0.0, 3.5, 500)
    lens = S.smail_bins(z, 0.26, 0.94, edges=np.arange(0.2, 0.8 + 1e-9, 0.2), sigma_z=0.03)
    src = S.smail_bins(z, 0.13, 0.78, nbins=4, sigma_z=0.05)
    bias = [1.2, 1.4, 1.6]
    stk, colls = db.build_3x2pt_stacked(cos, z, lens, src, bias)
    b1 = ccl.LibBatch(len(cos))
    b1.set_stacked(0, stk)
    b1.commit()
    b2, colls2, tabs = db.build_3x2pt_batch(cos, z, lens, src, bias)
    assert colls == colls2
    pairs = [(i, j) for i in range(len(colls)) for j in range(i, len(colls))]
    ells = np.unique(np.geomspace(2, 3000, 20).astype(int)).astype(float)
    c1, s1, st1 = b1.angular_cl(colls, pairs, ells)
    c2, s2, st2 = b2.angular_cl(colls, pairs, ells)
    assert st1 == st2 == 0 and b2.last_used_fast_path
    assert [b1.tracer_chi_range(2, t) for t in range(7)] == [b2.tracer_chi_range(2, t) for t in range(7)]
    assert np.array_equal(c1, c2)
    # the batch is reusable: rebuild into the same arena
    b3, _, tabs3 = db.build_3x2pt_batch(cos, z, lens, src, bias, batch=b2)
    c3, _, _ = b3.angular_cl(colls, pairs, ells)
    assert np.array_equal(c1, c3)


def test_api_objects_on_device_builders_match_host_twins():
    """Cosmology / tracer objects built by the device builders (the product default) against the same objects
    built by the numpy twins: tables to the builders' parity tolerances, README-example C_ell to 1e-6."""
    import time
    import ccl_b200 as ccl
    assert ccl.get_table_builder() == "device"
    kw = dict(Omega_c=0.27, Omega_b=0.045, h=0.67, sigma8=0.8, n_s=0.96, transfer_function='bbks')
    z = np.linspace(0., 1., 500)
    n = np.ones_like(z)
    ell = np.arange(2, 2001).astype(float)

    def run():
        c = ccl.Cosmology(**kw)
        t1 = ccl.WeakLensingTracer(c, dndz=(z, n))
        t2 = ccl.WeakLensingTracer(c, dndz=(z, n), ia_bias=(z, n))
        g = ccl.NumberCountsTracer(c, has_rsd=True, dndz=(z, n), bias=(z, np.sqrt(1 + z)), mag_bias=(z, 0.4 * n))
        return c, [ccl.angular_cl(c, a, b, ell, limber_integration_method='spline') for a, b in ((t1, t1), (t1, t2), (g, g), (g, t2))]

    run()                                   # CUDA context, library load
    t0 = time.perf_counter()
    c_dev, cl_dev = run()
    dt_dev = time.perf_counter() - t0
    assert c_dev._table_builder == "device" and c_dev._bgd is not None
    ccl.set_table_builder("host")
    try:
        t0 = time.perf_counter()
        c_host, cl_host = run()
        dt_host = time.perf_counter() - t0
    finally:
        ccl.set_table_builder("device")
    assert c_host._table_builder == "host" and c_host._bgd is None
    print("setup + 4 angular_cl calls: device builders %.1f ms, numpy twins %.1f ms" % (dt_dev * 1e3, dt_host * 1e3))
    for k in ("chi", "E", "growth", "fgrowth"):
        assert np.max(np.abs(getattr(c_dev._bg, k) / np.where(getattr(c_host._bg, k) == 0, 1, getattr(c_host._bg, k)) - 1)[:-1]) < 1e-7, k
    pd, ph = c_dev.get_nonlin_power(), c_host.get_nonlin_power()
    assert np.max(np.abs(pd._tab - ph._tab)) < 1e-6
    for a, b in zip(cl_dev, cl_host):
        assert np.max(np.abs(a - b) / np.max(np.abs(b))) < 1e-6
    assert dt_dev < dt_host
