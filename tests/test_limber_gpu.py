"""GPU parity: the CUDA Limber path (through the C ABI) against the CPU oracle on identical
tables.  Tolerances are BASELINE.json's: 1e-10 relative for 'spline', 1e-5 for 'qag_quad'."""
import numpy as np
import pytest

import oracle
from common import (chi_max_nokernel, edge_mask, edge_mask_subs, oracle_cosmo, oracle_pk, oracle_tracer,
                    parity_error)

pytestmark = pytest.mark.gpu

SPLINE_RTOL = 1e-10
QAG_RTOL = 1e-5
ELLS = np.unique(np.geomspace(2, 3000, 24).astype(int)).astype(float)


def _rel(cl, ref, scale_ref):
    # C_ell that cross zero (IA x shear) are compared against the curve's own scale
    floor = 1e-6 * np.max(np.abs(scale_ref))
    if floor == 0.0:   # identically zero curve (disjoint supports with nk >= 5): must match exactly
        return np.where(cl == ref, 0.0, np.inf)
    return np.abs(cl - ref) / np.maximum(np.abs(ref), floor)


@pytest.fixture(scope="module")
def setup(synth):
    import ccl_b200 as cb
    bg, pk, subs, colls = synth["bg"], synth["pk"], synth["subs"], synth["colls"]
    cmax = chi_max_nokernel(bg)
    ocos = oracle_cosmo(bg)
    opk = oracle_pk(pk)
    otr = [oracle_tracer(s, cmax) for s in subs]
    cos = cb.LibCosmology()
    cos.set_achi(bg["chi"], bg["a"])
    cos.set_growth(bg["a_growth"], bg["growth"])
    cos.set_chi_max_nokernel(cmax)
    psp = cb.LibF2D(a=pk["a"], lk=pk["lk"], fka=pk["logp"], extrap_order_lok=1, extrap_order_hik=2,
                    is_log=True, growth_exponent=2, growth_factor_0=0.0)
    ltr = [cb.LibTracer(cos, **s) for s in subs]
    return dict(ocos=ocos, opk=opk, otr=otr, cos=cos, psp=psp, ltr=ltr, colls=colls)


def _coll(objs, coll):
    return objs[coll[0]:coll[0] + coll[1]]


PAIRS = [(0, 0), (0, 3), (1, 7), (4, 14), (5, 5), (6, 12), (2, 2), (0, 9)]


@pytest.mark.parametrize("pair", PAIRS)
def test_single_pair_spline(setup, pair):
    import ccl_b200 as cb
    s = setup
    c1, c2 = s["colls"][pair[0]], s["colls"][pair[1]]
    ref, st_ref = oracle.angular_cl(s["ocos"], _coll(s["otr"], c1), _coll(s["otr"], c2), s["opk"], ELLS,
                                    oracle.INTEG_SPLINE)
    cl, st = cb.angular_cls_limber(s["cos"], _coll(s["ltr"], c1), _coll(s["ltr"], c2), s["psp"], ELLS,
                                   cb._lib.INTEGRATION_SPLINE)
    assert st == st_ref == 0
    assert np.all(np.isfinite(cl))
    redo = lambda: oracle.angular_cl(s["ocos"], _coll(s["otr"], c1), _coll(s["otr"], c2), s["opk"], ELLS,
                                     oracle.INTEG_SPLINE)[0]
    edge = edge_mask(_coll(s["otr"], c1), _coll(s["otr"], c2), ELLS)
    assert parity_error(cl, ref, redo, SPLINE_RTOL, edge) < SPLINE_RTOL


@pytest.mark.parametrize("pair", PAIRS[:5])
def test_single_pair_qag(setup, pair):
    import ccl_b200 as cb
    s = setup
    c1, c2 = s["colls"][pair[0]], s["colls"][pair[1]]
    ref, st_ref = oracle.angular_cl(s["ocos"], _coll(s["otr"], c1), _coll(s["otr"], c2), s["opk"], ELLS,
                                    oracle.INTEG_QAG)
    cl, st = cb.angular_cls_limber(s["cos"], _coll(s["ltr"], c1), _coll(s["ltr"], c2), s["psp"], ELLS,
                                   cb._lib.INTEGRATION_QAG_QUAD)
    assert st == st_ref == 0
    assert np.max(_rel(cl, ref, ref)) < QAG_RTOL


@pytest.mark.parametrize("pair", PAIRS)
def test_single_pair_cquad_fallback(setup, pair, monkeypatch):
    """The reference retries with gsl_integration_cquad when QAG reports GSL_EROUND
    (src/ccl_cls.c:245-259).  That branch is forced on both sides (no synthetic input of this suite
    triggers it): device CQUAD == oracle CQUAD to the 'qag_quad' tolerance, and both agree with
    QAG-GK41 to its own epsrel."""
    import ccl_b200 as cb
    s = setup
    c1, c2 = s["colls"][pair[0]], s["colls"][pair[1]]
    args = (s["ocos"], _coll(s["otr"], c1), _coll(s["otr"], c2), s["opk"], ELLS, oracle.INTEG_QAG)
    ref_qag, _ = oracle.angular_cl(*args)
    oracle.set_force_cquad(True)
    try:
        ref, st_ref = oracle.angular_cl(*args)
    finally:
        oracle.set_force_cquad(False)
    monkeypatch.setenv("CCL_B200_FORCE_CQUAD", "1")
    cl, st = cb.angular_cls_limber(s["cos"], _coll(s["ltr"], c1), _coll(s["ltr"], c2), s["psp"], ELLS,
                                   cb._lib.INTEGRATION_QAG_QUAD)
    monkeypatch.delenv("CCL_B200_FORCE_CQUAD")
    assert st == st_ref == 0 and np.all(np.isfinite(cl))
    assert np.max(_rel(cl, ref, ref)) < QAG_RTOL
    # two integrators at epsrel = 1e-4, where the curve is not a near-cancellation of disjoint kernels
    big = np.abs(ref_qag) > 1e-3 * np.max(np.abs(ref_qag))
    assert np.max(np.abs(cl[big] / ref_qag[big] - 1)) < 5e-4
    # and the normal path is back afterwards
    cl2, _ = cb.angular_cls_limber(s["cos"], _coll(s["ltr"], c1), _coll(s["ltr"], c2), s["psp"], ELLS,
                                   cb._lib.INTEGRATION_QAG_QUAD)
    assert np.max(_rel(cl2, ref_qag, ref_qag)) < QAG_RTOL


def test_accessors_match_oracle(setup):
    s = setup
    chi = np.linspace(0.0, 9000.0, 301)
    a_gpu, st = s["cos"].scale_factor_of_chi(chi)
    a_ref, st_ref = s["ocos"].scale_factor_of_chi(chi)
    assert st == st_ref == 0
    assert np.max(np.abs(a_gpu - a_ref)) < 1e-14
    for it in (0, 1, 2, 15, 16):
        w = s["ltr"][it].kernel(chi)
        wr = s["otr"][it].kernel(chi)
        assert np.allclose(w, wr, rtol=1e-13, atol=1e-300)
        assert s["ltr"][it].chi_min == s["otr"][it].chi_min
        assert s["ltr"][it].chi_max == s["otr"][it].chi_max
    lk = np.linspace(np.log(1e-5), np.log(3e3), 97)  # beyond both table ends: extrapolation orders 1 / 2
    for a in (1.0, 0.5, 0.0123, 0.004):              # below amin: growth extrapolation
        p, st = s["psp"].eval(lk, a, s["cos"])
        pr = np.array([s["opk"](x, a, s["ocos"])[0] for x in lk])
        assert st == 0
        inside = (lk >= s["opk"]._lk[0]) & (lk <= s["opk"]._lk[-1])
        assert np.max(np.abs(p / pr - 1)[inside]) < 1e-13
        # Outside the table the order-2 term multiplies d2/dlnk2 of a *natural* spline at its end
        # node, i.e. rounding noise of order eps*|z|/dx^2 ~ 1e-10 in both implementations.
        assert np.max(np.abs(p / pr - 1)[~inside]) < 5e-9


# ---------------------------------------------------------------------------------------------
# Batched entry: the grid-sharing fast kernels (limber_fast.cu) and their redo path
# ---------------------------------------------------------------------------------------------
def _fill(batch, ic, bg, pk, subs, cmax, logp_shift=0.0):
    batch.set_cosmology(ic, bg["chi"], bg["a"], bg["a_growth"], bg["growth"], cmax)
    batch.set_pk2d(ic, pk["lk"], pk["a"], pk["logp"] + logp_shift)
    for s in subs:
        batch.add_tracer(ic, **s)


def _oracle_all(bg, pk, subs, colls, pairs, ells, method=oracle.INTEG_SPLINE, logp_shift=0.0):
    cmax = chi_max_nokernel(bg)
    oc = oracle_cosmo(bg)
    op = oracle_pk(dict(pk, logp=pk["logp"] + logp_shift))
    otr = [oracle_tracer(s, cmax) for s in subs]
    out = np.empty((len(pairs), len(ells)))
    stat = np.zeros(len(pairs), dtype=int)
    for q, (i, j) in enumerate(pairs):
        out[q], st = oracle.angular_cl(oc, _coll(otr, colls[i]), _coll(otr, colls[j]), op, ells, method, each=True)
        stat[q] = 1030 if np.any(st) else 0
    return out, stat


@pytest.mark.parametrize("flavour", ["plain", "ia_mag", "rsd_ia_mag"])
def test_batch_fast_path_matches_oracle(flavour):
    """All 120 pairs x 24 ells of two different cosmologies in one launch, against the oracle."""
    import ccl_b200 as cb
    from ccl_b200 import synthetic as S
    kw = {"plain": {}, "ia_mag": dict(with_ia=True, with_mag=True),
          "rsd_ia_mag": dict(with_ia=True, with_mag=True, with_rsd=True)}[flavour]
    batch = cb.LibBatch(2)
    sets = []
    for ic, (Om, h) in enumerate(((0.3, 0.7), (0.26, 0.64))):
        bg = S.background(Om=Om, h=h)
        pk = S.power(bg)
        subs, colls = S.lsst_like_tracers(bg, **kw)
        _fill(batch, ic, bg, pk, subs, chi_max_nokernel(bg), 0.01 * ic)
        sets.append((bg, pk, subs, colls))
    batch.commit()
    colls = sets[0][3]
    pairs = [(i, j) for i in range(len(colls)) for j in range(i, len(colls))]
    cl, stat, st = batch.angular_cl(colls, pairs, ELLS)
    assert batch.last_used_fast_path
    assert st == 0 and not stat.any()
    for ic, (bg, pk, subs, colls_i) in enumerate(sets):
        ref, _ = _oracle_all(bg, pk, subs, colls_i, pairs, ELLS, logp_shift=0.01 * ic)
        for q in range(len(pairs)):
            redo = lambda: _oracle_all(bg, pk, subs, colls_i, [pairs[q]], ELLS, logp_shift=0.01 * ic)[0][0]
            edge = edge_mask_subs(subs, colls_i, pairs[q], ELLS, chi_max_nokernel(bg))
            assert parity_error(cl[ic, q], ref[q], redo, SPLINE_RTOL, edge) < SPLINE_RTOL, (ic, pairs[q])


def test_batch_fast_equals_general_kernel(monkeypatch):
    """Same launch through the general one-warp-per-C_ell kernel: identical to rounding."""
    import ccl_b200 as cb
    from ccl_b200 import synthetic as S
    bg = S.background()
    pk = S.power(bg)
    subs, colls = S.lsst_like_tracers(bg, with_ia=True, with_rsd=True, with_mag=True)
    pairs = [(i, j) for i in range(len(colls)) for j in range(i, len(colls))]
    ells = np.geomspace(2, 3000, 100)
    batch = cb.LibBatch(1)
    _fill(batch, 0, bg, pk, subs, chi_max_nokernel(bg))
    batch.commit()
    fast, s1, _ = batch.angular_cl(colls, pairs, ells)
    assert batch.last_used_fast_path
    monkeypatch.setenv("CCL_B200_NO_FAST", "1")
    gen, s2, _ = batch.angular_cl(colls, pairs, ells)
    assert not batch.last_used_fast_path
    assert not s1.any() and not s2.any()
    scale = np.max(np.abs(gen), axis=1, keepdims=True)
    assert np.max(np.abs(fast - gen) / np.maximum(np.abs(gen), 1e-6 * scale)) < 1e-12


def test_batch_fast_redo_and_failures():
    """Nodes below the P(k,a) table's amin go through the redo list (growth extrapolation in the
    general kernel); nearly disjoint supports fail with NaN + CCL_ERROR_INTEG like the reference."""
    import ccl_b200 as cb
    from ccl_b200 import synthetic as S
    bg = S.background()
    pk = S.power(bg)
    keep = pk["a"] >= 0.45                           # amin = 0.45: sources at z > 1.2 leave the table
    pk_cut = dict(lk=pk["lk"], a=pk["a"][keep], logp=pk["logp"][keep])
    subs, colls = S.lsst_like_tracers(bg)
    chi = bg["chi"]
    w1 = np.where((chi > 1000) & (chi < 1040), 1.0, 0.0)
    w2 = np.where((chi > 2000) & (chi < 2040), 1.0, 0.0)   # chi_max1 / chi_min2 < 0.546 -> nk < 5
    subs = subs + [dict(der_bessel=0, der_angles=0, chi=chi, w=w1), dict(der_bessel=0, der_angles=0, chi=chi, w=w2)]
    colls = colls + [(15, 1), (16, 1)]
    pairs = [(0, 0), (2, 9), (4, 14), (12, 14), (14, 14), (15, 16), (15, 15)]
    batch = cb.LibBatch(1)
    _fill(batch, 0, bg, pk_cut, subs, chi_max_nokernel(bg))
    batch.commit()
    cl, stat, st = batch.angular_cl(colls, pairs, ELLS)
    assert batch.last_used_fast_path
    ref, rstat = _oracle_all(bg, pk_cut, subs, colls, pairs, ELLS)
    assert rstat[5] == 1030 and stat[0, 5] == 1030 and st == 1030
    assert np.array_equal(np.isnan(cl[0]), np.isnan(ref))
    assert np.array_equal(stat[0] != 0, rstat != 0)
    ok = ~np.isnan(ref)
    for q in range(len(pairs)):
        if ok[q].any():
            redo = lambda: _oracle_all(bg, pk_cut, subs, colls, [pairs[q]], ELLS)[0][0][ok[q]]
            edge = edge_mask_subs(subs, colls, pairs[q], ELLS, chi_max_nokernel(bg))[ok[q]]
            assert parity_error(cl[0, q][ok[q]], ref[q][ok[q]], redo, SPLINE_RTOL, edge) < SPLINE_RTOL, pairs[q]


def test_batch_stacked_setters_equal_per_slot():
    """ccl_b200_batch_*_stacked fill the arena exactly like the per-slot setters."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import ccl_b200 as cb
    tables = bench.build_host_tables(3)
    colls, pairs, ells = bench.geometry()
    ells = ells[::9]
    n = 5
    b1 = cb.LibBatch(n)
    bench.fill_batch(b1, tables, n, 1)
    b1.commit()
    b2 = cb.LibBatch(n)
    b2.set_stacked(0, bench.stack_host_tables(tables, n, 1))
    b2.commit()
    c1, s1, _ = b1.angular_cl(colls, pairs, ells)
    c2, s2, _ = b2.angular_cl(colls, pairs, ells)
    assert np.array_equal(c1, c2) and np.array_equal(s1, s2)
    assert [b1.tracer_chi_range(4, t) for t in range(15)] == [b2.tracer_chi_range(4, t) for t in range(15)]


# ---------------------------------------------------------------------------------------------
# Full-size properties (BASELINE configs[3]: 4096 cosmologies x 120 pairs x 100 ells) and edge cases
# ---------------------------------------------------------------------------------------------
def test_full_size_batch_properties():
    """Size-independent properties at the bench's full size: the 4096 slots cycle through 8 table sets
    with log P shifted by 1e-3 per cycle, so (i) C_ell is linear in P: slots of the same set differ by
    exactly exp(shift); (ii) every entry is finite and flagged OK; (iii) a sample of (cosmology,
    pair) curves matches the oracle; (iv) a second launch is bit-identical (determinism)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import ccl_b200 as cb
    n = 4096
    tables = bench.build_host_tables(bench.N_DISTINCT)
    colls, pairs, ells = bench.geometry()
    batch = cb.LibBatch(n)
    batch.set_stacked(0, bench.stack_host_tables(tables, n, 0))
    batch.commit()
    cl, stat, st = batch.angular_cl(colls, pairs, ells)
    assert batch.last_used_fast_path and st == 0 and not stat.any()
    assert cl.shape == (n, len(pairs), len(ells)) and np.all(np.isfinite(cl))
    nd = bench.N_DISTINCT
    base = cl[:nd]
    scale = np.max(np.abs(base), axis=2, keepdims=True)
    for cyc in (1, 17, n // nd - 1):
        blk = cl[cyc * nd:(cyc + 1) * nd]
        denom = np.maximum(np.abs(base), 1e-6 * scale)
        diff = np.abs(blk - base * np.exp(1e-3 * cyc))
        zero = denom == 0.0                      # identically zero curves (disjoint supports) stay zero
        assert np.all(diff[zero] == 0.0)
        assert np.max(diff[~zero] / denom[~zero]) < 1e-11, cyc
    cl2, _, _ = batch.angular_cl(colls, pairs, ells)
    assert np.array_equal(cl, cl2)
    rng = np.random.default_rng(5)
    for ic, q in zip(rng.integers(0, n, 64), rng.integers(0, len(pairs), 64)):
        t = tables[ic % nd]
        shift = 1e-3 * (ic // nd)
        redo = lambda: _oracle_all(t["bg"], t["pk"], t["subs"], t["colls"], [pairs[q]], ells, logp_shift=shift)[0][0]
        edge = edge_mask_subs(t["subs"], t["colls"], pairs[q], ells, t["cmax"])
        assert parity_error(cl[ic, q], redo(), redo, SPLINE_RTOL, edge) < SPLINE_RTOL, (ic, pairs[q])


def test_batch_edge_cases(synth):
    """Empty ell / pair lists, ell <= 1 (f_ell = 0 for spin-2 tracers: src/ccl_tracers.c:707-711), a pair
    listed twice and in both orders, and the qag_quad method through the batched entry."""
    import ccl_b200 as cb
    bg, pk = synth["bg"], synth["pk"]
    from ccl_b200 import synthetic as S
    subs, colls = S.lsst_like_tracers(bg)
    batch = cb.LibBatch(1)
    _fill(batch, 0, bg, pk, subs, chi_max_nokernel(bg))
    batch.commit()
    cl, stat, st = batch.angular_cl(colls, [(0, 5)], np.zeros(0))
    assert cl.shape == (1, 1, 0) and st == 0
    cl, stat, st = batch.angular_cl(colls, [], np.array([10.0]))
    assert cl.shape == (1, 0, 1) and st == 0
    pairs = [(7, 9), (9, 7), (7, 9), (5, 5), (0, 0)]
    ells = np.array([0.0, 1.0, 2.0, 3.5, 10.0, 11.0, 1000.0, 1001.0, 40000.0])
    cl, stat, st = batch.angular_cl(colls, pairs, ells)
    assert st == 0 and not stat.any()
    assert np.array_equal(cl[0, 0], cl[0, 2])
    assert np.max(np.abs(cl[0, 0] - cl[0, 1]) / np.abs(cl[0, 0]).max()) < 1e-13
    assert np.all(cl[0, 0][:2] == 0.0) and np.all(cl[0, 3][:2] == 0.0)      # shear x shear at ell = 0, 1
    ref, rstat = _oracle_all(bg, pk, subs, colls, pairs, ells)
    for q in range(len(pairs)):
        redo = lambda: _oracle_all(bg, pk, subs, colls, [pairs[q]], ells)[0][0]
        edge = edge_mask_subs(subs, colls, pairs[q], ells, chi_max_nokernel(bg))
        assert parity_error(cl[0, q], ref[q], redo, SPLINE_RTOL, edge) < SPLINE_RTOL, pairs[q]
    clq, statq, stq = batch.angular_cl(colls, pairs, ells[2:7], cb._lib.INTEGRATION_QAG_QUAD)
    refq, _ = _oracle_all(bg, pk, subs, colls, pairs, ells[2:7], oracle.INTEG_QAG)
    assert stq == 0
    for q in range(len(pairs)):
        assert parity_error(clq[0, q], refq[q], None, QAG_RTOL) < QAG_RTOL


def test_batch_cquad_fallback(synth, monkeypatch):
    """The CQUAD hand-over list through the batched entry: forced on both sides, every C_ell of the batch
    goes QAG kernel -> list -> limber_cquad_kernel and matches the oracle's CQUAD."""
    import ccl_b200 as cb
    bg, pk = synth["bg"], synth["pk"]
    from ccl_b200 import synthetic as S
    subs, colls = S.lsst_like_tracers(bg)
    batch = cb.LibBatch(2)
    for ic in range(2):
        _fill(batch, ic, bg, pk, subs, chi_max_nokernel(bg))
    batch.commit()
    pairs = [(0, 0), (2, 9), (7, 9), (14, 14)]
    ells = ELLS[::3]
    oracle.set_force_cquad(True)
    try:
        ref, _ = _oracle_all(bg, pk, subs, colls, pairs, ells, oracle.INTEG_QAG)
    finally:
        oracle.set_force_cquad(False)
    monkeypatch.setenv("CCL_B200_FORCE_CQUAD", "1")
    cl, stat, st = batch.angular_cl(colls, pairs, ells, cb._lib.INTEGRATION_QAG_QUAD)
    monkeypatch.delenv("CCL_B200_FORCE_CQUAD")
    assert st == 0 and not stat.any() and np.array_equal(cl[0], cl[1])
    for q in range(len(pairs)):
        assert parity_error(cl[0, q], ref[q], None, QAG_RTOL) < QAG_RTOL, pairs[q]


def test_pipelined_host_entry_equals_single_batch():
    """PipelinedBatch (chunked stage/commit/compute on three threads) == one LibBatch, bit for bit."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import ccl_b200 as cb
    tables = bench.build_host_tables(3)
    colls, pairs, ells = bench.geometry()
    ells = ells[::7]
    n = 37                                      # ragged chunks
    stk = bench.stack_host_tables(tables, n, 2)
    b1 = cb.LibBatch(n)
    b1.set_stacked(0, stk)
    b1.commit()
    c1, s1, st1 = b1.angular_cl(colls, pairs, ells)
    pipe = cb.PipelinedBatch(n, nchunks=5)
    for _ in range(2):                          # arenas are reused from step to step
        c2, s2, st2 = pipe.angular_cl(stk, colls, pairs, ells)
        assert np.array_equal(c1, c2) and np.array_equal(s1, s2) and st1 == st2 == 0


def test_pinned_stacked_inputs_are_not_staged():
    """Stacked arrays in page-locked memory are DMA'd straight from the caller's buffers (landing
    region) instead of being staged: same C_ell bit for bit, through LibBatch and PipelinedBatch; the
    landed tables survive a second commit after the sources are gone and a later per-slot addition."""
    import sys, os, gc
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import ccl_b200 as cb
    tables = bench.build_host_tables(3)
    colls, pairs, ells = bench.geometry()
    ells = ells[::7]
    n = 21
    stk = bench.stack_host_tables(tables, n, 1)
    b1 = cb.LibBatch(n)
    b1.set_stacked(0, stk)
    b1.commit()
    c1, s1, _ = b1.angular_cl(colls, pairs, ells)
    pinned = cb.pin_stacked(stk)
    b2 = cb.LibBatch(n)
    b2.set_stacked(0, pinned)
    b2.commit()
    c2, s2, _ = b2.angular_cl(colls, pairs, ells)
    assert np.array_equal(c1, c2) and np.array_equal(s1, s2)
    assert [b1.tracer_chi_range(4, t) for t in range(15)] == [b2.tracer_chi_range(4, t) for t in range(15)]
    pipe = cb.PipelinedBatch(n, nchunks=4)
    for _ in range(2):
        c3, s3, st3 = pipe.angular_cl(pinned, colls, pairs, ells)
        assert np.array_equal(c1, c3) and np.array_equal(s1, s3) and st3 == 0
    # sources released, then one more sub-tracer added per slot and a second commit
    for t in pinned["tracers"]:
        for k in ("chi", "w", "fa"):
            if t.get(k) is not None:
                t[k][...] = np.nan
    pinned["pk"][2][...] = np.nan
    del pinned
    gc.collect()
    sub = tables[0]["subs"][7]
    for ic in range(n):
        b2.add_tracer(ic, chi=sub["chi"], w=sub["w"])
    b2.commit()
    c4, s4, _ = b2.angular_cl(colls, pairs, ells)
    assert np.array_equal(c1, c4) and np.array_equal(s1, s4)


def test_batch_fallbacks_to_general_kernel(synth):
    """(i) a collection with more sub-tracers than a fast-path task holds goes through the redo list;
    (ii) a table form outside the fast path (2-D k-a transfer) sends the whole batch to the general
    kernel.  Both must still match the oracle."""
    import ccl_b200 as cb
    from ccl_b200 import synthetic as S
    bg, pk = synth["bg"], synth["pk"]
    subs0, colls0 = S.lsst_like_tracers(bg, n_lens=2, n_src=2)
    cmax = chi_max_nokernel(bg)
    ells = ELLS[::3]
    # (i) one collection made of 25 copies of a shear kernel scaled by 1/25 each: same C_ell as one
    big = [dict(subs0[2], w=subs0[2]["w"] / 25.0) for _ in range(25)]
    subs = subs0 + big
    colls = colls0 + [(len(subs0), 25)]
    pairs = [(0, 4), (2, 4), (4, 4), (2, 2), (1, 3)]
    batch = cb.LibBatch(1)
    _fill(batch, 0, bg, pk, subs, cmax)
    batch.commit()
    cl, stat, st = batch.angular_cl(colls, pairs, ells)
    assert batch.last_used_fast_path and st == 0
    ref, _ = _oracle_all(bg, pk, subs, colls, pairs, ells)
    for q in range(len(pairs)):
        redo = lambda: _oracle_all(bg, pk, subs, colls, [pairs[q]], ells)[0][0]
        edge = edge_mask_subs(subs, colls, pairs[q], ells, chi_max_nokernel(bg))
        assert parity_error(cl[0, q], ref[q], redo, SPLINE_RTOL, edge) < SPLINE_RTOL, pairs[q]
    assert np.max(np.abs(cl[0, 2] / cl[0, 3] - 1)) < 1e-12          # 25 x (w/25) == w
    # (ii) a k-dependent 2-D transfer on one sub-tracer
    lk_t = np.linspace(np.log(1e-4), np.log(50.0), 40)
    a_t = np.linspace(0.2, 1.0, 16)
    tka = (1.0 + 0.1 * np.tanh(lk_t))[None, :] * (0.5 + a_t)[:, None]
    subs2 = [dict(s) for s in subs0]
    subs2[0] = dict(der_bessel=0, der_angles=0, chi=subs0[0]["chi"], w=subs0[0]["w"], a=a_t, lk=lk_t, fka=tka)
    batch2 = cb.LibBatch(1)
    _fill(batch2, 0, bg, pk, subs2, cmax)
    batch2.commit()
    pairs2 = [(0, 0), (0, 2), (1, 3)]
    cl2, stat2, st2 = batch2.angular_cl(colls0, pairs2, ells)
    assert not batch2.last_used_fast_path and st2 == 0
    ref2, _ = _oracle_all(bg, pk, subs2, colls0, pairs2, ells)
    for q in range(len(pairs2)):
        redo = lambda: _oracle_all(bg, pk, subs2, colls0, [pairs2[q]], ells)[0][0]
        edge = edge_mask_subs(subs2, colls0, pairs2[q], ells, cmax)
        assert parity_error(cl2[0, q], ref2[q], redo, SPLINE_RTOL, edge) < SPLINE_RTOL, pairs2[q]
