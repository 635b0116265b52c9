"""-m gpu tests of the McSAS-compatible host (mcsas_b200.mcsas.McSAS) -- written like a
user of the reference would drive it (McSAS.factory()(), .model, .data, .calc(),
.result[0], param.histograms())."""
import os

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu
NM = 1e-9


def _cfg1():
    from mcsas_b200.data import SASData
    g = common.load_chain("sphere_cfg1")
    return g, SASData(g["q"], g["I"], g["sigma"])


def test_calc_result_dictionary_and_histogram():
    from mcsas_b200.mcsas import McSAS
    from mcsas_b200.models import Sphere
    from mcsas_b200.parameter import Histogram
    g, data = _cfg1()
    algo = McSAS.factory()()
    algo.seed = 99
    algo.model = Sphere()
    algo.model.radius.setActiveRange((1 * NM, 100 * NM))
    algo.model.radius.histograms().append(
        Histogram(algo.model.radius, 1 * NM, 100 * NM, 50, 'log', 'vol'))
    algo.data = data
    algo.numReps.setValue(10)
    algo.calc()
    assert len(algo.result) == 1
    res = algo.result[0]
    for key in ("contribs", "fitMeasValMean", "fitMeasValStd", "fitX0", "dataX0", "dataMean",
                "dataStd", "scaling", "background", "times", "numIter"):       # mcsas.py:268-285
        assert key in res, key
    assert res["contribs"].shape == (300, 1, 10)
    assert res["fitMeasValMean"].shape == (1, 100)
    assert len(res["scaling"]) == 2 and len(res["background"]) == 2
    assert (res["conval"] <= 1.0).all()
    # reference, same configuration: 30 282 iterations/rep on average,
    # scaling 9.95e35 for seed 12345, vol-weighted mean radius 9.7455e-9 +- 3.06e-11 m
    assert 15000 < res["numIter"] < 60000
    assert abs(res["scaling"][0] / 9.95e35 - 1) < 0.05
    h = algo.model.radius.histograms()[0]
    mean, mean_std = h.moments.mean
    assert abs(mean - 9.7455e-9) < 4 * 3.06e-11
    var, var_std = h.moments.variance
    assert abs(var - 2.019e-18) < 5 * 1.45e-19
    assert h.bins.mean.shape == (50,) and h.cdf.mean[-1] == pytest.approx(1.0)
    # fit reproduces the data within its uncertainty on average
    chi = (((res["fitMeasValMean"][0] - res["dataMean"]) / res["dataStd"]) ** 2).mean()
    assert chi < 1.5
    assert len(algo.model.radius.activeValues()) == 10


def test_mcfit_api_and_deterministic_stream():
    from mcsas_b200 import FP64
    from mcsas_b200.mcsas import McSAS
    from mcsas_b200.models import Sphere
    g, data = _cfg1()
    algo = McSAS(precision=FP64)
    algo.model = Sphere()
    algo.model.radius.setActiveRange((1 * NM, 100 * NM))
    algo.data = data
    algo.maxIterations.setValue(g["num_iter"])
    rset, meas, conval, details = algo.mcFit(300, True, True, nRun=0,
                                             proposalStream=g["stream"])
    assert details["numIterations"] == g["num_iter"] and details["numMoves"] == g["num_moves"]
    assert abs(conval / g["conval"] - 1) < 1e-6
    np.testing.assert_allclose(rset, g["rset"], rtol=1e-12)
    np.testing.assert_allclose(meas, g["fit_meas"], rtol=1e-6)
    assert abs(details["scaling"] / g["scaling"] - 1) < 1e-6


def test_non_convergence_returns_no_result_unless_show_incomplete():
    from mcsas_b200.mcsas import McSAS
    from mcsas_b200.models import Sphere
    g, data = _cfg1()
    algo = McSAS(seed=5)
    algo.model = Sphere()
    algo.model.radius.setActiveRange((1 * NM, 100 * NM))
    algo.data = data
    algo.numReps.setValue(3)
    algo.maxIterations.setValue(200)          # cannot converge in 200 steps
    algo.maxRetries.setValue(2)
    algo.calc()
    assert algo.result == []                  # mcsas.py:221-230
    assert (algo.details["attempts"] == 3).all()      # maxRetries + 1 attempts (measured on the reference)
    algo.showIncomplete.setValue(True)
    algo.calc()
    assert len(algo.result) == 1 and (algo.result[0]["conval"] > 1.0).all()


def test_model_calc_on_gpu_matches_reference_vectors():
    from mcsas_b200.models import MODELS
    from mcsas_b200.data import SASData
    g = np.load(os.path.join(common.GOLDEN, "eval_models.npz"))
    q = g["q"]
    data = SASData(q, np.ones_like(q), np.ones_like(q))
    for name in ("Sphere", "EllipsoidsIsotropic", "GaussianChain"):
        m = MODELS[name]()
        for p in m.params():
            if hasattr(p, "setActive"):
                p.setActive(False)
        for p, v in zip(m.params(), g[name + "_pvalues"]):
            p.setValue(type(p())(v), clip=False)
        for a in g[name + "_active"]:
            getattr(m, str(a)).setActive(True)
        md = m.calc(data, g[name + "_pset"], 0.6666666)
        np.testing.assert_allclose(md.cumInt, g[name + "_it"].sum(0), rtol=1e-9)
        np.testing.assert_allclose(md.vset, g[name + "_v"], rtol=1e-12)
        np.testing.assert_allclose(md.wset, g[name + "_w"], rtol=1e-12)


@pytest.mark.parametrize("case,reps", [("config2", 100), ("sphere200", 100), ("config3_reduced", 48)])
def test_production_statistics_within_the_references_own_spread(case, reps):
    """north_star, production mode: "volume-weighted distribution moments and histograms must agree
    with the reference within its own repetition-to-repetition standard deviation".
    tests/golden/stats_<case>.npz holds 16 repetitions of the UNMODIFIED reference (analyse + histogram,
    oracle/gen_stats_golden.py) on: config2 = the headline workload (BASELINE configs[1]: 1000 q, 300
    contributions), sphere200 (200 q, 200 contributions), config3_reduced (CylindersIsotropic, radius +
    aspect active, 60 q, 60 contributions, 20000 iterations; does not reach chi2 <= 1 in the reference
    either: every repetition uses all its attempts, showIncomplete).  The GPU fit (fp32 production
    kernel, Philox proposals) must put the mean AND the variance of every active parameter's volume-
    weighted distribution within one reference standard deviation of the reference's mean, and the
    histogram bins within 1.5 of theirs."""
    from mcsas_b200.data import SASData
    from mcsas_b200.mcsas import McSAS
    from mcsas_b200.models import MODELS
    from mcsas_b200.parameter import Histogram
    g = np.load(os.path.join(common.GOLDEN, "stats_%s.npz" % case))
    model = MODELS[str(g["model"])]()
    for p in model.params():
        if hasattr(p, "setActive"):
            p.setActive(False)
    active = [str(a) for a in g["active"]]
    for name, lo, hi in zip(active, g["lo"], g["hi"]):
        p = getattr(model, name)
        p.setActive(True)
        p.setActiveRange((float(lo), float(hi)))
        cfg = g["%s_vol_cfg" % name]
        p.histograms().append(Histogram(p, float(cfg[0]), float(cfg[1]), int(cfg[2]), str(g["%s_vol_scale" % name]), "vol"))
    algo = McSAS.factory()()
    algo.seed = 20261018
    algo.model = model
    algo.data = SASData(g["q"], g["I"], g["sigma"])
    algo.numContribs.setValue(int(g["N"]))
    algo.numReps.setValue(reps)
    algo.maxIterations.setValue(int(g["max_iter"]))
    algo.convergenceCriterion.setValue(float(g["conv_crit"]))
    converges = bool((g["conval"] <= float(g["conv_crit"])).all())
    algo.showIncomplete.setValue(not converges)
    algo.calc()
    assert len(algo.result) == 1
    res = algo.result[0]
    if converges:
        assert (res["conval"] <= float(g["conv_crit"])).all()
        # the chains are as long as the reference's: means within 3 standard errors of their difference
        # (chain lengths are heavy-tailed: standard deviation ~ half the mean)
        mine, ref_it = res["numIterAll"].astype(float), g["num_iter"].astype(float)
        se = np.sqrt(mine.var(ddof=1) / mine.size + ref_it.var(ddof=1) / ref_it.size)
        assert abs(mine.mean() - ref_it.mean()) < 3 * se, (mine.mean(), ref_it.mean(), se)
    else:
        # neither implementation converges here: the same number of attempts, a similar final chi2
        assert (res["attempts"] == int(g["attempts"].max())).all()
        assert 0.5 < np.median(res["conval"]) / np.median(g["conval"]) < 2.0
    for name in active:
        h = getattr(model, name).histograms()[0]
        ref = g["%s_vol_moments" % name]                  # (value, std) x (total, mean, variance, skew, kurtosis)
        mean, mean_std = h.moments.mean
        var, var_std = h.moments.variance
        # within the reference's repetition-to-repetition standard deviation
        assert abs(mean - ref[2]) < ref[3], (name, "mean", mean, ref[2], ref[3])
        assert abs(var - ref[4]) < ref[5], (name, "variance", var, ref[4], ref[5])
        # and the GPU's own spread is of the same size
        assert 0.4 < mean_std / ref[3] < 2.5, (name, mean_std, ref[3])
        tot = h.moments.total
        assert abs(tot[0] - ref[0]) < max(ref[1], 1e-3 * abs(ref[0])), (name, "total", tot, ref[:2])
        bm, bs = np.asarray(h.bins.mean), np.asarray(h.bins.std)
        rb, rs = g["%s_vol_bins_mean" % name], g["%s_vol_bins_std" % name]
        tol = 1.5 * rs + 0.01 * rb.max()
        assert (np.abs(bm - rb) < tol).all(), (name, np.abs(bm - rb) / tol)
        np.testing.assert_allclose(h.cdf.mean, g["%s_vol_cdf_mean" % name], atol=0.03)


def test_stop_from_another_thread_reaches_the_scheduler():
    """McSAS.stop (mcsas.py:357) may be set while a fit runs (the reference's GUI does it from the Qt
    event loop): the scheduler polls it at every time-slice boundary; the run ends early and, like the
    reference (mcsas.py:238-243), returns no result unless showIncomplete."""
    import threading
    import time
    from mcsas_b200.mcsas import McSAS
    from mcsas_b200.models import Sphere
    g, data = _cfg1()
    algo = McSAS(seed=1)
    algo.model = Sphere()                     # default range 1-1000 nm: never converges
    algo.data = data
    algo.numReps.setValue(4)
    algo.maxIterations.setValue(10 ** 9)
    algo.convergenceCriterion.setValue(1e-9)
    t = threading.Timer(0.05, lambda: setattr(algo, "stop", True))
    t0 = time.perf_counter()
    t.start()
    algo.calc()
    el = time.perf_counter() - t0
    t.join()
    assert el < 5.0 and algo.stop
    assert algo.result == []
    assert (algo.details["num_iter"] > 0).all() and (algo.details["num_iter"] < 10 ** 8).all()


def test_stop_flag():
    from mcsas_b200.mcsas import McSAS
    from mcsas_b200.models import Sphere
    g, data = _cfg1()
    algo = McSAS(seed=1)
    algo.model = Sphere()                     # default range 1-1000 nm: never converges
    algo.data = data
    algo.numReps.setValue(2)
    algo.showIncomplete.setValue(True)
    algo.maxIterations.setValue(10 ** 9)
    algo.maxRetries.setValue(1)
    algo.stop = True                          # McSAS.stop, mcsas.py:357
    algo.analyse()
    assert (algo.details["num_iter"] < 10 ** 7).all()


def test_smearing_profile_on_the_data_reaches_the_kernels():
    """data.smearing (mcsas_b200.smearing.TrapezoidSmearing) is applied for a canSmear model in
    model.calc and in McSAS.mcFit, and ignored when disabled (sasmodel.py:56-60)."""
    from mcsas_b200 import synth
    from mcsas_b200.data import SASData
    from mcsas_b200.mcsas import McSAS
    from mcsas_b200.models import Sphere, GaussianChain
    from mcsas_b200.smearing import TrapezoidSmearing
    q, I, s = synth.sphere_curve(nq=60)
    data = SASData(q, I, s)
    m = Sphere()
    m.radius.setActiveRange((1 * NM, 100 * NM))
    plain = m.calc(data, np.array([[10 * NM]]), 0.6666666).cumInt
    data.smearing = TrapezoidSmearing(0.02e9, 0.05e9)
    smeared = m.calc(data, np.array([[10 * NM]]), 0.6666666).cumInt
    assert np.abs(smeared / plain - 1).max() > 0.05
    data.smearing = TrapezoidSmearing(0.02e9, 0.05e9, doSmear=False)
    np.testing.assert_array_equal(m.calc(data, np.array([[10 * NM]]), 0.6666666).cumInt, plain)
    gc = GaussianChain()                                    # canSmear = False: profile ignored
    data.smearing = TrapezoidSmearing(0.02e9, 0.05e9)
    a = gc.calc(data, np.array([[10 * NM]]), 0.6666666).cumInt
    data.smearing = None
    np.testing.assert_array_equal(gc.calc(data, np.array([[10 * NM]]), 0.6666666).cumInt, a)
    # a smeared fit runs and improves
    data.smearing = TrapezoidSmearing(0.02e9, 0.05e9)
    algo = McSAS(seed=3)
    algo.model, algo.data = m, data
    algo.maxIterations.setValue(1500)
    rset, conval, details = algo.mcFit(40, False, True, nRun=0)
    assert np.isfinite(conval) and details["numMoves"] > 10


# ---- batched data preparation on the device -------------------------------------
@pytest.mark.parametrize("tag", ["sphere10_default", "quick_default", "quick_nbin50_lim", "quick_nobin"])
def test_batched_dataprep_matches_reference(tag):
    """mcsas_prepare_datasets against the vectors the unmodified reference hands to mcFit after
    loaddatafile -> updateFuMin -> setConfig (tests/golden/dataprep.npz) and, bit for bit, against
    the host mirror (mcsas_b200.dataprep.prepare, itself pinned to the same golden on CPU)."""
    from mcsas_b200 import Engine, dataprep
    g = np.load(os.path.join(common.GOLDEN, "dataprep.npz"))
    raw, cfg = g[tag + "_raw"], g[tag + "_cfg"]
    kw = dict(n_bin=cfg[0], fu_min=cfg[1], x0_low=cfg[2], x0_high=cfg[3], mask_zero=bool(cfg[4]),
              mask_neg=bool(cfg[5]))
    eng = Engine(0)
    (q, I, s), = dataprep.prepare_batch([(raw[:, 0], raw[:, 1], raw[:, 2])], eng, **kw)
    eng.close()
    assert q.size == g[tag + "_q"].size
    np.testing.assert_allclose(q, g[tag + "_q"], rtol=1e-14)
    np.testing.assert_allclose(I, g[tag + "_I"], rtol=1e-14)
    np.testing.assert_allclose(s, g[tag + "_sigma"], rtol=1e-13)
    d = dataprep.prepare(raw[:, 0], raw[:, 1], raw[:, 2], **kw)
    np.testing.assert_array_equal(q, d.q)
    np.testing.assert_array_equal(I, d.f.binnedData)
    np.testing.assert_array_equal(s, d.f.binnedDataU)


@pytest.mark.parametrize("n_bin,with_sigma", [(100, True), (37, False), (0, True), (6, True)])
def test_batched_dataprep_ragged_batch_is_bit_identical_to_the_host_mirror(n_bin, with_sigma):
    """One device call for 60 raw tables of 3 .. 6000 rows with NaN / zero / negative intensities,
    invalid uncertainties, q limits and masks: every prepared vector equals numpy's bits (pairwise
    sums incl. bins of > 128 rows, max(SEM, propagated), empty bins dropped, one-row bins kept)."""
    from mcsas_b200 import Engine, dataprep
    rng = np.random.default_rng(11)
    tables = []
    for i in range(60):
        n = int(rng.choice([3, 8, 9, 129, 500, 1000, 6000])) if i else 6000
        q = np.sort(rng.uniform(0.01, 5.0, n))
        I = np.abs(rng.normal(100., 30., n)) * q ** -2
        I[rng.integers(0, n, max(1, n // 50))] = rng.choice([np.nan, 0.0, -3.0, np.inf])
        s = 0.02 * np.abs(I) * rng.uniform(0.1, 3.0, n)
        s[rng.integers(0, n, max(1, n // 80))] = rng.choice([np.nan, 0.0, np.inf])
        tables.append((q, I, s if with_sigma else None))
    kw = dict(n_bin=n_bin, fu_min=0.015, x0_low=0.03e9, x0_high=4.5e9, mask_zero=True, mask_neg=(n_bin != 37))
    eng = Engine(0)
    out = dataprep.prepare_batch(tables, eng, **kw)
    eng.close()
    n_multi = 0
    for (q, I, s), t in zip(out, tables):
        d = dataprep.prepare(t[0], t[1], t[2], **kw)
        np.testing.assert_array_equal(q, d.q)
        np.testing.assert_array_equal(I, d.f.binnedData)
        np.testing.assert_array_equal(s, d.f.binnedDataU)
        n_multi += q.size
    assert n_multi > 100


def test_integration_stub_drives_the_unmodified_reference_objects():
    """INTEGRATION.md way B executed as a McSAS maintainer would use it: the reference's OWN McSAS /
    Sphere / SASData objects (oracle/_ref or /root/reference through the test harness) handed to the
    stub's analyse_b200(); the contribution sets it returns go through the reference's own
    McSAS.histogram and land on the reference's statistics (stats_sphere200.npz)."""
    import refharness as rh
    if not rh.available():
        pytest.skip("no copy of the reference (oracle/_ref is vendored by __graft_entry__.build())")
    from test_host_cpu import _integration_stub
    from mcsas_b200 import _lib
    rh.install()
    from mcsas.utils.parameter import Histogram
    g = np.load(os.path.join(common.GOLDEN, "stats_sphere200.npz"))
    model = rh.model_class("Sphere")()
    model.radius.setActiveRange((float(g["lo"][0]), float(g["hi"][0])))
    data = rh.make_data(g["q"] * 1e-9, g["I"], g["sigma"], nbin=0)
    np.testing.assert_array_equal(np.asarray(data.f.binnedDataU).ravel(), g["sigma"])
    algo = rh.make_algo(model, data, numContribs=int(g["N"]), numReps=24)
    ns = _integration_stub()
    contribs, meas, sc, bg, conval, nit, el = ns["analyse_b200"](algo, lib_path=_lib.LIB_PATH, device=0, seed=77)
    assert contribs.shape == (int(g["N"]), 1, 24) and meas.shape == (1, g["q"].size, 24)
    assert (conval <= 1.0).all()
    cfg = g["radius_vol_cfg"]
    h = Histogram(model.radius, float(cfg[0]), float(cfg[1]), int(cfg[2]), str(g["radius_vol_scale"]), "vol")
    model.radius.histograms().append(h)
    algo.result = [dict(contribs=contribs)]
    algo.histogram()                                   # the reference's own post-processing
    ref = g["radius_vol_moments"]
    mean, var = h.moments.fields[2], h.moments.fields[4]
    assert abs(mean - ref[2]) < 1.5 * ref[3] and abs(var - ref[4]) < 1.5 * ref[5], (mean, var, ref[2:6])
