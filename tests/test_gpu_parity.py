"""-m gpu parity tests: the CUDA path (through the C-ABI) against the oracle and the
golden vectors recorded from the unmodified reference.

Tolerances (BASELINE.json north_star): deterministic mode -- accept/reject sequence
identical, chi-squared within 1e-6 relative (fp64); production mode -- distribution
moments within the reference's own repetition-to-repetition spread."""
import os

import numpy as np
import pytest

import common
import gpu_common
import oracle as orc

pytestmark = pytest.mark.gpu

FP64, FP32 = 0, 1
REL_CHI = 1e-6            # north_star: final chi-squared within 1e-6 relative


@pytest.mark.parametrize("name", sorted(orc.MODEL_SPECS))
@pytest.mark.parametrize("prec", [FP64, FP32])
def test_eval_matches_oracle_and_reference(name, prec):
    g = np.load(os.path.join(common.GOLDEN, "eval_models.npz"))
    spec = orc.MODEL_SPECS[name][1]
    values = dict(zip([p[0] for p in spec], g[name + "_pvalues"]))
    om = orc.Model(name, active=[str(a) for a in g[name + "_active"]], values=values)
    q = g["q"]
    eng = gpu_common.engine_for(om, [(q, np.ones_like(q), np.ones_like(q))])
    it, v, w, s = eng.eval(g[name + "_pset"], 0.6666666, precision=prec)
    eng.close()
    ref = g[name + "_it"]                     # from the unmodified reference
    oit, ov, ow, os_ = orc.eval_contribs(om, q, g[name + "_pset"], 0.6666666)
    np.testing.assert_allclose(v, g[name + "_v"], rtol=1e-12)
    np.testing.assert_allclose(w, g[name + "_w"], rtol=1e-12)
    np.testing.assert_allclose(s, g[name + "_s"], rtol=1e-12)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    if prec == FP64:
        tol = 1e-7 if name == "Kholodenko" else 1e-9
        np.testing.assert_allclose(it, ref, rtol=tol, atol=1e-13 * scale.max())
        np.testing.assert_allclose(it, oit, rtol=tol, atol=1e-13 * scale.max())
    else:
        # fp32 form factors: relative to the curve's envelope (minima of F^2 are ill-conditioned)
        env = np.maximum.accumulate(np.abs(ref[:, ::-1]), axis=1)[:, ::-1]
        err = np.abs(it - ref) / np.maximum(env, 1e-30)
        assert err.max() < 5e-4, err.max()
        assert np.median(err) < 2e-5


@pytest.mark.parametrize("prec", [FP64, FP32])
def test_kholodenko_series_integral_matches_oracle_at_high_q(prec):
    """kho_p0sq_series (Gauss-Legendre near the origin + closed-form exponential series) against
    the oracle's adaptive Gauss-Kronrod integral (kholodenko.py:31-38) on the config-4 q range
    (0.01 - 10 1/nm: q lK/3 up to ~170, ~1000 oscillations of the integrand)."""
    NM = 1e-9
    om = orc.Model("Kholodenko", active=["radius", "lenKuhn", "lenContour"])
    q = np.logspace(np.log10(0.01), np.log10(10.0), 40) * 1e9
    rng = np.random.default_rng(5)
    pset = np.stack([rng.uniform(1, 5, 6) * NM, rng.uniform(10, 50, 6) * NM,
                     rng.uniform(100, 1000, 6) * NM], 1)
    pset[0] = [1 * NM, 10 * NM, 1000 * NM]
    pset[1] = [5 * NM, 50 * NM, 100 * NM]
    eng = gpu_common.engine_for(om, [(q, np.ones_like(q), np.ones_like(q))])
    it, v, w, s = eng.eval(pset, 0.6666666, precision=prec)
    eng.close()
    oit, ov, ow, os_ = orc.eval_contribs(om, q, pset, 0.6666666)
    # relative to the running envelope: the cross-section factor 2 J1(qR)/(qR) has zeros
    env = np.maximum.accumulate(np.abs(oit[:, ::-1]), axis=1)[:, ::-1]
    err = np.abs(it - oit) / env
    # fp32 mode: q, R, lK enter as fp32 (6e-8 relative), amplified by q R <= 50 near the J1 lobes
    assert err.max() < (1e-8 if prec == FP64 else 1e-4), err.max()


CHAINS = [c for c in common.CHAIN_CASES if c != "sphere_posbg"]


@pytest.mark.parametrize("name", CHAINS)
@pytest.mark.parametrize("spec_max,cluster", [(1, 1), (32, 1), (32, 2), (8, 4), (32, 8)])
def test_deterministic_chain_matches_reference(name, spec_max, cluster):
    """Uploaded reference proposal stream, fp64: sequential (spec_max=1) and speculative
    batches, q axis on one CTA or split over a 2/4/8-CTA cluster -- all must reproduce the
    reference's accept/reject sequence."""
    g = common.load_chain(name)
    out = gpu_common.run_golden_chain(g, FP64, spec_max, cluster=cluster)
    assert int(out["num_iter"][0, 0]) == g["num_iter"]
    assert int(out["num_moves"][0, 0]) == g["num_moves"]
    assert np.array_equal(out["accept"], g["accept"])        # accept/reject sequence
    assert abs(out["conval0"][0, 0] / g["conval0"] - 1) < REL_CHI
    assert abs(out["conval"][0, 0] / g["conval"] - 1) < REL_CHI
    np.testing.assert_allclose(out["contribs"][0, :, :, 0], g["rset"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(out["fit_meas"][0][:, 0], g["fit_meas"], rtol=1e-6)
    assert abs(out["scaling"][0, 0] / g["scaling"] - 1) < 1e-6
    chi = g["step_chi"]
    mine = out["chi"] if chi.size == out["chi"].size else out["chi"][out["accept"] == 1]
    np.testing.assert_allclose(mine, chi, rtol=REL_CHI)


def test_positive_background_chain_matches_oracle():
    # the reference's LM does not converge on the |b| residual (maxfev warning when the
    # golden vector was recorded); parity for this flag is against the oracle's
    # constrained closed form
    g = common.load_chain("sphere_posbg")
    om, st = common.oracle_model(g), common.oracle_settings(g)
    o = orc.chain(om, st, g["q"], g["I"], g["sigma"], g["stream"], trace=True)
    out = gpu_common.run_golden_chain(g, FP64, 32, pad_steps=0)
    assert abs(out["conval0"][0, 0] / g["conval0"] - 1) < REL_CHI   # initial fit == reference
    assert np.array_equal(out["accept"], o["accept"])
    assert abs(out["conval"][0, 0] / o["conval"] - 1) < REL_CHI
    assert out["background"][0, 0] >= 0.


@pytest.mark.parametrize("name,prec", [("sphere_small", FP64), ("cyl_small", FP64),
                                       ("gauss_small", FP64)])
def test_philox_chain_matches_oracle(name, prec):
    """In-kernel Philox proposals: the oracle replays the same counter-based stream."""
    from mcsas_b200 import Engine
    g = common.load_chain(name)
    om = common.oracle_model(g)
    seed, R = 0x1234ABCD5678EF01, 3
    eng = gpu_common.engine_for(om, [(g["q"], g["I"], g["sigma"])])
    cfg = Engine.make_config(n_contribs=g["N"], n_reps=R, conv_crit=float(g["conv_crit"]),
                             max_iter=g["max_iter"], max_retries=-1, comp_exp=float(g["comp_exp"]),
                             precision=prec, seed=seed, record_trace=True, first_chain=5)
    eng.run(cfg)
    out = eng.fetch()
    st = common.oracle_settings(g)
    for r in range(R):
        stream = orc.philox_stream(seed, 5 + r, g["N"], om, False, g["max_iter"])
        o = orc.chain(om, st, g["q"], g["I"], g["sigma"], stream, trace=True)
        acc, chi = eng.fetch_trace(r, g["max_iter"])
        assert int(out["num_iter"][0, r]) == o["num_iter"]
        assert np.array_equal(acc[:o["num_iter"]], o["accept"])
        assert abs(out["conval"][0, r] / o["conval"] - 1) < REL_CHI
        np.testing.assert_allclose(out["contribs"][0, :, :, r], o["rset"], rtol=1e-12)
    eng.close()


@pytest.mark.parametrize("tag", ["cyl", "kho"])
def test_long_chain_at_full_size_matches_oracle(tag):
    """BASELINE configs[2] (CylindersIsotropic, radius + aspect active, 500 q, the reference's 100-node
    rule, 100 000 steps) and configs[3] (Kholodenko worm, 3 active parameters, 500 q; as many steps as the
    oracle's adaptive integral manages in half an hour of CPU) at the FULL problem size: one chain of the
    oracle on the kernels' Philox stream (tests/golden/long_<tag>.npz, oracle/gen_long_chains.py).
    The fp64 kernel reproduces its accept/reject sequence and the running chi-squared (1e-6); the fp32
    production kernel follows it.  Neither configuration reaches chi2 <= 1 within maxIterations in the
    oracle either (cylinder: 6248 -> 1.016 after 1e5 steps): the non-convergence reported for these
    configurations is the reference algorithm's, not the CUDA path's."""
    from mcsas_b200 import Engine
    path = os.path.join(common.GOLDEN, "long_%s.npz" % tag)
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    g = np.load(path)
    om = orc.Model(str(g["model"]), active=[str(a) for a in g["active"]],
                   ranges={str(a): (float(lo), float(hi)) for a, lo, hi in zip(g["active"], g["lo"], g["hi"])})
    n = int(g["num_iter"])
    acc_ref = np.unpackbits(g["accept"])[:n]
    assert g["conval"] > 1.0                                  # the oracle does not converge here
    outs = {}
    for prec in (FP64, FP32):
        eng = gpu_common.engine_for(om, [(g["q"], g["I"], g["sigma"])])
        cfg = Engine.make_config(n_contribs=int(g["N"]), n_reps=1, conv_crit=1.0, max_iter=int(g["max_iter"]),
                                 max_retries=-1, precision=prec, seed=int(g["seed"]), record_trace=True,
                                 first_chain=int(g["chain"]))
        eng.run(cfg)
        out = eng.fetch()
        acc, chi = eng.fetch_trace(0, n)
        eng.close()
        run = np.empty(n)
        cur = out["conval0"][0, 0]
        for t in range(n):
            if acc[t]:
                cur = chi[t]
            run[t] = cur
        outs[prec] = (out, acc[:n], run)
    out, acc, run = outs[FP64]
    assert int(out["num_iter"][0, 0]) == n and int(out["num_moves"][0, 0]) == int(g["num_moves"])
    assert np.array_equal(acc, acc_ref)                       # every accept/reject decision
    np.testing.assert_allclose(run[99::100], g["run_chi_every_100"], rtol=REL_CHI)
    assert abs(out["conval"][0, 0] / g["conval"] - 1) < REL_CHI
    np.testing.assert_allclose(out["contribs"][0, :, :, 0], g["rset"], rtol=1e-12)
    out32, acc32, run32 = outs[FP32]
    differ = np.nonzero(acc32 != acc_ref)[0]
    first = int(differ[0]) if differ.size else n
    assert first >= min(n, 2000), first                       # marginal flips appear late, if at all
    part = run32[99:first:100]
    np.testing.assert_allclose(part, g["run_chi_every_100"][:part.size], rtol=1e-3)
    assert out32["conval"][0, 0] > 1.0 and abs(np.log(out32["conval"][0, 0] / g["conval"])) < 0.05


def test_fractions_match_oracle():
    from mcsas_b200 import Engine
    g = common.load_chain("sphere_small")
    om, st = common.oracle_model(g), common.oracle_settings(g)
    eng = gpu_common.engine_for(om, [(g["q"], g["I"], g["sigma"])])
    cfg = Engine.make_config(n_contribs=g["N"], n_reps=1, comp_exp=float(g["comp_exp"]), precision=FP64)
    contribs = g["rset"].reshape(1, g["N"], 1, 1)
    f = eng.fractions(cfg, contribs)
    eng.close()
    o = orc.fractions(om, st, g["q"], g["I"], g["sigma"], g["rset"])
    for k in ("vol", "num", "volsqr", "surf", "min_vol", "min_num", "min_volsqr", "min_surf"):
        np.testing.assert_allclose(f[k][0, :, 0], o[k], rtol=1e-8, err_msg=k)
    np.testing.assert_allclose(f["totals"][0, :, 0], o["totals"], rtol=1e-8)
    np.testing.assert_allclose(f["sc"][0, :, 0], o["sc"], rtol=1e-6)


def test_production_sphere_statistics():
    """fp32 + Philox production mode on config 1 (sasfit_sphere-10-1 data, radius 1-100 nm,
    300 contributions): every repetition converges below chi2 = 1 and the volume-weighted
    mean radius agrees with the reference's 10-repetition result 9.7455e-9 +- 3.06e-11 m
    (measured with the unmodified reference) within 4 of its standard deviations."""
    from mcsas_b200 import Engine
    g = common.load_chain("sphere_cfg1")
    om = common.oracle_model(g)
    eng = gpu_common.engine_for(om, [(g["q"], g["I"], g["sigma"])])
    R = 20
    cfg = Engine.make_config(n_contribs=300, n_reps=R, conv_crit=1.0, max_iter=100000,
                             precision=FP32, seed=2024)
    eng.run(cfg)
    out = eng.fetch()
    assert (out["conval"] <= 1.0).all(), out["conval"]
    assert (out["attempts"] == 1).all()
    nit = out["num_iter"][0]
    assert 10000 < nit.mean() < 80000          # reference: 30 282 iterations/rep on average
    f = eng.fractions(Engine.make_config(n_contribs=300, n_reps=R, precision=FP64), out["contribs"])
    eng.close()
    r = out["contribs"][0, :, 0, :]                     # [N, R]
    vf = f["vol"][0]                                    # [N, R]
    mean_r = (vf * r).sum(0) / vf.sum(0)
    assert abs(mean_r.mean() - 9.7455e-9) < 4 * 3.06e-11, (mean_r.mean(), mean_r.std())
    # fp32 chi2 is what fp64 arithmetic says about the same contributions
    st = common.oracle_settings(g)
    for rep in range(3):
        ito, _, _, _ = orc.eval_contribs(om, g["q"], out["contribs"][0, :, :, rep], float(g["comp_exp"]))
        sc, conval, _ = orc.bgscale(g["I"], g["sigma"], ito.sum(0))
        assert abs(conval / out["conval"][0, rep] - 1) < 1e-3


def test_missing_device_or_bad_args_fail_loudly():
    from mcsas_b200 import Engine, McSASB200Error
    with pytest.raises(McSASB200Error):
        Engine(9999)
    eng = Engine(0)
    with pytest.raises(McSASB200Error):
        eng.run(Engine.make_config())          # no datasets / model
    eng.close()


# all 8 models and every settings variant (findBackground off, positiveBackground, startFromMinimum)
TRACK_CASES = ["sphere_small", "sphere_nobg", "sphere_posbg", "sphere_startmin", "lma_small", "sphcs_small",
               "ellcs_small", "elliso_small", "cyl_small", "gauss_small", "kho_small"]


@pytest.mark.parametrize("name", TRACK_CASES)
@pytest.mark.parametrize("spec_max,cluster", [(1, 1), (32, 1), (64, 1), (32, 2), (8, 4)])
def test_fp32_chain_tracks_fp64_chain(name, spec_max, cluster):
    """Production kernel (fp32 residual-delta sums, Gram batches, cached rows) against the
    fp64 parity kernel on the SAME Philox stream: identical accept/reject sequence over the
    first steps and the chi-squared trace within 1e-4 (fp32 form factors)."""
    from mcsas_b200 import Engine
    g = common.load_chain(name)
    om = common.oracle_model(g)
    # the goldens' lengths are set by the reference's speed (Kholodenko: 40 steps); the Philox
    # chains of both kernels can run longer on the same data
    max_iter = min(max(int(g["max_iter"]), 400), 1500)

    def run(prec, spec, cl):
        eng = gpu_common.engine_for(om, [(g["q"], g["I"], g["sigma"])])
        cfg = Engine.make_config(n_contribs=g["N"], n_reps=2, conv_crit=float(g["conv_crit"]),
                                 max_iter=max_iter, max_retries=-1, comp_exp=float(g["comp_exp"]),
                                 find_background=g["find_background"],
                                 positive_background=g["positive_background"],
                                 start_from_minimum=g["start_from_minimum"],
                                 precision=prec, seed=424242, record_trace=True, spec_max=spec,
                                 cluster_size=cl)
        eng.run(cfg)
        out = eng.fetch()
        tr = [eng.fetch_trace(r, max_iter) for r in range(2)]
        eng.close()
        return out, tr

    ref, rtr = run(FP64, 32, 1)
    out, tr = run(FP32, spec_max, cluster)
    for r in range(2):
        n = min(int(out["num_iter"][0, r]), int(ref["num_iter"][0, r]))
        assert n >= 100
        acc, chi = tr[r][0][:n], tr[r][1][:n]
        racc, rchi = rtr[r][0][:n], rtr[r][1][:n]
        differ = np.nonzero(acc != racc)[0]
        first = int(differ[0]) if differ.size else n
        assert first >= min(n, 300), (first, n)            # marginal flips may appear much later
        # gauss_small: 0.01 % uncertainties amplify the ~3e-7 error of the fp32 form factors
        np.testing.assert_allclose(chi[:first], rchi[:first], rtol=1e-2 if name == "gauss_small" else 1e-4)
        assert abs(out["conval0"][0, r] / ref["conval0"][0, r] - 1) < 1e-5


@pytest.mark.parametrize("name,crit", [("sphere_small", 3000.), ("sphere_small", 800.), ("sphere_small", 150.),
                                       ("sphcs_small", 1000.), ("lma_small", 400.), ("gauss_small", 300.),
                                       ("gauss_small", 20.), ("gauss_small", 1.0), ("cyl_small", 900.)])
@pytest.mark.parametrize("spec_max", [8, 64])
def test_chain_stops_at_the_step_that_reaches_the_criterion(name, crit, spec_max):
    """mcsas.py:354-357 tests `conval > convergenceCriterion` before EVERY step.  With a loose
    criterion the chains converge while moves are still frequent, i.e. inside a Gram batch of the
    production kernel (8 steps, several accepted): the steps behind the move that reaches the
    criterion must not be made.  Same Philox stream in the fp64 parity kernel: numIter, numMoves,
    the accept sequence and the contribution set are equal -- nothing is truncated."""
    from mcsas_b200 import Engine
    g = common.load_chain(name)
    om = common.oracle_model(g)
    R = 8
    outs = []
    for prec in (FP64, FP32):
        eng = gpu_common.engine_for(om, [(g["q"], g["I"], g["sigma"])])
        cfg = Engine.make_config(n_contribs=g["N"], n_reps=R, conv_crit=crit, max_iter=3000, max_retries=-1,
                                 comp_exp=float(g["comp_exp"]), find_background=g["find_background"],
                                 precision=prec, seed=90210, record_trace=True,
                                 spec_max=spec_max if prec == FP32 else 32, cluster_size=1)
        eng.run(cfg)
        out = eng.fetch()
        out["tr"] = [eng.fetch_trace(r, 3000)[0] for r in range(R)]
        eng.close()
        outs.append(out)
    a, b = outs
    assert (a["conval"] <= crit).all() and (a["num_iter"] < 3000).all()       # all converge early
    np.testing.assert_array_equal(b["num_iter"], a["num_iter"])
    np.testing.assert_array_equal(b["num_moves"], a["num_moves"])
    for r in range(R):
        n = int(a["num_iter"][0, r])
        np.testing.assert_array_equal(b["tr"][r][:n], a["tr"][r][:n])
        assert a["tr"][r][n - 1] == 1                      # the last step made is the converging move
    np.testing.assert_array_equal(b["contribs"], a["contribs"])
    # gauss_small: 0.01 % uncertainties amplify the ~3e-7 error of the fp32 form factors
    np.testing.assert_allclose(b["conval"], a["conval"], rtol=1e-2 if name == "gauss_small" else 1e-4)


def test_many_chains_run_in_waves_and_do_not_depend_on_packing():
    """More chains than SMs (3 datasets x 100 repetitions): the launch runs in waves; with a
    fixed layout every chain's result depends only on (seed, global chain id), not on which
    other chains share the launch."""
    from mcsas_b200 import Engine
    g = common.load_chain("sphere_cfg1")
    om = common.oracle_model(g)
    sets = [(g["q"], g["I"] * s, g["sigma"] * s) for s in (1.0, 2.0, 0.5)]
    kw = dict(n_contribs=100, conv_crit=1.0, max_iter=2000, max_retries=-1, precision=FP32,
              seed=77, cluster_size=1)
    eng = gpu_common.engine_for(om, sets)
    eng.run(Engine.make_config(n_reps=100, **kw))
    big = eng.fetch()
    eng.close()
    assert np.isfinite(big["conval"]).all() and (big["num_iter"] == 2000).all()
    eng = gpu_common.engine_for(om, sets[:1])
    eng.run(Engine.make_config(n_reps=100, **kw))
    one = eng.fetch()
    eng.close()
    np.testing.assert_array_equal(big["contribs"][0], one["contribs"][0])
    np.testing.assert_array_equal(big["conval"][0], one["conval"][0])
    # the other datasets are the same curve scaled by 2 and 0.5: chi-squared is scale invariant,
    # so their (different) chains reach the same quality of fit and proportional scalings
    assert abs(np.median(big["conval"][1]) / np.median(big["conval"][0]) - 1) < 0.2
    assert abs(np.median(big["scaling"][1]) / np.median(big["scaling"][0]) - 2.0) < 0.2
    assert abs(np.median(big["scaling"][2]) / np.median(big["scaling"][0]) - 0.5) < 0.05


def test_backup_attempts_do_not_change_results():
    """Scheduler: the next attempt of a late chain is started ahead of time on idle SMs
    (capi.cu: mcsas_wait).  An attempt depends on (seed, chain, attempt number) only, so the
    results must be bit-identical with and without backups -- here every attempt fails
    (unreachable criterion), i.e. every backup is promoted."""
    from mcsas_b200 import Engine
    g = common.load_chain("sphere_cfg1")
    om = common.oracle_model(g)
    kw = dict(n_contribs=60, n_reps=12, conv_crit=1e-6, max_iter=12000, max_retries=2, precision=FP32,
              seed=31)
    outs = []
    for flag in (0, -1):
        eng = gpu_common.engine_for(om, [(g["q"], g["I"], g["sigma"])])
        eng.set_tuning(backups=flag)
        eng.run(Engine.make_config(**kw))
        outs.append(eng.fetch())
        info = eng.scheduler_info()
        eng.close()
        if flag == 0:
            assert info["backups_started"] >= 12 and info["backups_used"] >= 12, info
        else:
            assert info["backups_started"] == 0
    a, b = outs
    assert (a["attempts"] == 3).all() and (b["attempts"] == 3).all()          # maxRetries + 1
    assert (a["num_iter"] == 12000).all()
    for k in ("contribs", "conval", "scaling", "background", "num_moves", "steps_all"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert (a["steps_all"] == 36000).all()


RESULT_KEYS = ("contribs", "conval", "scaling", "background", "num_iter", "num_moves", "steps_all",
               "attempts", "conval0")


def _assert_same_results(a, b, what):
    for k in RESULT_KEYS:
        np.testing.assert_array_equal(a[k], b[k], err_msg="%s: %s" % (what, k))
    for fa, fb in zip(a["fit_meas"], b["fit_meas"]):
        np.testing.assert_array_equal(fa, fb, err_msg="%s: fit_meas" % what)


@pytest.mark.parametrize("nq,N", [(1000, 300), (400, 120), (77, 40)])
def test_results_do_not_depend_on_cluster_size_or_slicing(nq, N):
    """A chain's result is a function of (seed, chain id) only.  The production kernel adds the
    partial sums of a chain's warps in one fixed order whatever the cluster size, and a launch
    that ends mid-chain saves everything the MC loop carries, so: one launch per chain on 1, 2, 4
    or 8 CTAs, step-sliced launches, the automatic scheduler (time-sliced, re-packing the chains on
    clusters of changing size, backup attempts) with different time budgets -- all bit-identical,
    including numIter and the fit curves."""
    from mcsas_b200 import Engine, synth
    q, I, sig = synth.sphere_curve(nq=nq, seed=3)
    kw = dict(n_contribs=N, n_reps=12, conv_crit=1.0, max_iter=20000, max_retries=1, precision=FP32, seed=99)

    def run(tuning=None, **over):
        eng = _sphere_engine([(q, I, sig)])
        if tuning:
            eng.set_tuning(**tuning)
        c = dict(kw)
        c.update(over)
        eng.run(Engine.make_config(**c))
        out = eng.fetch()
        info = eng.run_info()
        eng.close()
        return out, info

    ref, info = run(cluster_size=1, slice_steps=-1)                  # one launch, one CTA per chain
    assert info["launches"] == 1
    for c in (2, 4, 8):
        out, info = run(cluster_size=c, slice_steps=-1)
        assert info["last_cluster"] == c
        _assert_same_results(ref, out, "cluster %d" % c)
    out, info = run(cluster_size=2, slice_steps=700)                 # many launches of 700 steps
    assert info["launches"] > 5
    _assert_same_results(ref, out, "step slices")
    for us in (150, 400, 0):                                         # automatic scheduler, time budgets
        out, info = run(tuning=dict(slice_us=us))
        assert info["launches"] > 1
        _assert_same_results(ref, out, "auto scheduler slice_us=%d" % us)
    # the size of a speculative batch never changes a result: batches of at most 16 / 32 / 64 / 128
    # proposals (128 where the contribution set allows it), on one CTA and on clusters, and the
    # scheduler's own choice (128 from 4 CTAs on) forced on for every chain / off
    for sm, c in ((16, 1), (32, 4), (64, 8), (128, 1), (128, 8)):
        out, info = run(cluster_size=c, slice_steps=-1, spec_max=sm)
        _assert_same_results(ref, out, "spec_max %d on %d CTAs" % (sm, c))
    for bb in (1, -1):
        out, info = run(tuning=dict(big_batch=bb))
        _assert_same_results(ref, out, "auto scheduler big_batch=%d" % bb)
    # the L2 eviction priorities of the contribution-row cache are hints: everything evict_last (-1) or
    # only the first few rows of every chain (1 % of the L2 for the resident chains' rows)
    for pct in (-1, 1):
        out, info = run(tuning=dict(l2_keep_pct=pct))
        _assert_same_results(ref, out, "auto scheduler l2_keep_pct=%d" % pct)
        out, info = run(tuning=dict(l2_keep_pct=pct), cluster_size=2, slice_steps=-1)
        _assert_same_results(ref, out, "2 CTAs l2_keep_pct=%d" % pct)


def test_smeared_production_results_do_not_depend_on_the_schedule_or_the_batch():
    """The fp32 sphere kernel reads the smeared locations from a per-run table ([dataset][profile
    point][q point]; capi.cu: mcsas_run_async): a smeared chain's result is the same bits on 1 / 4 / 8
    CTAs and under the automatic scheduler, and a dataset gives the same result alone as behind another
    dataset of a different length in one upload (table rows, q indices and dataset offsets line up)."""
    from mcsas_b200 import Engine, synth
    qo = np.linspace(0., 4e7, 12)
    w = np.ones(12)
    w[6:] = np.linspace(1., 0., 6)
    w /= 2. * np.trapezoid(w, qo)
    sets = [synth.sphere_curve(nq=300, seed=3), synth.sphere_curve(nq=173, seed=4)]
    kw = dict(n_contribs=60, n_reps=3, conv_crit=1.0, max_iter=3000, max_retries=-1, precision=FP32, seed=21)

    def run(datasets, **over):
        eng = _sphere_engine(datasets)
        eng.set_smearing(qo, w, False)
        c = dict(kw)
        c.update(over)
        eng.run(Engine.make_config(**c))
        out = eng.fetch()
        eng.close()
        return out

    ref = run(sets, cluster_size=1, slice_steps=-1)
    assert (ref["conval"] < ref["conval0"]).all()
    for over in (dict(cluster_size=4, slice_steps=-1), dict(cluster_size=8, slice_steps=-1), dict()):
        _assert_same_results(ref, run(sets, **over), "smeared, %s" % (over,))
    # a 2-D profile with negative offsets: smeared locations below their q point, so the chunk evaluator may not
    # take its series vote from the unsmeared q (ChainArgs::smloc_mono = 0) -- and one without (vote on)
    for lo in (-1e7, 0.):
        qo2 = np.linspace(lo, 2e7, 7)
        w2 = np.full(7, 1. / 7.)

        def run2(**over):
            eng = _sphere_engine(sets)
            eng.set_smearing(qo2, w2, True)
            c = dict(kw)
            c.update(over)
            eng.run(Engine.make_config(**c))
            out = eng.fetch()
            eng.close()
            return out
        ref2 = run2(cluster_size=1, slice_steps=-1)
        assert (ref2["conval"] < ref2["conval0"]).all()
        for over in (dict(cluster_size=4, slice_steps=-1), dict(cluster_size=2, slice_steps=500), dict()):
            _assert_same_results(ref2, run2(**over), "2-D smeared from %g, %s" % (lo, over))
    alone = run(sets[1:2], first_chain=3, cluster_size=1, slice_steps=-1)       # same global chain ids
    for k in RESULT_KEYS:
        np.testing.assert_array_equal(alone[k][0], ref[k][1], err_msg="second dataset alone: %s" % k)
    np.testing.assert_array_equal(alone["fit_meas"][0], ref["fit_meas"][1])
    # and smearing is really on: the unsmeared fit of the same chains differs
    eng = _sphere_engine(sets)
    eng.run(Engine.make_config(cluster_size=1, slice_steps=-1, **kw))
    plain = eng.fetch()
    eng.close()
    assert not np.array_equal(plain["conval"], ref["conval"])


@pytest.mark.parametrize("qpt", [4, 2, 1])
def test_production_kernel_layouts_of_1_2_4_q_points_per_thread(qpt):
    """mcsas_tuning.qpt forces the fp32 kernel's q points per thread (4 in production: packed pairs;
    2: one pair; 1: the scalar sums).  Every layout fits (chi-squared falls by more than 100x), and the
    chi-squared it reports is what the fp64 evaluation says about the contribution set it returns."""
    from mcsas_b200 import Engine, synth
    q, I, sig = synth.sphere_curve(nq=200, seed=5)
    eng = _sphere_engine([(q, I, sig)])
    eng.set_tuning(qpt=qpt)
    eng.run(Engine.make_config(n_contribs=60, n_reps=4, conv_crit=1.0, max_iter=8000, max_retries=-1,
                               precision=FP32, seed=17))
    out = eng.fetch()
    assert (out["conval"][0] < 1e-2 * out["conval0"][0]).all(), (out["conval"], out["conval0"])
    for r in range(4):
        it, _, _, _ = eng.eval(out["contribs"][0, :, :, r], 0.6666666, 0, FP64)   # [N, nq], fp64
        m = it.sum(axis=0)
        wgt = 1. / sig ** 2                               # backgroundscalingfit.py:112-139 in closed form
        A = np.array([[(wgt * m * m).sum(), (wgt * m).sum()], [(wgt * m).sum(), wgt.sum()]])
        a, b = np.linalg.solve(A, np.array([(wgt * m * I).sum(), (wgt * I).sum()]))
        chi = float((wgt * (I - a * m - b) ** 2).sum() / q.size)
        assert abs(chi / out["conval"][0, r] - 1) < 1e-3, (chi, out["conval"][0, r])
    eng.close()


def _sphere_engine(sets, lo=1e-9, hi=100e-9):
    from mcsas_b200 import Engine
    eng = Engine(0)
    eng.upload_datasets(sets)
    eng.set_model(0, [10e-9, 1e14], [0], [lo], [hi], [0])
    return eng


@pytest.mark.parametrize("prec", [FP64, FP32])
def test_ragged_datasets_in_one_launch(prec):
    """Datasets of different length share a launch (nq 100, 37, 64): every chain uses its own
    nq; results equal those of the dataset run alone."""
    from mcsas_b200 import Engine, synth
    sets = [synth.sphere_curve(nq=n, seed=i) for i, n in enumerate((100, 37, 64))]
    kw = dict(n_contribs=40, n_reps=3, max_iter=600, max_retries=-1, precision=prec, seed=5, cluster_size=1)
    eng = _sphere_engine(sets)
    eng.run(Engine.make_config(**kw))
    out = eng.fetch()
    eng.close()
    assert [f.shape for f in out["fit_meas"]] == [(100, 3), (37, 3), (64, 3)]
    assert np.isfinite(out["conval"]).all() and (out["conval"] < out["conval0"]).all()
    eng = _sphere_engine(sets[1:2])
    eng.run(Engine.make_config(first_chain=3, **kw))          # same global chain ids as in the batch
    one = eng.fetch()
    eng.close()
    np.testing.assert_array_equal(out["contribs"][1], one["contribs"][0])
    np.testing.assert_allclose(out["fit_meas"][1], one["fit_meas"][0], rtol=1e-12)


def test_maximum_nq_and_tiny_inputs():
    """nq = MCSAS_MAX_NQ (4096: the q axis must be split over a cluster), nq = 3 and a single
    contribution (mcsas.py:354: no MC step is made with one contribution)."""
    from mcsas_b200 import Engine, McSASB200Error, synth
    q, I, s = synth.sphere_curve(nq=4096)
    eng = _sphere_engine([(q, I, s)])
    eng.run(Engine.make_config(n_contribs=300, n_reps=2, max_iter=3000, max_retries=-1, precision=FP32, seed=1))
    out = eng.fetch()
    assert np.isfinite(out["conval"]).all() and (out["conval"] < out["conval0"]).all()
    assert out["fit_meas"][0].shape == (4096, 2)
    eng.run(Engine.make_config(n_contribs=300, n_reps=2, max_iter=3000, max_retries=-1, precision=FP64, seed=1))
    ref = eng.fetch()                                      # the fp64 parity kernel on the same Philox streams
    np.testing.assert_allclose(out["conval0"], ref["conval0"], rtol=1e-5)
    np.testing.assert_allclose(out["conval"], ref["conval"], rtol=0.05)
    with pytest.raises(McSASB200Error):
        eng.upload_datasets([(np.ones(4097), np.ones(4097), np.ones(4097))])
    with pytest.raises(McSASB200Error):
        eng.upload_datasets([(np.array([1e8, np.nan]), np.ones(2), np.ones(2))])
    eng.close()
    q, I, s = synth.sphere_curve(nq=3)
    for prec in (FP64, FP32):
        eng = _sphere_engine([(q, I, s)])
        eng.run(Engine.make_config(n_contribs=5, n_reps=2, max_iter=200, max_retries=-1, precision=prec, seed=2))
        out = eng.fetch()
        assert np.isfinite(out["conval"]).all()
        eng.run(Engine.make_config(n_contribs=1, n_reps=2, max_iter=200, max_retries=-1, precision=prec, seed=2))
        out = eng.fetch()
        assert (out["num_iter"] == 0).all() and np.isfinite(out["conval"]).all()
        eng.close()


# ---- smearing (sasmodel.py:56-73) ------------------------------------------
@pytest.mark.parametrize("tag", ["sphere_slit", "sphere_2d", "lma_slit", "sphere_gauss"])
@pytest.mark.parametrize("prec", [FP64, FP32])
def test_smeared_eval_matches_reference(tag, prec):
    g = np.load(os.path.join(common.GOLDEN, "eval_smeared.npz"))
    name = str(g[tag + "_model"])
    spec = orc.MODEL_SPECS[name][1]
    om = orc.Model(name, active=["radius"], values=dict(zip([p[0] for p in spec], g[tag + "_pvalues"])))
    q = g["q"]
    eng = gpu_common.engine_for(om, [(q, np.ones_like(q), np.ones_like(q))],
                                smear=(g[tag + "_off"], g[tag + "_w"], bool(g[tag + "_two_d"])))
    it, v, w, s = eng.eval(g[tag + "_pset"], 0.6666666, precision=prec)
    eng.set_smearing(None)
    it_plain, _, _, _ = eng.eval(g[tag + "_pset"], 0.6666666, precision=prec)
    eng.close()
    ref = g[tag + "_it"]
    if prec == FP64:
        np.testing.assert_allclose(it, ref, rtol=1e-9)
    else:
        env = np.maximum.accumulate(np.abs(ref[:, ::-1]), axis=1)[:, ::-1]
        assert (np.abs(it - ref) / env).max() < 5e-4
    assert np.abs(it_plain / ref - 1).max() > 1e-2          # smearing was really on / is really off


@pytest.mark.parametrize("name", common.SMEAR_CHAINS)
@pytest.mark.parametrize("spec_max,cluster", [(1, 1), (32, 1), (32, 2)])
def test_deterministic_smeared_chain_matches_reference(name, spec_max, cluster):
    g = common.load_chain(name)
    out = gpu_common.run_golden_chain(g, FP64, spec_max, cluster=cluster)
    assert int(out["num_iter"][0, 0]) == g["num_iter"]
    assert int(out["num_moves"][0, 0]) == g["num_moves"]
    assert np.array_equal(out["accept"], g["accept"])
    assert abs(out["conval0"][0, 0] / g["conval0"] - 1) < REL_CHI
    assert abs(out["conval"][0, 0] / g["conval"] - 1) < REL_CHI
    np.testing.assert_allclose(out["contribs"][0, :, :, 0], g["rset"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(out["fit_meas"][0][:, 0], g["fit_meas"], rtol=1e-6)


def test_smeared_production_chain_tracks_fp64_chain():
    """fp32 production kernel with a smearing profile (the sphere's q fold is off, rows cached)
    against the fp64 kernel on the same Philox stream."""
    from mcsas_b200 import Engine
    g = common.load_chain("sphere_smear")
    om = common.oracle_model(g)
    outs = []
    for prec in (FP64, FP32):
        eng = gpu_common.engine_for(om, [(g["q"], g["I"], g["sigma"])], smear=common.smear_of(g))
        eng.run(Engine.make_config(n_contribs=g["N"], n_reps=2, conv_crit=1.0, max_iter=800, max_retries=-1,
                                   comp_exp=float(g["comp_exp"]), precision=prec, seed=11, record_trace=True))
        out = eng.fetch()
        out["tr"] = [eng.fetch_trace(r, 800) for r in range(2)]
        eng.close()
        outs.append(out)
    a, b = outs
    for r in range(2):
        n = 800
        differ = np.nonzero(a["tr"][r][0][:n] != b["tr"][r][0][:n])[0]
        first = int(differ[0]) if differ.size else n
        assert first >= 300, first
        np.testing.assert_allclose(b["tr"][r][1][:first], a["tr"][r][1][:first], rtol=2e-4)
