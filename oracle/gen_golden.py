# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the
UNMODIFIED reference (McSAS 1.3.1 at /root/reference, imported through
oracle/refharness.py) in the build container.  The reference cannot travel to
the GPU box, the vectors can.  Re-run:  python oracle/gen_golden.py [case ...]

Vectors per chain case: data (q, I, sigma as the reference hands them to the
hot path), generator-output stream, per-step chi-squared
and accept trace, final contributions, scaling/background, fitted intensity.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refharness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
NM = 1e-9

# (model, {param: value}, active [(name, lo, hi)], truth pset generator)
CASES = {
    "sphere_cfg1": dict(model="Sphere", file="sasfit_sphere-10-1.dat",
                        active=[("radius", 1 * NM, 100 * NM)], N=300, max_iter=100000,
                        seed=12345, store_stream=False),
    "sphere_small": dict(model="Sphere", active=[("radius", 1 * NM, 100 * NM)],
                         nq=64, N=40, max_iter=3000, seed=7),
    "sphere_nobg": dict(model="Sphere", active=[("radius", 1 * NM, 100 * NM)],
                        nq=48, N=30, max_iter=1500, seed=8, find_background=False),
    "sphere_posbg": dict(model="Sphere", active=[("radius", 1 * NM, 100 * NM)],
                         nq=48, N=30, max_iter=1500, seed=9, positive_background=True,
                         bg=-50.0),
    "sphere_startmin": dict(model="Sphere", active=[("radius", 1 * NM, 100 * NM)],
                            nq=48, N=30, max_iter=1500, seed=10, start_from_minimum=True),
    "sphere_smear": dict(model="Sphere", active=[("radius", 1 * NM, 100 * NM)],
                         nq=40, N=30, max_iter=1200, seed=21,
                         smear=dict(umbra=0.02e9, penumbra=0.05e9, two_d=False)),
    "lma_smear2d": dict(model="LMADenseSphere", active=[("radius", 1 * NM, 50 * NM)],
                        nq=40, N=24, max_iter=800, seed=22,
                        smear=dict(umbra=0.004e9, penumbra=0.012e9, two_d=True)),
    "lma_small": dict(model="LMADenseSphere", active=[("radius", 1 * NM, 50 * NM)],
                      nq=48, N=30, max_iter=1500, seed=11),
    "sphcs_small": dict(model="SphericalCoreShell", active=[("radius", 1 * NM, 50 * NM)],
                        values={"t": 2 * NM}, nq=48, N=30, max_iter=1500, seed=12),
    "ellcs_small": dict(model="EllipsoidalCoreShell", active=[("a", 1 * NM, 50 * NM)],
                        values={"b": 8 * NM, "t": 2 * NM}, nq=32, N=20, max_iter=600, seed=13),
    "elliso_small": dict(model="EllipsoidsIsotropic", active=[("a", 1 * NM, 50 * NM)],
                         values={"aspect": 3.0}, nq=32, N=20, max_iter=600, seed=14),
    "cyl_small": dict(model="CylindersIsotropic",
                      active=[("radius", 1 * NM, 50 * NM), ("aspect", 1.0, 20.0)],
                      nq=32, N=20, max_iter=600, seed=15),
    "gauss_small": dict(model="GaussianChain", active=[("rg", 1 * NM, 100 * NM)],
                        nq=48, N=30, max_iter=1500, seed=16),
    "kho_small": dict(model="Kholodenko",
                      active=[("radius", 1 * NM, 5 * NM), ("lenKuhn", 10 * NM, 50 * NM),
                              ("lenContour", 100 * NM, 1000 * NM)],
                      nq=16, N=8, max_iter=40, seed=17),
}


def make_model(case):
    m = rh.model_class(case["model"])()
    for k, v in case.get("values", {}).items():
        getattr(m, k).setValue(v)
    for p in m.params():
        if hasattr(p, "setActive"):
            p.setActive(False)
    for name, lo, hi in case["active"]:
        p = getattr(m, name)
        p.setActive(True)
        p.setActiveRange((lo, hi))
    return m


def synth_data(case, model):
    """Synthetic curve from the reference model itself: a handful of truth
    contributions mid-range, flat background, 1 % uncertainty + noise."""
    rng = np.random.default_rng(1000 + case["seed"])
    nq = case["nq"]
    q_nm = np.logspace(np.log10(0.03), np.log10(3.0), nq)
    if case["model"] == "Kholodenko":
        q_nm = np.logspace(np.log10(0.01), np.log10(5.0), nq)
    dummy = rh.make_data(q_nm, np.ones(nq), 0.01 * np.ones(nq))
    nt = 12 if case["model"] != "Kholodenko" else 3
    pset = np.zeros((nt, len(case["active"])))
    for i, (name, lo, hi) in enumerate(case["active"]):
        mid = np.sqrt(lo * hi) if lo > 0 else 0.5 * (lo + hi)
        pset[:, i] = np.clip(mid * np.exp(0.25 * rng.standard_normal(nt)), lo, hi)
    md = model.calc(dummy, pset, 0.6666666)
    I = md.cumInt / md.cumInt[0] * 1e6 + case.get("bg", 1.0)
    sig = 0.01 * np.abs(I)
    I = I + rng.standard_normal(nq) * sig
    return rh.make_data(q_nm, I, sig)


def apply_smearing(data, smear):
    """Enable the reference's TrapezoidSmearing on a SASData (sasconfig.py:76-149, 308-339)."""
    sm = data.config.smearing
    sm.doSmear.setValue(True)
    sm.twoDColl.setValue(bool(smear.get("two_d", False)))
    sm.umbra.setValue(smear["umbra"])
    sm.penumbra.setValue(smear["penumbra"])
    assert sm.inputValid()
    data.setConfig(data.config)
    assert data.locs.ndim == 2
    return sm


def run_chain_case(name, case):
    rh.install()
    model = make_model(case)
    if "file" in case:
        data = rh.load_data(os.path.join(rh.REFERENCE_TESTDATA, case["file"]))
    else:
        data = synth_data(case, model)
    smear_out = {}
    if "smear" in case:
        sm = apply_smearing(data, case["smear"])
        qo, w = sm.prepared
        smear_out = dict(smear_off=np.asarray(qo, float), smear_w=np.asarray(w, float),
                         smear_two_d=bool(sm.twoDColl()))
    st = dict(numContribs=case["N"], maxIterations=case["max_iter"])
    for k in ("findBackground", "positiveBackground", "startFromMinimum"):
        pyk = {"findBackground": "find_background", "positiveBackground": "positive_background",
               "startFromMinimum": "start_from_minimum"}[k]
        if pyk in case:
            st[k] = case[pyk]
    algo = rh.make_algo(model, data, **st)
    np.random.seed(case["seed"])
    t0 = time.time()
    with rh.RecordingGenerators() as gen, rh.RecordingFit() as fit:
        rset, fit_meas, conval, details = algo.mcFit(case["N"], True, True, nRun=0)
    el = time.time() - t0
    stream = gen.flat()
    chis = np.array(fit.conval)
    nit = details["numIterations"]
    assert chis.size == nit + 3, (chis.size, nit)
    step_chi = chis[2:2 + nit]
    # accept trace: strict improvement over the running conval (mcsas.py:379)
    accept = np.zeros(nit, dtype=np.uint8)
    cur = chis[1]
    for t in range(nit):
        if step_chi[t] < cur:
            accept[t] = 1
            cur = step_chi[t]
    assert accept.sum() == details["numMoves"], (accept.sum(), details["numMoves"])
    # parameter vector in the reference's `parameters` order
    pvals = np.array([float(p()) for p in model.params()])
    out = dict(
        model=case["model"], pnames=np.array([p.name() for p in model.params()]),
        pvalues=pvals, active=np.array([a[0] for a in case["active"]]),
        lo=np.array([a[1] for a in case["active"]]), hi=np.array([a[2] for a in case["active"]]),
        q=np.asarray(data.q, dtype=np.float64).flatten(),
        I=np.asarray(data.f.binnedData, dtype=np.float64).flatten(),
        sigma=np.asarray(data.f.binnedDataU, dtype=np.float64).flatten(),
        q_limit_hi=float(data.x0.limit[1]),
        N=case["N"], max_iter=case["max_iter"], seed=case["seed"],
        conv_crit=float(algo.convergenceCriterion()), comp_exp=float(algo.compensationExponent()),
        find_background=bool(algo.findBackground.value()),
        positive_background=bool(algo.positiveBackground.value()),
        start_from_minimum=bool(algo.startFromMinimum()),
        conval0=chis[1], conval=float(conval), num_iter=nit, num_moves=details["numMoves"],
        scaling=float(details["scaling"]), background=float(details["background"]),
        rset=np.asarray(rset, dtype=np.float64), fit_meas=np.asarray(fit_meas).flatten(),
        accept=np.packbits(accept), step_chi=step_chi if nit <= 5000 else step_chi[accept == 1],
        stream_len=stream.size, stream_sum=float(stream.sum()),
    )
    if case.get("store_stream", True):
        out["stream"] = stream
    out.update(smear_out)
    np.savez_compressed(os.path.join(OUT, "chain_%s.npz" % name), **out)
    print("chain %-16s nit=%6d moves=%5d chi0=%.6g chi=%.6g  (%.1fs)" % (
        name, nit, details["numMoves"], chis[1], conval, el))


EVAL_CASES = {
    "Sphere": dict(active=["radius"], pset=[[1.3 * NM], [10 * NM], [97 * NM]]),
    "LMADenseSphere": dict(active=["radius"], pset=[[1.3 * NM], [10 * NM], [47 * NM]],
                           values={"volFrac": 0.2}),
    "SphericalCoreShell": dict(active=["radius", "t"],
                               pset=[[1.3 * NM, 0.5 * NM], [10 * NM, 2 * NM], [47 * NM, 30 * NM]]),
    "EllipsoidalCoreShell": dict(active=["a", "b", "t"],
                                 pset=[[1.3 * NM, 4 * NM, 1 * NM], [10 * NM, 15 * NM, 5 * NM],
                                       [47 * NM, 20 * NM, 50 * NM]]),
    "EllipsoidsIsotropic": dict(active=["a", "aspect"],
                                pset=[[1.3 * NM, 0.3], [10 * NM, 10.], [47 * NM, 2.5]]),
    "CylindersIsotropic": dict(active=["radius", "aspect"],
                               pset=[[1.3 * NM, 0.5], [10 * NM, 10.], [47 * NM, 3.]]),
    "GaussianChain": dict(active=["rg"], pset=[[1.3 * NM], [10 * NM], [97 * NM]]),
    "Kholodenko": dict(active=["radius", "lenKuhn", "lenContour"],
                       pset=[[1 * NM, 10 * NM, 1000 * NM], [3 * NM, 45 * NM, 150 * NM],
                             [4.9 * NM, 20 * NM, 500 * NM]]),
}


def run_eval_cases():
    """ScatteringModel.calc on single contributions (scatteringmodel.py:79-105)."""
    rh.install()
    out = {}
    q_nm = np.logspace(np.log10(0.01), np.log10(5.0), 60)
    data = rh.make_data(q_nm, np.ones(60), 0.01 * np.ones(60))
    out["q"] = np.asarray(data.q).flatten()
    for name, c in EVAL_CASES.items():
        m = rh.model_class(name)()
        for k, v in c.get("values", {}).items():
            getattr(m, k).setValue(v)
        for p in m.params():
            if hasattr(p, "setActive"):
                p.setActive(False)
        for a in c["active"]:
            getattr(m, a).setActive(True)
        pset = np.array(c["pset"], dtype=np.float64)
        its, vs, ws, ss = [], [], [], []
        for row in pset:
            md = m.calc(data, row.reshape(1, -1), 0.6666666)
            its.append(md.cumInt.copy()); vs.append(md.vset[0]); ws.append(md.wset[0]); ss.append(md.sset[0])
        out[name + "_pset"] = pset
        out[name + "_active"] = np.array(c["active"])
        out[name + "_pvalues"] = np.array([float(p()) for p in m.params()])
        out[name + "_it"] = np.array(its)
        out[name + "_v"] = np.array(vs)
        out[name + "_w"] = np.array(ws)
        out[name + "_s"] = np.array(ss)
        print("eval", name, "ok")
    np.savez_compressed(os.path.join(OUT, "eval_models.npz"), **out)


def run_smear_eval_cases():
    """Smeared ScatteringModel.calc (sasmodel.py:56-73) of the canSmear models for a slit and a
    2-D averaged trapezoidal profile and a Gaussian one."""
    rh.install()
    from mcsas.dataobj.sasconfig import GaussianSmearing
    q_nm = np.logspace(np.log10(0.03), np.log10(3.0), 40)
    out = dict(q=q_nm * 1e9)
    for tag, mname, pset, conf in [
            ("sphere_slit", "Sphere", [[1.3 * NM], [10 * NM], [60 * NM]], dict(umbra=0.02e9, penumbra=0.05e9, two_d=False)),
            ("sphere_2d", "Sphere", [[1.3 * NM], [10 * NM], [60 * NM]], dict(umbra=0.004e9, penumbra=0.012e9, two_d=True)),
            ("lma_slit", "LMADenseSphere", [[1.3 * NM], [10 * NM], [40 * NM]], dict(umbra=0.01e9, penumbra=0.03e9, two_d=False)),
            ("sphere_gauss", "Sphere", [[2 * NM], [25 * NM]], dict(variance=0.01e9, two_d=False))]:
        data = rh.make_data(q_nm, np.ones(q_nm.size), 0.01 * np.ones(q_nm.size))
        if "variance" in conf:
            sm = GaussianSmearing()
            data.config.smearing = sm
            sm.doSmear.setValue(True)
            sm.twoDColl.setValue(conf["two_d"])
            sm.variance.setValue(conf["variance"])
            data.setConfig(data.config)
        else:
            sm = apply_smearing(data, conf)
        model = rh.model_class(mname)()
        for p in model.params():
            if hasattr(p, "setActive"):
                p.setActive(False)
        model.radius.setActive(True)
        md = [model.calc(data, np.array([row]), 0.6666666) for row in pset]
        qo, w = sm.prepared
        out[tag + "_model"] = np.array(mname)
        out[tag + "_pvalues"] = np.array([float(p()) for p in model.params()])
        out[tag + "_pset"] = np.array(pset)
        out[tag + "_it"] = np.array([np.asarray(m.cumInt).flatten() for m in md])
        out[tag + "_off"] = np.asarray(qo, float)
        out[tag + "_w"] = np.asarray(w, float)
        out[tag + "_two_d"] = np.array(bool(sm.twoDColl()))
        print("smear eval", tag, out[tag + "_it"].shape, len(qo))
    np.savez_compressed(os.path.join(OUT, "eval_smeared.npz"), **out)


def run_fit_cases():
    """BackgroundScalingFit.calc (backgroundscalingfit.py:112-139) vectors."""
    rh.install()
    from mcsas.mcsas.backgroundscalingfit import BackgroundScalingFit
    from mcsas.bases.model import SASModelData
    rng = np.random.default_rng(5)
    recs = []
    q_nm = np.logspace(-1.5, 0.5, 50)
    for trial in range(12):
        M = np.exp(-q_nm ** 2 * rng.uniform(5, 50)) * rng.uniform(0.1, 10) + 1e-3
        a, b = rng.uniform(0.5, 50), rng.uniform(-2, 2) * 0.01
        I = a * M + b
        sig = 0.02 * np.abs(I) + 1e-4
        I = I + rng.standard_normal(50) * sig
        if trial == 3:
            sig[5] = 0.0    # sigma == 0 -> 1  (:117)
        data = rh.make_data(q_nm, I, np.maximum(sig, 1e-300))
        # overwrite prepared vectors so that exactly (I, sig) reach the fit
        Iv, sv = I.copy(), sig.copy()

        class D(object):
            pass
        d = D(); d.f = D(); d.f.binnedData = Iv; d.f.binnedDataU = sv.copy()
        md = SASModelData(M.copy(), np.ones(1), np.ones(1), np.ones(1), 1)
        for fb, pb in ((True, False), (False, False), (True, True)):
            B = BackgroundScalingFit(fb, pb, None)
            sc0 = np.array([I.max() / M.max(), I.min()])
            sc, conval, _, agofs = B.calc(d, md, sc0.copy(), ver=1)
            sc, conval, _, agofs = B.calc(d, md, sc.copy())
            recs.append(dict(I=Iv, sigma=sv, M=M, fb=fb, pb=pb, sc=np.array(sc), conval=conval,
                             agofs=agofs))
    out = {}
    for k in recs[0]:
        out[k] = np.array([r[k] for r in recs])
    np.savez_compressed(os.path.join(OUT, "bgfit.npz"), **out)
    print("bgfit ok", len(recs))


def run_sasfit_fixtures():
    """SASfit golden curves shipped with the reference (testdata/sasfit_*.dat;
    thresholds sphere.py:69, gaussianchain.py:84, kholodenko.py:98-103)."""
    out = {}
    for fn in ("sasfit_sphere-2-1.dat", "sasfit_sphere-10-1.dat", "sasfit_sphere-20-1.dat",
               "sasfit_sphere-50-1.dat", "sasfit_sphere-100-1.dat", "sasfit_kho-1-10-1000.dat",
               "sasfit_gauss2-5-1.5-2-1.dat", "sasfit_gauss2-1-100-1-1.dat"):
        rows = []
        with open(os.path.join(rh.REFERENCE_TESTDATA, fn)) as fh:
            for line in fh:
                try:
                    vals = [float(x) for x in line.replace(";", " ").replace(",", " ").split()]
                except ValueError:
                    continue
                if len(vals) >= 2:
                    rows.append(vals[:2])
        a = np.array(rows)
        out[fn.replace(".dat", "").replace("-", "_").replace(".", "p")] = a
        print("fixture", fn, a.shape)
    np.savez_compressed(os.path.join(OUT, "sasfit_curves.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:]
    if not which or "eval" in which:
        run_eval_cases()
    if not which or "bgfit" in which:
        run_fit_cases()
    if not which or "smear" in which:
        run_smear_eval_cases()
    if not which or "sasfit" in which:
        run_sasfit_fixtures()
    for name, case in CASES.items():
        if not which or name in which or "chains" in which:
            run_chain_case(name, case)


def run_histogram_case():
    """McSAS.calc -> analyse + histogram (mcsas.py:149-179, 445-615) on a small synthetic
    sphere case: contribs, the per-repetition fractions handed to Histograms.calc, and the
    resulting Histogram / Moments values (utils/parameter.py:20-122, 420-479)."""
    rh.install()
    from mcsas.utils.parameter import Histogram, Histograms
    case = dict(model="Sphere", active=[("radius", 1 * NM, 100 * NM)], nq=64, N=40, seed=7)
    model = make_model(case)
    data = synth_data(case, model)
    h1 = Histogram(model.radius, 1 * NM, 100 * NM, 20, 'log', 'vol')
    h2 = Histogram(model.radius, 2 * NM, 60 * NM, 12, 'lin', 'num', autoFollow=False)
    model.radius.histograms().append(h1)
    model.radius.histograms().append(h2)
    algo = rh.make_algo(model, data, numContribs=40, numReps=4, maxIterations=3000,
                        convergenceCriterion=40.0)
    captured = {}
    orig = Histograms.calc

    def calc(self, contribs, paramIndex, fractions):
        captured["fractions"] = fractions
        return orig(self, contribs, paramIndex, fractions)
    Histograms.calc = calc
    np.random.seed(21)
    try:
        algo.calc()
    finally:
        Histograms.calc = orig
    res = algo.result[0]
    fr = captured["fractions"]
    out = dict(q=np.asarray(data.q).flatten(), I=np.asarray(data.f.binnedData).flatten(),
               sigma=np.asarray(data.f.binnedDataU).flatten(), contribs=res["contribs"],
               comp_exp=float(algo.compensationExponent()))
    for k in ("vol", "num", "int", "surf"):
        out["frac_" + k] = fr[k][0]
        out["minreq_" + k] = fr[k][1]
    for i, h in enumerate(model.radius.histograms()):
        out["h%d_edges" % i] = h.xLowerEdge
        out["h%d_mean" % i] = h.bins.mean
        out["h%d_std" % i] = h.bins.std
        out["h%d_cdf" % i] = h.cdf.mean
        out["h%d_obs" % i] = h.observability
        out["h%d_moments" % i] = np.array(h.moments.fields)
        out["h%d_cfg" % i] = np.array([h.lower, h.upper, h.binCount])
    np.savez_compressed(os.path.join(OUT, "histogram_sphere.npz"), **out)
    print("histogram case ok; moments", out["h0_moments"][:4])


if __name__ == "__main__" and (not sys.argv[1:] or "hist" in sys.argv[1:]):
    run_histogram_case()


def run_dataprep_cases():
    """Data preparation of the reference: loaddatafile -> SASData ->
    config.updateFuMin() -> setConfig (dataobj.py:204-226, 239-269, 288-345) on the shipped
    test data; raw table + DataConfig values + the vectors mcFit would read."""
    rh.install()
    from mcsas.datafile import loaddatafile
    out = {}

    def case(tag, fn, nbin=None, qlow=None, qhigh=None, maskneg=None, fumin=None):
        d = loaddatafile(os.path.join(rh.REFERENCE_TESTDATA, fn)).getDataObj()
        c = d.config
        if nbin is not None:
            c.nBin.setValue(nbin)
        if fumin is not None:
            c.fuMin.setValue(fumin)
        if qlow is not None:
            c.x0Low.setValue(qlow)
        if qhigh is not None:
            c.x0High.setValue(qhigh)
        if maskneg is not None:
            c.fMaskNeg.setValue(maskneg)
        c.updateFuMin()
        d.setConfig(c)
        out[tag + "_raw"] = np.stack([np.asarray(d.x0.rawData, float), np.asarray(d.f.rawData, float),
                                      np.asarray(d.f.rawDataU, float)], 1)
        out[tag + "_cfg"] = np.array([c.nBin(), c.fuMin(), c.x0Low(), c.x0High(),
                                      float(c.fMaskZero()), float(c.fMaskNeg())])
        binned = d.f.binnedData is not None
        out[tag + "_q"] = np.asarray(d.q, float).ravel()
        out[tag + "_I"] = np.asarray(d.f.binnedData if binned else d.f.sanitized, float).ravel()
        out[tag + "_sigma"] = np.asarray(d.f.binnedDataU if binned else d.f.sanitizedU, float).ravel()
        print("dataprep", tag, out[tag + "_raw"].shape, "->", out[tag + "_q"].shape)

    case("sphere10_default", "sasfit_sphere-10-1.dat")
    case("quick_default", "quickstartdemo1.csv")
    case("quick_nbin50_lim", "quickstartdemo1.csv", nbin=50, qlow=0.05e9, qhigh=1.5e9, maskneg=True)
    case("quick_nobin", "quickstartdemo1.csv", nbin=0, fumin=0.02)
    np.savez_compressed(os.path.join(OUT, "dataprep.npz"), **out)


if __name__ == "__main__" and (not sys.argv[1:] or "dataprep" in sys.argv[1:]):
    run_dataprep_cases()


def run_output_cases():
    """Text layout of the reference's result files: Calculator.writeResults
    (gui/calc.py:128-155) = AsciiFile.writeHeaderLine + appendFile (asciifile.py:15-68) on a
    fit table and on a statistics-like table of pre-formatted cells."""
    import tempfile
    rh.install()
    from mcsas.datafile.asciifile import AsciiFile
    rng = np.random.default_rng(5)
    cols = {"fitX0": np.logspace(7, 9, 7), "dataMean": rng.uniform(1, 1e6, 7),
            "dataStd": rng.uniform(0.01, 10, 7), "fitMeasValMean": rng.uniform(1, 1e6, (1, 7)),
            "fitMeasValStd": rng.uniform(0, 1, (1, 7))}
    names = ('fitX0', 'dataMean', 'dataStd', 'fitMeasValMean', 'fitMeasValStd')
    fn = tempfile.mktemp()
    AsciiFile.writeHeaderLine(fn, names)
    AsciiFile.appendFile(fn, np.vstack([cols[c] for c in names]).T)
    fit_txt = open(fn).read()
    stats = {"lower": [AsciiFile.formatValue(1e-9), AsciiFile.formatValue(5e-9)],
             "upper": [AsciiFile.formatValue(1e-7)] * 2,
             "weighting": [AsciiFile.formatValue("vol"), AsciiFile.formatValue("num")],
             "mean": [AsciiFile.formatValue(9.7455e-9), AsciiFile.formatValue(-1.5)]}
    sn = ("lower", "upper", "weighting", "mean")
    AsciiFile.writeHeaderLine(fn, sn)
    AsciiFile.appendFile(fn, np.vstack([stats[c] for c in sn]).T)
    stats_txt = open(fn).read()
    os.remove(fn)
    np.savez_compressed(os.path.join(OUT, "output_formats.npz"), fit_txt=np.array(fit_txt),
                        stats_txt=np.array(stats_txt), **{"col_" + k: v for k, v in cols.items()})
    print("output formats ok:", len(fit_txt), len(stats_txt))


if __name__ == "__main__" and (not sys.argv[1:] or "output" in sys.argv[1:]):
    run_output_cases()
