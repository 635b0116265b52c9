# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- production-mode statistical goldens from the UNMODIFIED
reference (BASELINE.json north_star: "volume-weighted distribution moments and histograms
must agree with the reference within its own repetition-to-repetition standard deviation").

Every repetition is one process running the reference's own McSAS.analyse with numReps = 1
(mcsas.py:191-285: mcFit + the retry loop) on its own numpy seed; the parent then hands all
contribution sets to the reference's McSAS.histogram (mcsas.py:445-615) and records what
Histogram / Moments (utils/parameter.py:20-122, 420-479) make of them.  Run in the build
container only (/root/reference is not on the GPU box); outputs go to tests/golden/.

  python oracle/gen_stats_golden.py [case ...] [--reps R] [--procs P]
"""
import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
NM = 1e-9


def _curve(case):
    """(q [1/nm], I, sigma) of a case -- the generators of mcsas_b200/synth.py (data synthesis
    only) so that the GPU test can rebuild the identical vectors; they are stored anyway."""
    from mcsas_b200 import synth
    if case["kind"] == "sphere":
        q, I, s = synth.sphere_curve(nq=case["nq"], seed=case.get("data_seed", 0))
    elif case["kind"] == "cylinder":
        # BASELINE configs[2] truth (radius ~ N(5 nm, 1 nm), aspect 5) evaluated by the reference's
        # own CylindersIsotropic; the vectors are stored in the fixture
        import refharness as rh
        rh.install()
        rng = np.random.default_rng(3)
        q_nm = np.logspace(np.log10(0.03), np.log10(3.0), case["nq"])
        m = rh.model_class("CylindersIsotropic")()
        m.radius.setActive(True)
        m.aspect.setActive(True)
        radii = np.clip(rng.normal(5., 1., 40), 0.5, None) * NM
        dummy = rh.make_data(q_nm, np.ones(q_nm.size), 0.01 * np.ones(q_nm.size))
        it = m.calc(dummy, np.stack([radii, np.full_like(radii, 5.0)], 1), 1.0).cumInt
        I = it / it[0] * 1e6 + 1.0
        s = 0.01 * I
        I = I + rng.normal(0., 1., I.size) * s
        return q_nm, I, s
    else:
        raise KeyError(case["kind"])
    return q * 1e-9, I, s


CASES = {
    # BASELINE configs[1]: the headline workload
    "config2": dict(kind="sphere", model="Sphere", nq=1000, N=300, max_iter=100000,
                    active=[("radius", 1 * NM, 100 * NM)], seed0=20260, hist=("log", 50)),
    # the same generator on 200 points (cheap cross-check, also used for N > 1 GPU tests)
    "sphere200": dict(kind="sphere", model="Sphere", nq=200, N=200, max_iter=100000,
                      active=[("radius", 1 * NM, 100 * NM)], seed0=777, hist=("log", 40)),
    # reduced BASELINE configs[2]: CylindersIsotropic, radius + aspect active, 100-node rule
    "config3_reduced": dict(kind="cylinder", model="CylindersIsotropic", nq=60, N=60, max_iter=20000,
                            active=[("radius", 1 * NM, 50 * NM), ("aspect", 1.0, 20.0)],
                            seed0=4242, hist=("log", 20), conv_crit=1.0, show_incomplete=True),
}


def _build(case):
    import refharness as rh
    rh.install()
    model = rh.model_class(case["model"])()
    for p in model.params():
        if hasattr(p, "setActive"):
            p.setActive(False)
    for name, lo, hi in case["active"]:
        p = getattr(model, name)
        p.setActive(True)
        p.setActiveRange((lo, hi))
    q_nm, I, sig = _curve(case)
    data = rh.make_data(q_nm, I, sig, nbin=0)
    algo = rh.make_algo(model, data, numContribs=case["N"], numReps=1,
                        maxIterations=case["max_iter"],
                        convergenceCriterion=case.get("conv_crit", 1.0),
                        showIncomplete=bool(case.get("show_incomplete", False)))
    return rh, model, data, algo


def _one_rep(args):
    name, r = args
    case = CASES[name]
    rh, model, data, algo = _build(case)
    np.random.seed(case["seed0"] + r)
    t0 = time.time()
    attempts = [0]
    orig = algo.mcFit

    def counted(*a, **kw):
        attempts[0] += 1
        return orig(*a, **kw)
    algo.mcFit = counted
    algo.analyse()
    el = time.time() - t0
    if not algo.result:
        return None
    res = algo.result[0]
    return dict(contribs=res["contribs"][:, :, 0], scaling=res["scaling"][0],
                background=res["background"][0], num_iter=res["numIter"],
                fit=res["fitMeasValMean"].ravel(), attempts=attempts[0], elapsed=el)


def run_case(name, reps, procs):
    case = CASES[name]
    t0 = time.time()
    with mp.Pool(procs) as pool:
        outs = pool.map(_one_rep, [(name, r) for r in range(reps)], chunksize=1)
    outs = [o for o in outs if o is not None]
    R = len(outs)
    rh, model, data, algo = _build(case)
    from mcsas.utils.parameter import Histogram
    contribs = np.stack([o["contribs"] for o in outs], axis=2)              # [N, P, R]
    hists = []
    for name_p, lo, hi in case["active"]:
        p = getattr(model, name_p)
        scale, nb = case["hist"]
        for w in ("vol", "num"):
            h = Histogram(p, lo, hi, nb, scale, w)
            p.histograms().append(h)
            hists.append((name_p, w, h))
    algo.result = [dict(contribs=contribs)]
    algo.histogram()
    q = np.asarray(data.q, float).ravel()
    I = np.asarray(data.f.binnedData, float).ravel()
    sig = np.asarray(data.f.binnedDataU, float).ravel()
    out = dict(q=q, I=I, sigma=sig, contribs=contribs, N=case["N"], max_iter=case["max_iter"],
               model=np.array(case["model"]), active=np.array([a[0] for a in case["active"]]),
               lo=np.array([a[1] for a in case["active"]]), hi=np.array([a[2] for a in case["active"]]),
               conv_crit=float(algo.convergenceCriterion()), comp_exp=float(algo.compensationExponent()),
               seed0=case["seed0"],
               scaling=np.array([o["scaling"] for o in outs]),
               background=np.array([o["background"] for o in outs]),
               num_iter=np.array([o["num_iter"] for o in outs]),
               attempts=np.array([o["attempts"] for o in outs]),
               elapsed=np.array([o["elapsed"] for o in outs]),
               fit_mean=np.mean([o["fit"] for o in outs], axis=0))
    # per-repetition chi-squared of the returned sets (BackgroundScalingFit on the model sum)
    from mcsas.mcsas.backgroundscalingfit import BackgroundScalingFit
    bg = BackgroundScalingFit(True, False, model)
    conv = []
    for r in range(R):
        md = model.calc(data, contribs[:, :, r], float(algo.compensationExponent()))
        sc = np.array([data.f.limit[1] / md.chisqrInt.max(), data.f.limit[0]])
        sc, cv, _, _ = bg.calc(data, md, sc)
        conv.append(cv)
    out["conval"] = np.array(conv)
    for pname, w, h in hists:
        tag = "%s_%s" % (pname, w)
        out[tag + "_edges"] = np.asarray(h.xLowerEdge, float)
        out[tag + "_bins_mean"] = np.asarray(h.bins.mean, float)
        out[tag + "_bins_std"] = np.asarray(h.bins.std, float)
        out[tag + "_cdf_mean"] = np.asarray(h.cdf.mean, float)
        out[tag + "_moments"] = np.array(h.moments.fields, float)       # (value, std) x 5
        out[tag + "_cfg"] = np.array([h.lower, h.upper, h.binCount])
        out[tag + "_scale"] = np.array(case["hist"][0])
    np.savez_compressed(os.path.join(OUT, "stats_%s.npz" % name), **out)
    m = out["%s_vol_moments" % case["active"][0][0]]
    print("stats %-16s R=%d  numIter mean %.0f  attempts %s  chi2 %.4f..%.4f  vol mean %.5g +- %.2g  "
          "var %.4g +- %.2g  (%.0f s wall, %.0f s cpu)" % (
              name, R, out["num_iter"].mean(), out["attempts"].tolist(), out["conval"].min(),
              out["conval"].max(), m[2], m[3], m[4], m[5], time.time() - t0, out["elapsed"].sum()))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--procs", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    for nm in (a.cases or list(CASES)):
        run_case(nm, a.reps, a.procs)
