#!/usr/bin/env python
"""BASELINE.json configs 3-5 at full size on one B200 (not bench lines: evidence that the
other model families and the multi-dataset path run and converge; prints one JSON line
per config).  Data are synthesised with the GPU model evaluation (fp64)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, FP32, FP64  # noqa: E402
from mcsas_b200.data import SASData  # noqa: E402
from mcsas_b200.mcsas import McSAS  # noqa: E402
from mcsas_b200.models import CylindersIsotropic, EllipsoidsIsotropic, EllipsoidalCoreShell, Kholodenko  # noqa: E402

NM = 1e-9


def noisy(I, seed):
    rng = np.random.default_rng(seed)
    I = I / I[0] * 1e6 + 1.0
    sig = 0.01 * I
    return I + rng.normal(0., 1., I.size) * sig, sig


def fit(tag, model, q, I, sig, reps, max_iter=100000):
    algo = McSAS(seed=4242)
    algo.model, algo.data = model, SASData(q, I, sig)
    algo.numContribs.setValue(300)
    algo.numReps.setValue(reps)
    algo.maxIterations.setValue(max_iter)
    algo.showIncomplete.setValue(True)
    if os.environ.get("MAX_RETRIES"):
        algo.maxRetries.setValue(int(os.environ["MAX_RETRIES"]))
    t0 = time.perf_counter()
    algo.analyse()
    wall = time.perf_counter() - t0
    res = algo.result[0]
    steps = int(res["stepsAll"].sum())
    print(json.dumps({"config": tag, "n_reps": reps, "nq": int(q.size), "kernel_ms": algo.lastKernelMs,
                      "wall_s": wall, "mc_steps": steps, "mc_steps_per_s": steps / (algo.lastKernelMs * 1e-3),
                      "converged_fraction": float((res["conval"] <= 1.0).mean()),
                      "median_conval": float(np.median(res["conval"])),
                      "mean_iter": float(res["numIter"]), "launches": algo.engine.run_info()["launches"]}), flush=True)


def config3(reps):
    q = np.logspace(np.log10(0.03), np.log10(3.0), 500) * 1e9
    m = CylindersIsotropic()
    m.aspect.setActive(True)
    m.radius.setActiveRange((1 * NM, 50 * NM))
    m.aspect.setActiveRange((1., 20.))
    if os.environ.get("QUAD_GL"):                       # Gauss-Legendre orientation average, fewer nodes
        from mcsas_b200 import _lib
        m.quadRule, m.quadNodes = _lib.QUAD_GAUSS_LEGENDRE, int(os.environ["QUAD_GL"])
    rng = np.random.default_rng(3)
    radii = np.clip(rng.normal(5., 1., 400), 0.5, None) * NM
    pset = np.stack([radii, np.full_like(radii, 5.0)], 1)
    it = m.calc(q, pset, 1.0).cumInt          # sum of V^2 F^2
    I, sig = noisy(it, 3)
    fit("configs[2]: CylindersIsotropic radius+aspect, 500 q%s" % (
        ", GL %s nodes" % os.environ["QUAD_GL"] if os.environ.get("QUAD_GL") else ", reference 100-node rule"),
        m, q, I, sig, reps, max_iter=int(os.environ.get("CYL_MAXIT", "100000")))


def config_ell(reps):
    """Not a BASELINE config: the two other orientation-averaged models on the config-3 q grid."""
    q = np.logspace(np.log10(0.03), np.log10(3.0), 500) * 1e9
    rng = np.random.default_rng(6)
    for cls in (EllipsoidsIsotropic, EllipsoidalCoreShell):
        m = cls()
        act = [p for p in m.params() if hasattr(p, "isActive") and p.isActive()]
        act[0].setActiveRange((1 * NM, 50 * NM))
        radii = np.clip(rng.normal(8., 1.5, 300), 0.5, None) * NM
        it = m.calc(q, radii[:, None], 1.0).cumInt
        I, sig = noisy(it, 5)
        fit("%s (%s active), 500 q, reference %d-node rule" % (cls.__name__, act[0].name(), 100), m, q, I, sig, reps,
            max_iter=int(os.environ.get("ELL_MAXIT", "100000")))


def config4(reps):
    q = np.logspace(np.log10(0.01), np.log10(10.0), 500) * 1e9
    m = Kholodenko()
    pset = np.array([[1 * NM, 10 * NM, 1000 * NM]])
    it = m.calc(q, pset, 1.0).cumInt
    I, sig = noisy(it, 4)
    fit("configs[3]: Kholodenko worm, 500 q", m, q, I, sig, reps, max_iter=int(os.environ.get("KHO_MAXIT", "20000")))


def config5(n_sets, reps, nq):
    sets = synth.sphere_batch(n_sets, nq=nq)
    eng = Engine(0)
    eng.upload_datasets(sets)
    eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
    cfg = Engine.make_config(n_contribs=300, n_reps=reps, precision=FP32, seed=55)
    t0 = time.perf_counter()
    ms = eng.run(cfg)
    wall = time.perf_counter() - t0
    r = eng.fetch()
    steps = int(r["steps_all"].sum())
    print(json.dumps({"config": "configs[4] (subset): %d sphere datasets x %d reps, %d q" % (n_sets, reps, nq),
                      "kernel_ms": ms, "wall_s": wall, "mc_steps": steps, "mc_steps_per_s": steps / (ms * 1e-3),
                      "converged_fraction": float((r["conval"] <= 1.0).mean()),
                      "launches": eng.run_info()["launches"]}), flush=True)
    eng.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["3", "5"]
    if "3" in which:
        config3(int(os.environ.get("REPS", "100")))
    if "ell" in which:
        config_ell(int(os.environ.get("REPS", "100")))
    if "4" in which:
        config4(int(os.environ.get("REPS", "100")))
    if "5" in which:
        config5(int(os.environ.get("SETS", "64")), int(os.environ.get("REPS", "100")), int(os.environ.get("NQ", "1000")))
