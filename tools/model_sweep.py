"""Experiment: production (fp32) chains of every model on its golden small case, timing + sanity."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import common, gpu_common
import oracle as orc
from mcsas_b200 import Engine, FP32
for name in ["sphere_small", "lma_small", "sphcs_small", "ellcs_small", "elliso_small", "cyl_small", "gauss_small", "kho_small"]:
    g = common.load_chain(name)
    om = common.oracle_model(g)
    reps = 8 if name == "kho_small" else 32
    mi = 400 if name == "kho_small" else 4000
    eng = gpu_common.engine_for(om, [(g["q"], g["I"], g["sigma"])])
    cfg = Engine.make_config(n_contribs=g["N"], n_reps=reps, conv_crit=float(g["conv_crit"]), max_iter=mi, max_retries=-1,
                             comp_exp=float(g["comp_exp"]), precision=FP32, seed=5)
    t0 = time.perf_counter()
    ms = eng.run(cfg)
    r = eng.fetch()
    eng.close()
    # fp64 oracle re-evaluation of the returned sets
    errs = []
    for rep in range(2):
        ito, _, _, _ = orc.eval_contribs(om, g["q"], r["contribs"][0, :, :, rep], float(g["comp_exp"]))
        _, conval, _ = orc.bgscale(g["I"], g["sigma"], ito.sum(0))
        errs.append(abs(conval / r["conval"][0, rep] - 1))
    steps = int(r["steps_all"].sum())
    print("%-14s nq %4d N %3d reps %2d: %8.2f ms  %.3e steps/s  conval median %.4g  oracle-vs-kernel chi2 rel %.1e" % (
        name, g["q"].size, g["N"], reps, ms, steps / ms * 1e3, np.median(r["conval"]), max(errs)), flush=True)
