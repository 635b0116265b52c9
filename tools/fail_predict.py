"""Experiment: how well does chi2 at 40/50/60 % of maxIterations predict a failed attempt?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, FP32
q, I, sig = synth.sphere_curve(nq=1000)
for seed in (7919, 7919 * 4, 7919 * 7):
    res = {}
    for mi in (40000, 50000, 60000, 100000):
        eng = Engine(0)
        eng.upload_datasets([(q, I, sig)])
        eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
        cfg = Engine.make_config(n_contribs=300, n_reps=100, precision=FP32, seed=seed, max_iter=mi, max_retries=-1, cluster_size=1)
        eng.run(cfg)
        r = eng.fetch()
        res[mi] = (r["conval"][0].copy(), r["num_iter"][0].copy())
        eng.close()
    fail = res[100000][0] > 1.0
    print("seed", seed, "failed attempts:", int(fail.sum()), "iters of converged: median %d max %d" % (np.median(res[100000][1][~fail]), res[100000][1][~fail].max()))
    for mi in (40000, 50000, 60000):
        c = res[mi][0]
        live = res[mi][1] >= mi
        print("  at %d: live %d; conval of eventual failures %s; of live non-failures: median %.3f  90%% %.3f max %.3f" % (
            mi, live.sum(), np.round(np.sort(c[fail]), 3), np.median(c[live & ~fail]), np.percentile(c[live & ~fail], 90), c[live & ~fail].max()))
