#!/usr/bin/env python
"""Experiment: how well does a chain's chi-squared after n steps predict its total length?
(the scheduler gives the largest clusters to the chains it expects to run longest)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, FP32
q, I, sig = synth.sphere_curve(nq=1000)
eng = Engine(0)
eng.upload_datasets([(q, I, sig)])
eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
for seed in (7919, 7919 * 2, 7919 * 4):
    kw = dict(n_contribs=300, n_reps=100, precision=FP32, seed=seed, max_retries=-1)
    eng.run(Engine.make_config(max_iter=100000, **kw))
    full = eng.fetch()
    tot = full["num_iter"][0]
    order = np.argsort(-tot)
    print("seed", seed, "longest 8 chains:", order[:8], tot[order[:8]], "moves", full["num_moves"][0][order[:8]])
    for n in (2000, 5000, 10000, 20000, 40000):
        eng.run(Engine.make_config(max_iter=n, **kw))
        part = eng.fetch()
        cv = part["conval"][0]
        alive = tot > n
        r = np.argsort(-cv)                      # predicted-longest first
        top8 = set(order[:8])
        hit = [len(top8 & set(r[:k])) for k in (8, 16, 32, 48)]
        rho = np.corrcoef(np.argsort(np.argsort(cv[alive])), np.argsort(np.argsort(tot[alive])))[0, 1]
        print("  after %6d steps: alive %3d  rank corr(conval, total steps) %.2f  of the 8 longest in predicted top 8/16/32/48: %s"
              % (n, alive.sum(), rho, hit), flush=True)
eng.close()
