#!/usr/bin/env python
"""Experiment: which repetitions does a config-2 fit wait for at the end?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, FP32
q, I, sig = synth.sphere_curve(nq=1000)
eng = Engine(0)
eng.upload_datasets([(q, I, sig)])
eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
eng.set_tuning(verbose=1)
for seed in [int(x) for x in os.environ.get("SEEDS", "1,13").split(",")]:
    ms = eng.run(Engine.make_config(n_contribs=300, n_reps=100, precision=FP32, seed=7919 * seed))
    r = eng.fetch()
    order = np.argsort(-r["steps_all"][0])[:8]
    print("seed %d: %.2f ms; repetitions with the most steps: %s" % (seed, ms, [
        (int(c), int(r["attempts"][0][c]), int(r["num_iter"][0][c]), int(r["steps_all"][0][c]), round(float(r["elapsed"][0][c]) * 1e3, 1))
        for c in order]), flush=True)
    sys.stderr.flush()
eng.close()
