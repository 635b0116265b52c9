#!/usr/bin/env python
"""Experiment (not product): production-mode throughput of a slit-smeared sphere fit
(config-2 curve, 26-point trapezoidal profile as sasconfig.py builds it for nSteps = 25)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, FP32  # noqa: E402

nq = int(os.environ.get("NQ", "1000"))
reps = int(os.environ.get("REPS", "100"))
q, I, sig = synth.sphere_curve(nq=nq)
qo = np.linspace(0., 5e7, 26)                     # 0 .. 0.05 1/nm
w = np.ones(26)
w[13:] = np.linspace(1., 0., 13)
w /= 2. * np.trapezoid(w, qo)
eng = Engine(0)
eng.upload_datasets([(q, I, sig)])
eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
for smear in (False, True):
    eng.set_smearing(qo, w, False) if smear else eng.set_smearing()
    cfg = Engine.make_config(n_contribs=300, n_reps=reps, precision=FP32, seed=5, conv_crit=0.0, max_retries=-1,
                             max_iter=int(os.environ.get("MAXIT", "20000")))
    ms = eng.run(cfg)
    r = eng.fetch()
    steps = int(r["steps_all"].sum())
    print("smeared %-5s nq %d reps %d: %9.2f ms  %.3e MC steps/s  median chi2 %.4g" % (
        smear, nq, reps, ms, steps / ms * 1e3, float(np.median(r["conval"]))), flush=True)
eng.close()
