#!/usr/bin/env python
"""Experiment driver: per-phase cycle profile of the fp32 chain kernel (tuning knob `profile`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import _lib
# (cd mcsas_b200/csrc && make PROF=1 BUILD=../../build/csrc_prof OUT=../../build/libmcsas_b200_prof.so)
_lib.LIB_PATH = os.path.join(ROOT, "build", "libmcsas_b200_prof.so")
from mcsas_b200 import Engine, synth, FP32
q, I, sig = synth.sphere_curve(nq=1000)
for cl in [int(x) for x in os.environ.get("CLUSTERS", "1,2,4,8").split(",")]:
    eng = Engine(0)
    eng.upload_datasets([(q, I, sig)])
    eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
    eng.set_tuning(profile=1)
    if os.environ.get("SMEAR"):                          # the 26-point slit profile of tools/smear_throughput.py
        import numpy as np
        qo = np.linspace(0., 5e7, 26)
        w = np.ones(26)
        w[13:] = np.linspace(1., 0., 13)
        w /= 2. * np.trapezoid(w, qo)
        eng.set_smearing(qo, w, False)
    reps = int(os.environ.get('REPS', '16'))
    cfg = Engine.make_config(n_contribs=300, n_reps=reps, precision=FP32, seed=7919, max_iter=int(os.environ.get('MAXIT', '100000')), max_retries=-1,
                             cluster_size=cl)
    print("== cluster %d, %d chains" % (cl, reps), flush=True)
    sys.stderr.flush()
    try:
        ms = eng.run(cfg)
    except Exception as ex:
        print("FAILED", ex)
        continue
    r = eng.fetch()
    print("kernel ms %.2f steps %d conv %.2f" % (ms, r["steps_all"].sum(), (r["conval"] <= 1).mean()), flush=True)
    eng.close()
