#!/usr/bin/env python
"""Experiment driver: per-phase cycle profile of the fp32 chain kernel (MCSAS_B200_PROF)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, FP32
os.environ["MCSAS_B200_PROF"] = "1"
q, I, sig = synth.sphere_curve(nq=1000)
runs = []
for qpt in ("4", "2"):
    for cache in ("1",):
        runs.append(("qpt %s cache %s" % (qpt, cache), {"MCSAS_B200_QPT": qpt, "MCSAS_B200_CACHE": cache, "MCSAS_B200_CLUSTER": "1"}))
if len(sys.argv) > 1:
    runs = [("default", {})]
for tag, env in runs:
    os.environ.update(env)
    eng = Engine(0)
    eng.upload_datasets([(q, I, sig)])
    eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
    cfg = Engine.make_config(n_contribs=300, n_reps=int(os.environ.get('REPS', '100')), precision=FP32, seed=7919, max_iter=100000, max_retries=-1)
    print("==", tag, flush=True)
    sys.stderr.flush()
    try:
        ms = eng.run(cfg)
    except Exception as ex:
        print("FAILED", ex)
        continue
    r = eng.fetch()
    print("kernel ms %.2f steps %d conv %.2f" % (ms, r["steps_all"].sum(), (r["conval"] <= 1).mean()), flush=True)
    eng.close()
