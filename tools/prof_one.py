#!/usr/bin/env python
"""One fit of a reduced config-2 workload for ncu captures (tools, not product)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, FP32
reps = int(os.environ.get("TUNE_REPS", "148"))
q, I, sig = synth.sphere_curve(nq=1000)
eng = Engine(0)
eng.upload_datasets([(q, I, sig)])
eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
cfg = Engine.make_config(n_contribs=300, n_reps=reps, precision=FP32, seed=5, max_iter=int(os.environ.get("TUNE_MAXIT", "4096")),
                         conv_crit=0.0, max_retries=-1, slice_steps=int(os.environ.get("TUNE_SLICE", "-1")))
ms = eng.run(cfg)
r = eng.fetch()
print("kernel ms", ms, "steps", int(r["num_iter"].sum()), "steps/s %.3e" % (r["num_iter"].sum() / ms * 1e3), eng.run_info())
