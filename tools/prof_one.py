#!/usr/bin/env python
"""A short config-2 workload for ncu captures (tools, not product): 148 chains (one per SM), 300
contributions, 1000 q, criterion 0 (chains never stop), launches of TUNE_SLICE steps: the third
launch (`ncu -s 2 -c 1`) is a steady-state slice (steps 16 384 .. 24 576 of every chain: moves every
~15 steps, mostly speculative batches, some Gram batches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, FP32
reps = int(os.environ.get("TUNE_REPS", "148"))
q, I, sig = synth.sphere_curve(nq=1000)
eng = Engine(0)
eng.upload_datasets([(q, I, sig)])
eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
if os.environ.get("TUNE_L2"):
    eng.set_tuning(l2_keep_pct=int(os.environ["TUNE_L2"]))
cfg = Engine.make_config(n_contribs=300, n_reps=reps, precision=FP32, seed=5, max_iter=int(os.environ.get("TUNE_MAXIT", "24576")),
                         conv_crit=0.0, max_retries=-1, slice_steps=int(os.environ.get("TUNE_SLICE", "8192")),
                         cluster_size=int(os.environ.get("TUNE_CLUSTER", "1")))
ms = eng.run(cfg)
r = eng.fetch()
print("span ms", ms, "steps", int(r["num_iter"].sum()), "moves", int(r["num_moves"].sum()), eng.run_info())
