#!/usr/bin/env python
"""A few short fits for compute-sanitizer (racecheck / memcheck / synccheck): the fp32 production
kernel on 1, 2, 4 and 8 CTAs per chain (Gram and speculative batches, accepted moves, the
cluster exchange through st.async + mbarrier, the ring refill by the last warp), and the fp64
parity kernel.

    compute-sanitizer --tool racecheck python tools/san_one.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, FP32, FP64
q, I, sig = synth.sphere_curve(nq=int(os.environ.get("NQ", "1000")))
eng = Engine(0)
eng.upload_datasets([(q, I, sig)])
eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
for prec, clusters in ((FP32, [int(c) for c in os.environ.get("CLUSTERS", "1,2,4,8").split(",")]), (FP64, [1, 2])):
    for cl in clusters:
        cfg = Engine.make_config(n_contribs=300, n_reps=int(os.environ.get("REPS", "2")), precision=prec, seed=7919,
                                 max_iter=int(os.environ.get("MAXIT", "400")), max_retries=-1, cluster_size=cl,
                                 slice_steps=int(os.environ.get("SLICE", "150")))
        ms = eng.run(cfg)
        r = eng.fetch()
        print("precision %d cluster %d: kernel ms %.2f steps %d moves %d launches %d chi2 %s" % (
            prec, cl, ms, r["steps_all"].sum(), r["num_moves"].sum(), eng.run_info()["launches"], r["conval"].round(2)), flush=True)
eng.close()
