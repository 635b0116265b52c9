#!/usr/bin/env python
"""Experiment driver (not part of the product): time the config-2 fit under scheduler / kernel
knobs (Engine.set_tuning = struct mcsas_tuning) and run-config overrides."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, FP32  # noqa: E402

NQ = int(os.environ.get("TUNE_NQ", "1000"))
REPS = int(os.environ.get("TUNE_REPS", "100"))
VERBOSE = int(os.environ.get("TUNE_VERBOSE", "0"))


def run(tag, tuning, seeds=(1, 2, 4, 7), **cfgkw):
    if os.environ.get("SEEDS"):
        seeds = tuple(int(x) for x in os.environ["SEEDS"].split(","))
    q, I, sig = synth.sphere_curve(nq=NQ)
    eng = Engine(0)
    eng.upload_datasets([(q, I, sig)])
    eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
    eng.set_tuning(verbose=VERBOSE, **tuning)
    out = []
    for seed in seeds:
        cfg = Engine.make_config(n_contribs=300, n_reps=REPS, precision=FP32, seed=7919 * seed, **cfgkw)
        t0 = time.perf_counter()
        try:
            ms = eng.run(cfg)
        except Exception as ex:
            print("%-34s FAILED: %s" % (tag, ex), flush=True)
            eng.close()
            return
        wall = time.perf_counter() - t0
        r = eng.fetch()
        info = eng.run_info()
        steps = int(r["steps_all"].sum())
        out.append((ms, wall * 1e3, steps, info["launches"], int(r["num_iter"].max()),
                    2.0 * steps / r["evals"].sum(), float((r["conval"] <= 1).mean()),
                    int(r["contribs"].view(np.uint64).sum() & 0xFFFFFFFF)))
    eng.close()
    for ms, wall, steps, nl, mx, eff, conv, chk in out[1:] or out:
        print("%-34s kernel %8.2f ms wall %8.2f ms  %.3e steps/s  launches %3d  max_iter %6d  spec_eff %.2f conv %.2f  chk %08x"
              % (tag, ms, wall, steps / ms * 1e3, nl, mx, eff, conv, chk), flush=True)


def saturated(tag, tuning, n_sets=16):
    """16 datasets x 100 repetitions: every SM busy (bench.py's `saturated` record)."""
    eng = Engine(0)
    eng.upload_datasets(synth.sphere_batch(n_sets, nq=NQ))
    eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
    eng.set_tuning(verbose=VERBOSE, **tuning)
    for seed in (99, 100):
        ms = eng.run(Engine.make_config(n_contribs=300, n_reps=100, precision=FP32, seed=seed, max_retries=0))
        r = eng.fetch()
        steps = float(r["steps_all"].sum())
        print("%-34s saturated: kernel %8.2f ms  %.3e steps/s  (algorithmic MUFU fraction %.3f of 4.63e12/s)  chk %08x"
              % (tag, ms, steps / ms * 1e3, steps / ms * 1e3 * 6000 / 4.63e12,
                 int(r["contribs"].view(np.uint64).sum() & 0xFFFFFFFF)), flush=True)
    eng.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["base"]
    if "base" in which:
        run("auto", {})
    if "seeds" in which:                                 # fit time by seed (the first run warms up and is not printed)
        run("auto", {}, seeds=tuple(range(1, 27)))
    if "maxc" in which:
        for c in (8, 4, 2, 1, 8):
            run("max cluster=%d" % c, {"max_cluster": c}, seeds=(1, 2, 4, 7, 9, 11))
    if "slice" in which:
        for us in [int(x) for x in os.environ.get("SLICE_US", "500,1000,1500,2000,3000").split(",")]:
            run("slice_us=%d" % us, {"slice_us": us}, seeds=(1, 2, 4, 7, 9, 11))
    if "late" in which:
        for d in (-1, 20000, 30000, 40000, 50000, -1):
            run("late_steps=%d" % d, {"late_steps": d}, seeds=(1, 2, 4, 7, 9, 11))
    if "retry" in which:
        for rp in [int(x) for x in os.environ.get("RETRY_PCT", "-1,0,10,20,-1").split(",")]:
            run("retry_pct=%d" % rp, {"retry_pct": rp}, seeds=tuple(range(1, 27)))
    if "bpct" in which:
        for pct in [int(x) for x in os.environ.get("BPCT", "40,25,30,35,50,40").split(",")]:
            for bd in [int(x) for x in os.environ.get("BDIV", "2").split(",")]:
                run("backup_pct=%d,div=%d" % (pct, bd), {"backup_pct": pct, "backup_div": bd}, seeds=tuple(range(1, 27)))
    if "repack" in which:
        for rp in [int(x) for x in os.environ.get("REPACK", "0,15,25,40,0").split(",")]:
            run("repack_pct=%d" % rp, {"repack_pct": rp}, seeds=tuple(range(1, 27)))
    if "btop" in which:
        for bt in [int(x) for x in os.environ.get("BTOP", "0,8,16,24,36,0").split(",")]:
            run("backup_top=%d" % bt, {"backup_top": bt}, seeds=(1, 2, 4, 7, 9, 11, 13, 17, 19))
    if "l2" in which:
        for pct in [int(x) for x in os.environ.get("L2_PCT", "-1,40,55,70,85,-1").split(",")]:
            saturated("l2_keep_pct=%d" % pct, {"l2_keep_pct": pct})
        run("l2_keep_pct=-1", {"l2_keep_pct": -1}, seeds=(1, 2, 4, 7, 9, 11))
        run("l2_keep_pct=default", {}, seeds=(1, 2, 4, 7, 9, 11))
    if "backup" in which:
        for pct in (30, 40, 50, 60):
            for bd in (2, 4):
                run("backup_pct=%d div=%d" % (pct, bd), {"backup_pct": pct, "backup_div": bd}, seeds=(1, 2, 4, 7, 9, 11))
    if "half" in which:
        for h in (0, 1, 0, 1):
            run("half=%d" % h, {"half": h}, seeds=(1, 2, 4, 7, 9, 11))
        for h in (0, 1):
            saturated("half=%d" % h, {"half": h})
    if "early" in which:
        for top in [int(x) for x in os.environ.get("EARLY_TOP", "-1,8,12,16,24").split(",")]:
            for pct in [int(x) for x in os.environ.get("EARLY_PCT", "25,30,35").split(",")]:
                run("early_top=%d early_pct=%d" % (top, pct), {"early_top": top, "early_pct": pct}, seeds=(1, 2, 4, 7, 9, 11, 13))
                if top < 0:
                    break
    if "hedge" in which:
        for h in (-1, 5, 10, 20, -1, 10):
            run("hedge_pct=%d (%d reps)" % (h, REPS), {"hedge_pct": h}, seeds=(1, 2, 4, 7, 9, 11, 13))
    if "gram" in which:
        for g in (0, -1):
            run("gram=%d" % g, {"gram": g}, seeds=(1, 2, 4, 7, 9, 11))
    if "sched" in which:
        for us in [int(x) for x in os.environ.get("SLICE_US", "1000,1500,2500").split(",")]:
            for bd in [int(x) for x in os.environ.get("SCHED_BD", "2,4,8").split(",")]:
                run("slice_us=%d backup_div=%d" % (us, bd), {"slice_us": us, "backup_div": bd}, seeds=(1, 2, 4, 7, 9, 11))
    if "fixed" in which:                 # one cluster size for all chains (no re-packing)
        for c in (1, 2, 4, 8):
            run("cluster=%d" % c, {}, seeds=(1, 2), cluster_size=c)
