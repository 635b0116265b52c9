#!/usr/bin/env python
"""Experiment driver (not part of the product): time the config-2 fit under layout
overrides passed through MCSAS_B200_* environment variables."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, FP32  # noqa: E402

NQ = int(os.environ.get("TUNE_NQ", "1000"))
REPS = int(os.environ.get("TUNE_REPS", "100"))


def run(tag, env, seeds=(1, 2, 4, 7), **cfgkw):
    for k in list(os.environ):
        if k.startswith("MCSAS_B200_") and k not in ("MCSAS_B200_DEBUG", "MCSAS_B200_BACKUP", "MCSAS_B200_BACKUP_DIV", "MCSAS_B200_KMAX", "MCSAS_B200_SLICE_BATCHES"):
            del os.environ[k]
    os.environ.update(env)
    q, I, sig = synth.sphere_curve(nq=NQ)
    eng = Engine(0)
    eng.upload_datasets([(q, I, sig)])
    eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
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
                    2.0 * steps / r["evals"].sum(), float((r["conval"] <= 1).mean())))
    eng.close()
    for ms, wall, steps, nl, mx, eff, conv in out[1:] or out:
        print("%-34s kernel %8.2f ms wall %8.2f ms  %.3e steps/s  launches %3d  max_iter %6d  spec_eff %.2f conv %.2f"
              % (tag, ms, wall, steps / ms * 1e3, nl, mx, eff, conv), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["base"]
    if "base" in which:
        run("auto", {})
    if "mixed" in which:
        for m in ("0", "1", "0", "1"):
            run("mixed=%s" % m, {"MCSAS_B200_MIXED": m}, seeds=(1, 2, 4, 7, 9, 11))
    if "factor" in which:
        for f in ("2", "4", "2", "4"):
            run("mixed factor=%s" % f, {"MCSAS_B200_MIXED_FACTOR": f}, seeds=(1, 2, 4, 7, 9, 11))
    if "eighths" in which:
        for e in ("9", "10", "11", "12", "13", "14"):
            run("mixed eighths=%s" % e, {"MCSAS_B200_MIXED_EIGHTHS": e}, seeds=(1, 2, 4, 7, 9, 11))
    if "gram" in which:
        for g in ("1", "0"):
            run("gram=%s" % g, {"MCSAS_B200_GRAM": g}, seeds=(1, 2, 4, 7, 9, 11))
    if "sched" in which:
        for sb in os.environ.get("SCHED_SB", "256,384,576").split(","):
            for bd in os.environ.get("SCHED_BD", "4,8,16").split(","):
                os.environ["MCSAS_B200_SLICE_BATCHES"] = sb
                os.environ["MCSAS_B200_BACKUP_DIV"] = bd
                run("slice_batches=%s backup_div=%s" % (sb, bd), {}, seeds=(1, 2, 4, 7, 9, 11))
        del os.environ["MCSAS_B200_SLICE_BATCHES"], os.environ["MCSAS_B200_BACKUP_DIV"]
    if "grid" in which:
        for c in (1, 2, 4, 8):
            for qpt in (1, 2, 4):
                run("cluster=%d qpt=%d" % (c, qpt), {"MCSAS_B200_CLUSTER": str(c), "MCSAS_B200_QPT": str(qpt)})
    if "slice" in which:
        for sl in (2048, 4096, 16384, 65536, 1000000):
            run("auto slice=%d" % sl, {"MCSAS_B200_SLICE": str(sl)})
    if "fast" in which:
        for c in (1, 2, 4, 8):
            for qpt in (1, 2, 4):
                for ps in (1, 2, 4):
                    run("cluster=%d qpt=%d ps=%d" % (c, qpt, ps), {"MCSAS_B200_CLUSTER": str(c), "MCSAS_B200_QPT": str(qpt),
                                                               "MCSAS_B200_PS": str(ps)}, seeds=(1,))
    if "ct" in which:
        for c in (1, 2, 4, 8):
            for t in (64, 128, 256):
                run("cluster=%d T=%d" % (c, t), {"MCSAS_B200_CLUSTER": str(c), "MCSAS_B200_T": str(t)}, seeds=(1,))
