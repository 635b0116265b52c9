#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""bench.py -- MC steps/s and time-to-convergence of the McSAS fitting core on B200
(BASELINE.json metric).

Workload (configs[1]): Sphere model, synthetic 1000-point q curve prepared as the reference
prepares it (no rebinning, 1 % uncertainty floor), nContribs=300, nRep=100, convergenceCriterion=1,
production mode (fp32 form factors and delta sums, fp64 base sums / fit, in-kernel Philox).  One
bench "step" = one complete McSAS.analyse of the repetitions with a fresh seed; MC steps = every
executed iteration of the loop at mcsas.py:354-404 of the attempts that make up the result.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched with torchrun, one rank per GPU.  The headline value is weak scaling (100
repetitions per rank, no data-path collective, one NCCL all-gather of the final contribution
sets); the same line carries a strong-scaling record (ONE 100-repetition fit partitioned over
the ranks: time to convergence at N GPUs).  `--impl reference` times the UNMODIFIED reference
(oracle/_ref, vendored by __graft_entry__.build()) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NQ, N_CONTRIBS, N_REPS = 1000, 300, 100
RADIUS_RANGE = (1e-9, 100e-9)
MUFU_PER_STEP = 6 * NQ          # 2 evaluations x (sin, cos, rcp) per q point (algorithmic)
MUFU_EXECUTED_PER_STEP = 2 * NQ # what the kernel issues: 1 evaluation x (sin, cos); the old row is cached
FLOP_PER_STEP = 44 * NQ         # ~14 flop per evaluation + 16 flop LSQ per q point
METRIC = "MC steps/s (nRep=100 sphere fit, 1000 q, 300 contributions)"
# identical in both arms (the driver compares them)
CONFIG = {"workload": "configs[1]: Sphere, synthetic 1000-point q curve (no rebinning, 1 % uncertainty floor), "
                      "nContribs=300, nRep=100, convergenceCriterion=1, radius 1-100 nm",
          "nq": NQ, "n_contribs": N_CONTRIBS, "n_reps": N_REPS}


def config2_data():
    """(q [1/m], I, sigma) exactly as the reference's SASData hands them to mcFit for this input:
    nBin = 0, sigma <- max(sigma, fuMin * I) (dataobj.py:204-226)."""
    from mcsas_b200 import synth, dataprep
    q, I, sig = synth.sphere_curve(nq=NQ)
    return q, I, dataprep.prepare_uncertainty(I, sig, 0.01)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super(ClockSampler, self).__init__(daemon=True)
        self.gpu, self.rows, self._stop_evt = gpu, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.25)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU baselines --
def oracle_cpu_steps_per_s(steps_per_chain, threads):
    """The oracle's C restatement of McSAS.mcFit on `threads` host threads, one independent
    chain each (kind "port")."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.lib()
    q, I, sig = config2_data()
    om = orc.Model("Sphere", ranges={"radius": RADIUS_RANGE})
    st = orc.settings(n_contribs=N_CONTRIBS, conv_crit=0.0, max_iter=steps_per_chain)
    done = [0] * threads

    def work(i):
        stream = np.random.RandomState(1000 + i).uniform(size=N_CONTRIBS + steps_per_chain)
        o = orc.chain(om, st, q, I, sig, stream)
        done[i] = o["num_iter"]
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    el = time.perf_counter() - t0
    return sum(done) / el, el


def _ref_worker_init():
    """One process per core: import the unmodified reference once (oracle/refharness.py stubs)."""
    global _REF
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import logging
    import refharness as rh
    rh.install()
    logging.getLogger().setLevel(logging.ERROR)
    q, I, sig = config2_data()
    model = rh.model_class("Sphere")()
    model.radius.setActiveRange(RADIUS_RANGE)
    data = rh.make_data(q * 1e-9, I, sig, nbin=0)       # fuMin floor applied by the reference itself
    _REF = (rh, model, data)


def _ref_worker(args):
    """One bounded sample: the reference's own McSAS.mcFit on the config-2 data for `steps`
    iterations (criterion 0: it uses them all); returns (steps made, seconds inside mcFit)."""
    seed, steps = args
    rh, model, data = _REF
    algo = rh.make_algo(model, data, numContribs=N_CONTRIBS, maxIterations=steps, convergenceCriterion=1e-12)
    np.random.seed(seed)
    t0 = time.perf_counter()
    rset, meas, conval, details = algo.mcFit(N_CONTRIBS, True, True, nRun=0)
    return int(details["numIterations"]), time.perf_counter() - t0


class ReferencePool(object):
    """The unmodified reference (single-threaded Python) as `procs` independent processes."""

    def __init__(self, procs):
        import multiprocessing as mp
        self.procs = procs
        self.pool = mp.get_context("spawn").Pool(procs, initializer=_ref_worker_init)
        self.pool.map(_ref_worker, [(i, 20) for i in range(procs)])       # imports + first-call costs

    def sample(self, steps, seed0):
        t0 = time.perf_counter()
        out = self.pool.map(_ref_worker, [(seed0 + i, steps) for i in range(self.procs)], chunksize=1)
        wall = time.perf_counter() - t0
        n = sum(o[0] for o in out)
        per_core = float(np.mean([o[0] / o[1] for o in out]))
        return n / wall, wall, per_core

    def close(self):
        self.pool.close()
        self.pool.join()


def reference_available():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refharness as rh
    return rh.available()


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout: libraries also write there (NCCL's version banner
    on the first collective), so everything else is sent to stderr until emit()."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path on all host cores."""
    if rank != 0:
        return
    threads = host_threads()
    line = {"impl": "reference", "metric": METRIC, "unit": "MC steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(CONFIG),
            "gpu_launches": 0}
    if reference_available():
        # bounded sample per bench step: every core runs the reference's McSAS.mcFit for 2500
        # iterations of the config-2 chain (~2 s at ~1.3 k steps/s/core)
        per_chain = 2500
        pool = ReferencePool(threads)
        vals, tot_t, cores = [], 0.0, []
        for i in range(args.warmup + args.steps):
            v, el, pc = pool.sample(per_chain, 1000 * (i + 1))
            if i >= args.warmup:
                vals.append(v); tot_t += el; cores.append(pc)
        pool.close()
        value = float(np.mean(vals))
        # the C port beside it (second number; not the reported baseline)
        port_v, port_el = oracle_cpu_steps_per_s(20000, threads)
        import scipy
        line.update({
            "value": value, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
            "cpu_baseline": {"value": value, "unit": "MC steps/s", "cores": threads, "kind": "reference",
                             "sample": "%d processes x McSAS.mcFit of the UNMODIFIED reference (oracle/_ref, "
                                       "mcsas.py:287-439) x %d MC steps (+300 initial contributions) of the "
                                       "config-2 chain per bench step; %.0f steps/s per core; numpy %s, scipy %s"
                                       % (threads, per_chain, float(np.mean(cores)), np.__version__, scipy.__version__),
                             "per_core": float(np.mean(cores))},
            "cpu_baseline_port": {"value": port_v, "unit": "MC steps/s", "cores": threads, "kind": "port",
                                  "sample": "%d threads x 1 chain x 20000 MC steps, oracle/mcsas_oracle.c (%.1f s)"
                                            % (threads, port_el)},
            "e2e": {"value": value, "unit": "MC steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    else:
        per_chain = 20000
        vals, tot_t = [], 0.0
        for i in range(args.warmup + args.steps):
            v, el = oracle_cpu_steps_per_s(per_chain, threads)
            if i >= args.warmup:
                vals.append(v); tot_t += el
        value = float(np.mean(vals))
        line.update({
            "value": value, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
            "cpu_baseline": {"value": value, "unit": "MC steps/s", "cores": threads, "kind": "port",
                             "sample": "oracle/_ref absent (run __graft_entry__.build() where /root/reference is "
                                       "mounted): %d threads x 1 chain x %d MC steps of oracle/mcsas_oracle.c, the C "
                                       "port of McSAS.mcFit" % (threads, per_chain)},
            "e2e": {"value": value, "unit": "MC steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(line)


# ------------------------------------------------------------------ other BASELINE configs --
def sub_records(peaks, sm_count):
    """Bounded runs of BASELINE configs[2..4] on one B200 (sub-records of the N = 1 line): MC steps/s,
    algorithmic MUFU-roofline fraction (SURVEY 8(d) work per step), converged fraction."""
    from mcsas_b200 import Engine, synth, FP32, _lib
    from mcsas_b200.data import SASData
    from mcsas_b200.mcsas import McSAS
    from mcsas_b200.models import CylindersIsotropic, Kholodenko
    NM = 1e-9
    out = {}

    def noisy(I, seed):
        rng = np.random.default_rng(seed)
        I = I / I[0] * 1e6 + 1.0
        sig = 0.01 * I
        return I + rng.normal(0., 1., I.size) * sig, sig

    def fit(model, q, I, sig, max_iter, mufu_per_step, note):
        algo = McSAS(seed=4242)
        algo.model, algo.data = model, SASData(q, I, sig)
        algo.numContribs.setValue(300)
        algo.numReps.setValue(100)
        algo.maxIterations.setValue(max_iter)
        algo.showIncomplete.setValue(True)
        eng = algo.engine
        cfg = algo._config(300, 100)
        cfg.max_retries = -1                              # bounded: one attempt per repetition
        q_, I_, s_ = q, I, sig
        eng.upload_datasets([(q_, I_, s_)])
        model.compensationExponent = float(algo.compensationExponent())
        model.configureEngine(eng)
        eng.set_smearing(None)
        ms = eng.run(cfg)
        r = eng.fetch()
        steps = float(r["steps_all"].sum())
        v = steps / (ms * 1e-3)
        rec = {"value": v, "unit": "MC steps/s", "kernel_ms": ms, "mc_steps": steps,
               "converged_fraction": float((r["conval"] <= 1.0).mean()),
               "median_chi2": float(np.median(r["conval"])),
               "mufu_per_step_algorithmic": mufu_per_step,
               "roofline_frac": v * mufu_per_step * 1e-9 / peaks["mufu_gops"], "note": note}
        eng.close()
        algo._engine = None
        return rec

    # configs[2]: CylindersIsotropic, radius + aspect active, 500 q; data from the model itself
    q = np.logspace(np.log10(0.03), np.log10(3.0), 500) * 1e9
    for tag, rule, nodes in (("configs[2] cylinder, reference 100-node rule", _lib.QUAD_REFERENCE, None),
                             ("configs[2] cylinder, 32-node Gauss-Legendre", _lib.QUAD_GAUSS_LEGENDRE, 32)):
        m = CylindersIsotropic()
        m.aspect.setActive(True)
        m.radius.setActiveRange((1 * NM, 50 * NM))
        m.aspect.setActiveRange((1., 20.))
        rng = np.random.default_rng(3)
        radii = np.clip(rng.normal(5., 1., 400), 0.5, None) * NM
        it = m.calc(q, np.stack([radii, np.full_like(radii, 5.0)], 1), 1.0).cumInt
        I, sig = noisy(it, 3)
        m.quadRule = rule
        if nodes:
            m.quadNodes = nodes
        D = nodes or 100
        out[tag] = fit(m, q, I, sig, 20000, 2 * 500 * D * 4,
                       "100 repetitions x 20000 steps, one attempt (bounded); 2*nq*D*4 MUFU per step")
    # configs[3]: Kholodenko worm, 3 active parameters, 500 q from 0.01 to 10 1/nm
    q = np.logspace(np.log10(0.01), np.log10(10.0), 500) * 1e9
    m = Kholodenko()
    it = m.calc(q, np.array([[1 * NM, 10 * NM, 1000 * NM]]), 1.0).cumInt
    I, sig = noisy(it, 4)
    out["configs[3] Kholodenko"] = fit(m, q, I, sig, 10000, 1.5e6,
                                       "100 repetitions x 10000 steps, one attempt (bounded); 1.5e6 MUFU per step "
                                       "(SURVEY 8(d): 2*nq*E*3 with E = 512 integrand evaluations)")
    # configs[4]: 64-dataset slice of the 4096-dataset batch, 100 repetitions each, every SM busy
    eng = Engine(0)
    eng.upload_datasets(synth.sphere_batch(64, nq=NQ))
    eng.set_model(0, [10e-9, 1e14], [0], [RADIUS_RANGE[0]], [RADIUS_RANGE[1]], [0])
    ms = eng.run(Engine.make_config(n_contribs=N_CONTRIBS, n_reps=N_REPS, precision=FP32, seed=55))
    r = eng.fetch()
    steps = float(r["steps_all"].sum())
    v = steps / (ms * 1e-3)
    out["configs[4] slice: 64 datasets x 100 repetitions"] = {
        "value": v, "unit": "MC steps/s", "kernel_ms": ms, "mc_steps": steps, "chains": 6400,
        "converged_fraction": float((r["conval"] <= 1.0).mean()), "mufu_per_step_algorithmic": MUFU_PER_STEP,
        "roofline_frac": v * MUFU_PER_STEP * 1e-9 / peaks["mufu_gops"],
        "note": "full criterion and retries; 64 of the 4096 datasets (tools/run_config5_dist.py runs all of them)"}
    eng.close()
    return out


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel, measured with ncu on THIS round's build over the
    launches of the headline fit (profiles/ncu_traffic.json, written by tools/launch_traffic.py from the
    kept launch list's dram__bytes_read.sum + dram__bytes_write.sum; its `saturated_slice` is the
    ncu --set full capture of one steady-state slice with every SM busy); None when nothing is on record."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            t = json.load(fh)
        return t
    except Exception:
        return None


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true")
    ap.add_argument("--tune", default="", help="experiments: scheduler knobs of struct mcsas_tuning, k=v[,k=v...]")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mcsas_b200 import dist as mdist
    from mcsas_b200.mcsas import McSAS
    from mcsas_b200.models import Sphere
    from mcsas_b200.data import SASData

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # one rank per GPU shares the host's cores: the few host-side tensor copies of a gather must not
        # fan out into cpu_count() threads per rank
        torch.set_num_threads(max(1, (os.cpu_count() or world) // (2 * world)))
    q, I, sig = config2_data()
    data = SASData(q, I, sig)

    def make_algo(n_reps, first_chain):
        model = Sphere()
        model.radius.setActiveRange(RADIUS_RANGE)
        algo = McSAS(device=local_rank, seed=0, first_chain=first_chain)
        algo.model, algo.data = model, data
        algo.numContribs.setValue(N_CONTRIBS)
        algo.numReps.setValue(n_reps)
        algo.convergenceCriterion.setValue(1.0)
        return algo

    algo = make_algo(N_REPS, rank * N_REPS)                         # weak scaling: 100 repetitions per rank
    if args.tune:
        algo.engine.set_tuning(**{k: int(v) for k, v in (kv.split("=") for kv in args.tune.split(","))})
    s_first, s_count = mdist.partition(N_REPS, world, rank)         # strong scaling: ONE fit over all ranks
    algo_strong = make_algo(max(1, s_count), s_first) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    gatherer = mdist.Gatherer(local_rank, max_rows=N_REPS, n_contribs=N_CONTRIBS, n_active=1) if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(al, i):
        """One complete analyse() with host buffers in and out; at N > 1 the all-gather of the
        previous step's contribution sets completes behind this step's fit."""
        flush.zero_()
        torch.cuda.synchronize()
        al.seed = 7919 * (i + 1)
        al.result = []                      # analyse() appends: never read a stale fit
        t0 = time.perf_counter()
        al.analyse()                        # upload (H2D) + chains + fetch (D2H)
        if not al.result:
            raise RuntimeError("bench step %d: the fit did not converge (no result)" % i)
        res = al.result[-1]
        launches[0] += al.engine.run_info()["launches"]
        if gatherer is not None:
            gatherer.wait()                 # previous step's gather (overlapped with this fit)
            gatherer.start(res)             # NCCL all_gather_into_tensor, asynchronous
        t1 = time.perf_counter()
        return res, al.lastKernelMs, t1 - t0

    launches = [0]
    for i in range(args.warmup):
        one_step(algo, i)
    if gatherer is not None:
        gatherer.wait()
    launches[0] = 0
    # rank 0 samples its GPU (every nvidia-smi call takes a driver lock: 8 ranks polling at 10 Hz slowed
    # every rank's scheduler loop by ~10 %)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.start()
    barrier()
    mc_steps, kern_ms, e2e_s, evals, t2c, nconv = 0, 0.0, 0.0, 0, [], 0
    wall0 = time.perf_counter()
    for i in range(args.steps):
        res, kms, el = one_step(algo, args.warmup + i)
        mc_steps += int(res["stepsAll"].sum())      # every executed MC step, retries included
        evals += int(res["evals"].sum())
        nconv += int((res["conval"] <= 1.0).sum())
        kern_ms += kms
        e2e_s += el
        t2c.append(kms)
    if gatherer is not None:
        t0 = time.perf_counter()
        gatherer.wait()                     # the last step's gather is inside the timed region
        e2e_s += time.perf_counter() - t0
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.finish() if sampler is not None else None

    # ---- strong scaling: one nRep=100 fit partitioned over the ranks -----------------------------
    strong = None
    if world > 1:
        launches_weak = launches[0]
        st_ms, st_e2e, st_steps = [], [], 0
        for i in range(2 + args.steps):
            barrier()
            res, kms, el = one_step(algo_strong, 100 + i)
            t0 = time.perf_counter()
            gatherer.wait()                 # a single fit has nothing to overlap its gather with
            el += time.perf_counter() - t0
            tt = torch.tensor([kms, el], dtype=torch.float64, device="cuda")
            ws = torch.tensor([float(res["stepsAll"].sum()) if res is not None else 0.], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ws, op=dist.ReduceOp.SUM)
            if i >= 2:
                st_ms.append(float(tt[0])); st_e2e.append(float(tt[1])); st_steps += float(ws[0])
        launches[0] = launches_weak
        strong = {"what": "ONE nRep=100 fit, repetitions block-partitioned over the ranks (dist.partition)",
                  "time_to_convergence_ms": float(np.mean(st_ms)), "e2e_ms": 1e3 * float(np.mean(st_e2e)),
                  "value": st_steps / (sum(st_ms) * 1e-3), "unit": "MC steps/s",
                  "reps_per_rank": [mdist.partition(N_REPS, world, r)[1] for r in range(world)], "steps": args.steps}

    # max over ranks of the time, sum over ranks of the work
    tt = torch.tensor([kern_ms, e2e_s, wall], dtype=torch.float64, device="cuda")
    ww = torch.tensor([mc_steps, evals, nconv], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ww, op=dist.ReduceOp.SUM)
    kern_ms, e2e_s, wall = [float(x) for x in tt.tolist()]
    mc_steps, evals, nconv = [float(x) for x in ww.tolist()]

    if rank == 0:
        peaks = algo.engine.measure_peaks()
        info = algo.engine.device_info()
        value = mc_steps / (kern_ms * 1e-3)
        e2e_value = mc_steps / e2e_s
        # the roofline is per GPU: rank-0-equivalent share of the work
        per_gpu = value / world
        mufu_ach = per_gpu * MUFU_PER_STEP * 1e-9
        fp32_ach = per_gpu * FLOP_PER_STEP * 1e-12
        h2d = 3 * NQ * 8 + 256
        d2h = N_REPS * (N_CONTRIBS * 8 + NQ * 8 + 8 * 8)
        traffic = measured_traffic()
        cfg = dict(CONFIG)                  # identical to the reference arm's
        run_details = {"n_reps_per_gpu": N_REPS, "precision": "fp32 production (Philox proposals)",
                       "l2": "256 MB flush between steps; the chain state is register/shared-memory resident "
                             "(24 KB of input per launch)"}
        line = {
            "metric": METRIC, "value": value, "unit": "MC steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kern_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "run": run_details,
            "time_to_convergence_ms": float(np.mean(t2c)),
            "converged_fraction": nconv / (args.steps * N_REPS * world),
            "mc_steps_per_fit": mc_steps / (args.steps * world),
            "speculation_efficiency": 2.0 * mc_steps / max(1.0, evals),
            "e2e": {"value": e2e_value, "unit": "MC steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "efficiency_vs_device_timed": e2e_value / value},
            "wall_s": wall,
            "gpu_launches": launches[0],     # chain-kernel launches in the timed region (rank 0)
            "clocks": clocks,
            "roofline": {"bound": "sfu", "achieved": mufu_ach, "peak": peaks["mufu_gops"],
                         "unit": "GMUFU/s", "frac": mufu_ach / peaks["mufu_gops"],
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_per_mc_step": traffic.get("dram_bytes_per_mc_step") if traffic else None,
                         "traffic_source": traffic.get("source") if traffic else
                         "no ncu --set full capture of this build on record (profiles/ncu_traffic.json)",
                         "executed_mufu_frac": per_gpu * MUFU_EXECUTED_PER_STEP * 1e-9 / peaks["mufu_gops"],
                         "peak_source": "measured live: mcsas_measure_peaks (sin+cos+rcp MUFU mix)",
                         "fp32": {"achieved": fp32_ach, "peak": peaks["ffma_tflops"],
                                  "unit": "TFLOP/s", "frac": fp32_ach / peaks["ffma_tflops"]},
                         "sm_count": info["sm_count"], "chains": N_REPS},
        }
        if strong is not None:
            line["strong_scaling"] = strong
        if world == 1:
            # the same kernel with every SM busy: 16 datasets x 100 repetitions (not part of `value`)
            from mcsas_b200 import Engine, FP32, synth
            eng = algo.engine
            eng.upload_datasets(synth.sphere_batch(16, nq=NQ))
            eng.set_model(0, [10e-9, 1e14], [0], [RADIUS_RANGE[0]], [RADIUS_RANGE[1]], [0])
            ms_sat = eng.run(Engine.make_config(n_contribs=N_CONTRIBS, n_reps=N_REPS, precision=FP32, seed=99,
                                                max_retries=0))
            rs = eng.fetch()
            sat = float(rs["steps_all"].sum()) / (ms_sat * 1e-3)
            line["saturated"] = {"value": sat, "unit": "MC steps/s", "chains": 1600, "kernel_ms": ms_sat,
                                 "mufu_frac_algorithmic": sat * MUFU_PER_STEP * 1e-9 / peaks["mufu_gops"],
                                 "mufu_frac_executed": sat * MUFU_EXECUTED_PER_STEP * 1e-9 / peaks["mufu_gops"]}
            if traffic and traffic.get("saturated_slice"):      # ncu --set full capture of a steady-state slice
                line["saturated"]["dram_bytes_per_mc_step"] = traffic["saturated_slice"]["dram_bytes_per_mc_step"]
                line["saturated"]["traffic_source"] = traffic["saturated_slice"]["source"]
            if not args.no_sub_records:
                line["sub_records"] = sub_records(peaks, info["sm_count"])
        if not args.no_cpu_baseline and world == 1:
            threads = host_threads()
            v, el = oracle_cpu_steps_per_s(100000, threads)
            port = {"value": v, "unit": "MC steps/s", "cores": threads, "kind": "port",
                    "sample": "%d threads x 1 chain x 100000 MC steps of the same workload (%.1f s); "
                              "oracle/mcsas_oracle.c" % (threads, el)}
            line["cpu_baseline"] = port
            if reference_available():
                pool = ReferencePool(threads)
                rv, rel, pc = pool.sample(6000, 4242)
                pool.close()
                line["cpu_baseline"] = {
                    "value": rv, "unit": "MC steps/s", "cores": threads, "kind": "reference", "per_core": pc,
                    "sample": "%d processes x McSAS.mcFit of the UNMODIFIED reference (oracle/_ref) x 6000 MC steps of "
                              "the same chain (%.1f s)" % (threads, rel)}
                line["cpu_baseline_port"] = port
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
