#!/usr/bin/env python
"""BASELINE.json configs[4]: a batch of synthetic sphere datasets x nRep = 100, partitioned over
the GPUs of one box (dataset x repetition chains are independent, no data-path
collective; one all_gather of the per-dataset statistics at the end).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 tools/run_config5_dist.py            # SETS=4096 REPS=100 NQ=1000 BLOCK=64

Each rank synthesises and fits its own contiguous block of datasets, BLOCK datasets per launch
(BLOCK x REPS chains run in waves over the SMs).  Prints one JSON line on rank 0 (not a bench
line: evidence for the multi-dataset configuration at full size)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, dataprep, FP32  # noqa: E402
from mcsas_b200.dist import partition  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_sets = int(os.environ.get("SETS", "4096"))
    reps = int(os.environ.get("REPS", "100"))
    nq = int(os.environ.get("NQ", "1000"))
    block = int(os.environ.get("BLOCK", "64"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    first, count = partition(n_sets, world, rank)
    t0 = time.perf_counter()
    sets = synth.sphere_batch(count, nq=nq, first=first)
    t_synth = time.perf_counter() - t0
    eng = Engine(local)
    eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
    # per dataset: [converged fraction, mean chi2, mean contribution radius over all repetitions, MC steps]
    stats = np.zeros((count, 4))
    kernel_ms = 0.0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for b0 in range(0, count, block):
        # raw tables -> the vectors mcFit reads, for the whole block in ONE device call (SURVEY 8(f) f3:
        # uncertainty floor of DataConfig.fuMin = 1 %, no rebinning -- nBin = 0 keeps the nq points)
        raw = [(q * 1e-9, I, sig) for q, I, sig in sets[b0:b0 + block]]
        chunk = dataprep.prepare_batch(raw, eng, n_bin=int(os.environ.get("NBIN", "0")), fu_min=0.01)
        eng.upload_datasets(chunk)
        eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
        cfg = Engine.make_config(n_contribs=300, n_reps=reps, precision=FP32, seed=55,
                                 first_chain=(first + b0) * reps)
        kernel_ms += eng.run(cfg)
        r = eng.fetch()
        rad = r["contribs"][:, :, 0, :]                          # [S, N, R]
        stats[b0:b0 + len(chunk), 0] = (r["conval"] <= 1.0).mean(axis=1)
        stats[b0:b0 + len(chunk), 1] = r["conval"].mean(axis=1)
        stats[b0:b0 + len(chunk), 2] = rad.mean(axis=(1, 2))
        stats[b0:b0 + len(chunk), 3] = r["steps_all"].sum(axis=1)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    eng.close()
    t = torch.tensor([wall, kernel_ms * 1e-3], device="cuda", dtype=torch.float64)
    mine = torch.from_numpy(stats).cuda()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sizes = [partition(n_sets, world, k)[1] for k in range(world)]
        out = [torch.empty((s, 4), device="cuda", dtype=torch.float64) for s in sizes]
        dist.all_gather(out, mine)                               # NCCL over NVLink: the only collective
        allstats = torch.cat(out, 0).cpu().numpy()
    else:
        allstats = stats
    if rank == 0:
        wall, kern = float(t[0]), float(t[1])
        steps = float(allstats[:, 3].sum())
        print(json.dumps({
            "config": "configs[4]: %d synthetic sphere datasets x %d reps, %d q, %d GPU(s), %d datasets per launch"
                      % (n_sets, reps, nq, world, block),
            "n_gpus": world, "chains": n_sets * reps, "mc_steps": steps,
            "wall_s_max_over_ranks": wall, "kernel_s_max_over_ranks": kern,
            "mc_steps_per_s_wall": steps / wall, "mc_steps_per_s_kernel": steps / kern,
            "datasets_per_s": n_sets / wall, "synth_s_rank0": t_synth,
            "converged_fraction": float(allstats[:, 0].mean()),
            "mean_conval": float(allstats[:, 1].mean()),
            "datasets_fully_converged": int((allstats[:, 0] == 1.0).sum())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
