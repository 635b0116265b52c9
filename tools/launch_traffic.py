#!/usr/bin/env python
"""DRAM traffic of the headline fit from a kept ncu launch list (no GPU needed):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1
    python tools/launch_traffic.py gpurun_out/launches.csv <MC steps per fit> [saturated-slice summary json]
The fits of bench.py are the groups of chain_fast_kernel<0,...> launches between two L2 flushes
(at::vectorized_elementwise_kernel).  Writes profiles/ncu_traffic.json, which bench.py reports as
roofline.traffic (per launch of the dominant kernel, like roofline.achieved)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, steps_per_fit = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
L = {}
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    d = L.setdefault(int(r[ix["ID"]]), {"k": r[ix["Kernel Name"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
fits, cur = [], None
for i in sorted(L):
    k = L[i]["k"]
    if "vectorized_elementwise" in k:
        if cur:
            fits.append(cur)
        cur = []
    elif cur is not None and "chain_fast_kernel<0" in k:
        cur.append(L[i])
    elif cur:
        fits.append(cur)
        cur = None
if cur:
    fits.append(cur)
fits = [f for f in fits if f]
n = sum(len(f) for f in fits)
byt = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for f in fits for d in f)
ms = sum(d["gpu__time_duration.sum"] for f in fits for d in f) * 1e-6
out = {"kernel": fits[0][0]["k"],
       "dram_bytes_per_launch": byt / n,
       "launches_per_fit": n / len(fits),
       "dram_bytes_per_fit": byt / len(fits),
       "mc_steps_per_fit": steps_per_fit,
       "dram_bytes_per_mc_step": byt / len(fits) / steps_per_fit,
       "launch_ms": ms / n,
       "source": "%s: %d fits of bench.py (nRep = 100, 100 chains, %d launches), ncu --metrics dram__bytes_read.sum,"
                 "dram__bytes_write.sum --clock-control none (per-launch, caches flushed before every launch)"
                 % (os.path.join("profiles", os.path.basename(path)), len(fits), n)}
if len(sys.argv) > 3:
    out["saturated_slice"] = json.load(open(sys.argv[3]))
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
