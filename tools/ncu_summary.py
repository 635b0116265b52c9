#!/usr/bin/env python
"""Summarise an ncu --set full report of the chain kernel (run where ncu is installed, no GPU
needed):  python tools/ncu_summary.py gpurun_out/prof.ncu-rep <steps in the launch> > profiles/...txt
Also writes profiles/ncu_traffic.json (read by bench.py for roofline.traffic)."""
import csv, io, json, os, subprocess, sys
rep, steps = sys.argv[1], float(sys.argv[2])
tag = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def f(k):
    return float(m[k][0].replace(",", ""))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_tensor.sum"]
print("ncu --set full --clock-control none: %s %s" % (rep, tag))
for k in keys:
    if k in m:
        print("%-68s %-16s %s" % (k, m[k][1], m[k][0]))
# warp-state samples (why warps do not issue)
st = [(k, f(k)) for k in m if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")]
tot = sum(v for _, v in st) or 1.
print("\nwarp-state samples (share of all samples):")
for k, v in sorted(st, key=lambda x: -x[1])[:10]:
    print("  %-60s %5.1f %%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot))
unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.}
dram = f("dram__bytes_read.sum") * unit[m["dram__bytes_read.sum"][1]] + f("dram__bytes_write.sum") * unit[m["dram__bytes_write.sum"][1]]
tms = f("gpu__time_duration.sum") * {"ms": 1., "us": 1e-3, "s": 1e3, "ns": 1e-6}[m["gpu__time_duration.sum"][1]]
print("\nMC steps in this launch: %.0f  -> %.3e steps/s, %.0f DRAM bytes per MC step, %.2f TB/s" % (
    steps, steps / (tms * 1e-3), dram / steps, dram / (tms * 1e-3) * 1e-12))
if "smsp__inst_executed_pipe_xu.sum" in m:
    print("XU (MUFU + conversions) warp instructions per MC step: %.0f (x32 lanes = %.0f lane-ops)" % (
        f("smsp__inst_executed_pipe_xu.sum") / steps, 32 * f("smsp__inst_executed_pipe_xu.sum") / steps))
out = {"kernel": m["Kernel Name"][0], "dram_bytes_per_launch": dram, "mc_steps_per_launch": steps,
       "dram_bytes_per_mc_step": dram / steps, "launch_ms": tms,
       "source": "profiles/%s (ncu --set full --clock-control none, one steady-state slice of 148 chains x 8192 steps, "
                 "tools/prof_one.py)" % os.path.basename(sys.argv[4]) if len(sys.argv) > 4 else rep}
if len(sys.argv) > 4:
    # (the steady-state slice with every SM busy; tools/launch_traffic.py merges it into ncu_traffic.json)
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(sys.argv[4])), "ncu_saturated_slice.json"), "w"), indent=1)
