"""Turn gpurun_out/{launches.csv,prof_sweeps.ncu-rep} into the tracked summaries under profiles/."""
import csv, json, os, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
# ---- launch list: per-kernel totals and share of one iteration ----
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", "launches.csv"))) if len(r) > 5]
hdr = rows[0]
iN, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[iV].replace(",", ""))
    except ValueError:
        continue
    k = r[iN].split("(")[0]
    agg[k][0] += 1
    agg[k][1] += v
unit = rows[1][hdr.index("Metric Unit")]
tot = sum(v for _, v in agg.values())
with open(os.path.join(out, tag + "_launches_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 200 python scripts/profile_target.py c2 3\n")
    f.write("(3 un-graphed fused iterations of the bench workload; per-launch times are cold-cache and serialised)\n")
    f.write("%-48s %8s %14s %7s\n" % ("kernel", "launches", "total " + unit, "share"))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-48s %8d %14.1f %6.1f%%\n" % (k[:48], n, v, 100 * v / tot))
os.replace(os.path.join(ROOT, "gpurun_out", "launches.csv"), os.path.join(out, tag + "_launches.csv")) if False else None
import shutil
shutil.copy(os.path.join(ROOT, "gpurun_out", "launches.csv"), os.path.join(out, tag + "_launches.csv"))
# ---- full capture: key raw metrics of the sweep kernels ----
rep = os.path.join(ROOT, "gpurun_out", "prof_sweeps.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__cluster_dim_x", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__icc_request_hit_rate.pct",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle"]
traffic = {}
with open(os.path.join(out, tag + "_sweeps_ncu_full.txt"), "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k 'regex:cluster_(fwd|bwd)_kernel|rollout_' -s 2 -c 2 python scripts/profile_target.py c2 2\n")
    f.write("units row: " + ", ".join("%s=%s" % (h, rows[1][hdr.index(h)]) for h in want[1:8] if h in hdr) + "\n\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        f.write("== %s\n" % name)
        for w in want[1:]:
            if w in hdr:
                f.write("   %-72s %s\n" % (w, r[hdr.index(w)]))
        rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
        mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[rows[1][hdr.index("dram__bytes_read.sum")]]
        traffic[name.split("<")[0].split("(")[0].strip().replace("void ", "")] = (rd + wr) * mult
        f.write("\n")
json.dump(traffic, open(os.path.join(out, tag + "_dram_traffic.json"), "w"), indent=1)
print(open(os.path.join(out, tag + "_launches_summary.txt")).read())
print(traffic)
