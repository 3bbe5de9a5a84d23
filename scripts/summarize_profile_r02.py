"""Turn gpurun_out/{launches*.csv, prof_*.ncu-rep} of scripts/gpu_profile_r02.sh into the tracked summaries under profiles/."""
import csv, json, os, shutil, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")


def launch_summary(src, dst, cmd, note):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
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
    with open(dst, "w") as f:
        f.write(cmd + "\n" + note + "\n")
        f.write("%-48s %8s %14s %7s\n" % ("kernel", "launches", "total " + unit, "share"))
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-48s %8d %14.1f %6.1f%%\n" % (k[:48], n, v, 100 * v / tot))
    print(open(dst).read())


launch_summary(os.path.join(G, "launches.csv"), os.path.join(out, tag + "_launches_summary.txt"),
               "ncu --metrics gpu__time_duration.sum --clock-control none -c 200 python scripts/profile_target.py c2 3",
               "(3 un-graphed fused iterations of the bench workload; per-launch times are cold-cache and serialised)")
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(out, tag + "_launches.csv"))
if os.path.exists(os.path.join(G, "launches_c3.csv")):
    launch_summary(os.path.join(G, "launches_c3.csv"), os.path.join(out, tag + "_launches_c3_summary.txt"),
                   "ncu --metrics gpu__time_duration.sum --clock-control none -c 60 python scripts/profile_target.py c3 2",
                   "(c3: moment matching on the cluster-resident sweeps; the cooperative cluster launches are listed when ncu can time them)")
if os.path.exists(os.path.join(G, "launches_fit.csv")):
    launch_summary(os.path.join(G, "launches_fit.csv"), os.path.join(out, tag + "_launches_fit_summary.txt"),
                   "ncu --metrics gpu__time_duration.sum --clock-control none -c 120 python scripts/fit_target.py 4",
                   "(fused train_regressor iterations, 2x[200], batch 100; torch's noise kernels included)")

want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_sleeping"]
traffic = {}
for cfg, rep, cmd in (("c2", "prof_c2", "-k 'regex:cluster_(fwd|bwd)_kernel' -s 2 -c 2 python scripts/profile_target.py c2 2"),
                      ("c4", "prof_c4", "-k 'regex:rollout_(fwd|bwd)' -s 2 -c 2 python scripts/profile_target.py c4 2"),
                      ("c5", "prof_c5", "-k 'regex:cw_(fwd|bwd)_kernel' -s 2 -c 2 python scripts/profile_target.py c5 2"),
                      ("c5tc", "prof_c5tc", "PMB_STREAM_MODE=4 ... -k 'regex:tc_(fwd|bwd)_kernel' -s 2 -c 2 python scripts/profile_target.py c5 2")):
    path = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(path):
        print("missing", path)
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        print("empty", path)
        continue
    hdr = rows[0]
    with open(os.path.join(out, "%s_%s_ncu_full.txt" % (tag, cfg)), "w") as f:
        f.write("ncu --set full --clock-control none " + cmd + "\n")
        f.write("units: " + ", ".join("%s=%s" % (h, rows[1][hdr.index(h)]) for h in want if h in hdr and rows[1][hdr.index(h)]) + "\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write("== %s\n" % name)
            for w in want:
                if w in hdr:
                    f.write("   %-72s %s\n" % (w, r[hdr.index(w)]))
            rd, wr = float(r[hdr.index("dram__bytes_read.sum")].replace(",", "")), float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            mr = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[rows[1][hdr.index("dram__bytes_read.sum")]]
            mw = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[rows[1][hdr.index("dram__bytes_write.sum")]]
            kn = name.split("<")[0].split("(")[0].strip().replace("void ", "")
            traffic["%s:%s" % (cfg, kn)] = rd * mr + wr * mw
            f.write("\n")
    print(open(os.path.join(out, "%s_%s_ncu_full.txt" % (tag, cfg))).read())
json.dump(traffic, open(os.path.join(out, tag + "_dram_traffic.json"), "w"), indent=1)
print(traffic)
for src, dst in (("bench.log", "_bench_final.json"), ("bench_ref.log", "_bench_reference_arm_final.json")):
    p = os.path.join(G, src)
    if os.path.exists(p) and os.path.getsize(p):
        shutil.copy(p, os.path.join(out, tag + dst))
