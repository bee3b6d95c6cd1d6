#!/usr/bin/env python
"""Regenerate profiles/r2_* from the scratch captures in gpurun_out/ (tools/capture_round2.sh; see profiles/README.md)."""
import collections, csv, json, os, shutil, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + "/"
out = {}
for tag, rep in (("f64_4096_T100_rows (bench workload, contact on)", "prof_r2_f64_4096"), ("f64_131072_T2", "prof_r2_f64_131072"),
                 ("f32_131072_T2", "prof_r2_f32_131072"), ("f64_32envs_group18_T100 (the slowest team of the bench batch, alone)", "prof_r2_group18")):
    p = R + f"gpurun_out/{rep}.ncu-rep"
    if not os.path.exists(p):
        continue
    js = subprocess.run([sys.executable, R + "tools/ncu_summary.py", p], capture_output=True, text=True).stdout
    out[tag] = json.loads(js)[0]
json.dump(out, open(R + "profiles/r2_ncu_full_summary.json", "w"), indent=1)
d = out["f64_4096_T100_rows (bench workload, contact on)"]
def val(k):
    v, u = d[k]
    return float(v) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
json.dump({"kernel": "k_rollout<double,double,SPLIT=true> (team kernel, table contact simulated)",
           "config": "4096 envs x 100 control steps x 10 sub-steps, rows written (bench workload)",
           "dram_bytes_per_launch_config1": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
           "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
           "algorithmic_bytes_per_launch": 4096 * 100 * 208, "source": "ncu --set full, profiles/r2_ncu_full_summary.json"},
          open(R + "profiles/r2_kernel_traffic.json", "w"), indent=1)
rows = [r for r in csv.reader(open(R + "gpurun_out/r2_bench_launches.csv")) if len(r) > 5]
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    agg[r[ci["Kernel Name"]][:90]][0] += 1
    agg[r[ci["Kernel Name"]][:90]][1] += float(r[ci["Metric Value"]]) / 1e6
tot = sum(v[1] for v in agg.values())
with open(R + "profiles/r2_bench_launch_list_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 --no-extra\n")
    f.write("(per-launch times are cold-cache and serialised: compare shares)\n\n launches   total_ms   share  kernel\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{n:9d} {t:10.3f} {100*t/tot:6.1f}%  {k}\n")
shutil.copy(R + "gpurun_out/r2_bench_launches.csv", R + "profiles/r2_bench_launches.csv")
for src, dst in (("r2_bench_default.json", "r2_bench_default.json"), ("r2_bench_reference.json", "r2_bench_reference.json"),
                 ("r2_contact_perf.log", "r2_contact_perf.txt"), ("r2_mpc_frame.log", "r2_mpc_frame.txt")):
    if os.path.exists(R + "gpurun_out/" + src):
        shutil.copy(R + "gpurun_out/" + src, R + "profiles/" + dst)
print(open(R + "profiles/r2_bench_launch_list_summary.txt").read())
for k, v in out.items():
    print(k, {kk: v[kk][0] for kk in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
                                       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active") if kk in v}, v.get("stall_per_issue"))
