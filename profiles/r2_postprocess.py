#!/usr/bin/env python
"""Turns what profiles/r2_final.sh + r2_flops2.sh left under gpurun_out/r2/ into the committed evidence under profiles/:
bench lines, test logs, counters -> profiles/flops.json, DRAM traffic -> profiles/r2/traffic.json, ncu --set full summaries."""
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out", "r2"), os.path.join(ROOT, "profiles", "r2")


def metrics(path):
    with open(path) as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    h = rows[0]
    return {d["Metric Name"]: float(d["Metric Value"].replace(",", "")) for d in (dict(zip(h, r)) for r in rows[1:])}


for pat in ("bench_*final*.json", "gpu_tests_final.log", "launches_final.csv", "flops_fast_*_4096x300.csv", "traffic_c3_fp32_fast_hk_131072x500.csv",
            "traffic_c4_fp32_rates_2097152x300.csv", "traffic_c4_fp32_full_2097152x200.csv", "smoke_final.log", "ncu_metric_names.txt"):
    for f in glob.glob(os.path.join(SRC, pat)):
        shutil.copy(f, DST)
# FLOP / instruction counters
fl = json.load(open(os.path.join(ROOT, "profiles", "flops.json")))
keys = ("fp32_uwb_hk", "fp32_uwb", "fp32_rates_hk", "fp32_rates", "fp64_uwb_hk", "fp64_rates_hk")
new = json.loads(subprocess.run([sys.executable, "profiles/flops_from_ncu.py"] + [os.path.join(SRC, "flops_fast_%s_4096x300.csv" % k) for k in keys],
                                capture_output=True, text=True, cwd=ROOT).stdout)
fl.update(new)
json.dump(fl, open(os.path.join(ROOT, "profiles", "flops.json"), "w"), indent=1)
for k, v in new.items():
    print(k, v)
# DRAM traffic
tr = json.load(open(os.path.join(DST, "traffic.json")))
for key, fn, n, t, what in (("c3_fp32_fast_hk", "traffic_c3_fp32_fast_hk_131072x500.csv", 131072, 500, "step_kernel<float,0,1,1,0,0> (housekeeping on): the state round trip"),
                            ("c4_fp32_rates", "traffic_c4_fp32_rates_2097152x300.csv", 2097152, 300, "step_kernel<float,0,0,1,1,0> (per-vehicle parameters), log every tick"),
                            ("c4_fp32_full", "traffic_c4_fp32_full_2097152x200.csv", 2097152, 200, "step_kernel<float,0,1,1,1,0>, log every tick")):
    m = metrics(os.path.join(SRC, fn))
    rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
    per = (rd + wr) / (n * t)
    tr[key] = dict(bytes=int(rd + wr), vehicles=n, ticks=t,
                   source="profiles/r2/%s: %.2f GB read + %.2f GB write in one %d-vehicle x %d-tick launch of %s = %.1f B per vehicle-step%s"
                          % (fn, rd / 1e9, wr / 1e9, n, t, what, per, "" if key.startswith("c3") else " against 68 B of log = %.3f x the algorithmic bytes" % (per / 68.0)))
    print(key, tr[key]["source"])
json.dump(tr, open(os.path.join(DST, "traffic.json"), "w"), indent=1)
# ncu --set full summaries
for rep, obj, kern, out, title in (
        ("prof_f32_uwb_hk_final", "agf_kernels_fast_f32_uwb_final.o", "step_kernelIfLb0ELb1ELb1ELb0ELb0", "step_f32_uwb_hk_summary.txt",
         "ncu --set full, step_kernel<float,0,1,1,0,0> (FP32 fast, full mode, housekeeping ON: the benchmarked kernel), 131072 vehicles x 200 ticks, end of round 2 (profiles/r2_final.sh)"),
        ("prof_f64_uwb_hk_final", "agf_kernels_fast_f64_uwb_final.o", "step_kernelIdLb0ELb1ELb1ELb0ELb0", "step_f64_uwb_hk_summary.txt",
         "ncu --set full, step_kernel<double,0,1,1,0,0> (FP64 plant + FP32 onboard logic, fast, full mode, housekeeping on), 65536 vehicles x 200 ticks, end of round 2"),
        ("prof_c4_rates_final", "agf_kernels_fast_f32_rates_final.o", "step_kernelIfLb0ELb0ELb1ELb1ELb0", "step_c4_rates_log_summary.txt",
         "ncu --set full, step_kernel<float,0,0,1,1,0> (C4: rates mode, per-vehicle parameters, trajectory log EVERY tick), 2097152 vehicles x 300 ticks, end of round 2")):
    with open(os.path.join(DST, out), "w") as f:
        subprocess.run([sys.executable, "profiles/ncu_summary.py", os.path.join(SRC, rep + ".ncu-rep"), os.path.join(SRC, obj), kern, title], stdout=f, cwd=ROOT)
# headline numbers
for f in ("bench_final", "bench_ref_final", "bench_fp64_final", "bench_parity_final", "bench_hkoff_final", "bench_c4_rates_final", "bench_c4_full_final"):
    try:
        d = json.loads([l for l in open(os.path.join(DST, f + ".json")) if l.startswith("{")][0])
    except Exception as e:
        print(f, "ERR", e)
        continue
    r = d.get("roofline", {})
    print(f, "value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]), {k: round(v, 4) for k, v in r.items() if k in ("frac", "frac_instrumented", "frac_executed", "traffic_over_algorithmic") and v is not None})
    for k, v in d.get("extra", {}).items():
        print("   ", k, {kk: ("%.4g" % vv) for kk, vv in v.items() if kk in ("vehicle_steps_per_s", "plans_per_s", "ms_per_launch")})
