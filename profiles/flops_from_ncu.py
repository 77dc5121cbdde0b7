"""profiles/flops_from_ncu.py <csv> [<csv> ...] -- FLOP per vehicle-step from ncu op-counter csv files (flop_count.py runs).

Each csv holds one launch of n vehicles x ticks (taken from the file name: flops_<math>_<prec>_<mode>[_hk]_<n>x<ticks>.csv, or
4096x300 for the round-1 files).  Prints a json fragment for profiles/flops.json."""
import csv
import json
import os
import re
import sys

out = {}
for path in sys.argv[1:]:
    name = os.path.basename(path)
    m = re.match(r"flops_(.+?)(?:_(\d+)x(\d+))?\.csv$", name)
    key = m.group(1).replace("uwb", "full")
    n, ticks = (int(m.group(2)), int(m.group(3))) if m.group(2) else (4096, 300)
    vals = {}
    with open(path) as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    hdr = rows[0]
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        vals[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    g = lambda op: vals.get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % op, 0.0)
    steps = float(n * ticks)
    # packed FP32 (sm_100 FADD2 / FMUL2 / FFMA2): one thread-instruction carries two lanes
    packed = 2 * g("fadd2") + 2 * g("fmul2") + 4 * g("ffma2")
    out[key] = {"fp32": round((g("fadd") + g("fmul") + 2 * g("ffma") + packed) / steps, 1),
                "fp32_packed": round(packed / steps, 1),
                "fp64": round((g("dadd") + g("dmul") + 2 * g("dfma")) / steps, 1),
                "warp_inst_per_vehicle_tick": round(vals.get("smsp__inst_executed.sum", 0.0) * 32 / steps, 1) if "smsp__inst_executed.sum" in vals else None,
                "source": "profiles/r2/" + name}
print(json.dumps(out, indent=1))
