#!/usr/bin/env python
"""Static code size of a kernel per source function (instructions in the SASS, attributed through nvdisasm line info).
    python profiles/code_size_by_fn.py <object.o> <mangled-kernel-substring>"""
import os, re, subprocess, sys, tempfile
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_by_line import function_spans

obj, kern = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
dis = ""
for c in os.listdir(tmp):
    if c.endswith(".cubin"):
        dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, c)], capture_output=True, text=True).stdout
csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "agri-fly_b200", "csrc")
owners = {n: function_spans(os.path.join(csrc, n))[0] for n in ("agf_step.cuh", "agf_rappids_plan.cuh", "agf_types.h", "agf_math.h")}
inside, cur, cnt, total = False, ("?", 0), defaultdict(int), 0
for ln in dis.split("\n"):
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln):
        f, l = cur
        o = owners.get(f)
        cnt[o[l] if o is not None and l < len(o) and o[l] else f + ":other"] += 1
        total += 1
print("%s: %d instructions = %.1f KB" % (kern, total, total * 16 / 1024))
for fn, c in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
    print("  %-32s %6d  %5.1f%%" % (fn, c, 100.0 * c / total))
