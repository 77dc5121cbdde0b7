#!/usr/bin/env python
"""Join an ncu report's per-instruction SASS page with nvdisasm line info and aggregate the executed
instructions and stall samples per source line / per named source region.

    python profiles/ncu_by_line.py <report.ncu-rep> <object.o|.so> <mangled-kernel-substring> [--top N]

Works on the CPU box (ncu -i, cuobjdump, nvdisasm).  The listing order of the kernel's SASS is the same
in both tools, so rows are matched by instruction offset."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

def function_spans(src_path):
    """very small C++ scanner: maps each line of the .cuh to the name of the enclosing function"""
    lines = open(src_path).read().split("\n")
    owner = [None] * (len(lines) + 2)
    cur, depth, pending = None, 0, None
    sig = re.compile(r"^\s*(?:static\s+)?(?:template<[^>]*>\s*)?(?:AGF_DEV|AGFR_DEV|AGF_COLD|static AGF_COLD|__global__|__device__(?:\s+__noinline__)?|AGF_HDI|static AGF_DEV|static __device__ __noinline__)[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    for i, ln in enumerate(lines, 1):
        if depth == 0:
            m = sig.match(ln)
            if m:
                pending = m.group(1)
            elif re.match(r"^step_kernel\(", ln):
                pending = "step_kernel"
        if pending and cur is None and "{" in ln:
            cur = pending
            pending = None
        if cur:
            owner[i] = cur
        if ln.startswith("namespace") or ln.startswith("}  // namespace"):
            continue
        depth += ln.count("{") - ln.count("}")
        if cur and depth <= 0:
            cur, depth = None, 0
    return owner, lines


def main():
    rep, obj, kern = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
    dis = ""
    for c in cub:
        dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, c)], capture_output=True, text=True).stdout
    # instruction offset -> (file, line) for the kernel
    off2line = {}
    inside, cur = False, (None, 0)
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
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            off2line[int(m.group(1), 16)] = cur
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.split("\n")))
    hdr = rows[1]
    ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    per_line = defaultdict(lambda: [0, 0])
    per_line_stall = defaultdict(lambda: defaultdict(int))
    tot_ex = tot_smp = 0
    opcount = defaultdict(int)
    for r in rows[2:]:
        if len(r) <= iex or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        key = off2line.get(a - base, ("?", 0))
        ex, smp = int(r[iex] or 0), int(r[ismp] or 0)
        per_line[key][0] += ex
        per_line[key][1] += smp
        tot_ex += ex
        tot_smp += smp
        op = r[isrc].split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        opcount[op.split(".")[0]] += ex
        for i, h in stall_cols:
            if r[i] and r[i] != "0":
                per_line_stall[key][h] += int(r[i])
    print("kernel %s: %d warp-instructions executed, %d stall samples" % (kern, tot_ex, tot_smp))
    # per function
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "agri-fly_b200", "csrc")
    owners = {}
    for name in ("agf_step.cuh", "agf_rappids_plan.cuh", "agf_types.h", "agf_math.h"):
        owners[name] = function_spans(os.path.join(csrc, name))[0]
    per_fn = defaultdict(lambda: [0, 0])
    per_fn_stall = defaultdict(lambda: defaultdict(int))
    for (f, l), (ex, smp) in per_line.items():
        owner = owners.get(f)
        fn = owner[l] if owner is not None and l < len(owner) and owner[l] else f + ":other"
        per_fn[fn][0] += ex
        per_fn[fn][1] += smp
        for h, v in per_line_stall[(f, l)].items():
            per_fn_stall[fn][h] += v
    print("\n%-28s %8s %8s   top stalls" % ("function", "inst %", "samp %"))
    for fn, (ex, smp) in sorted(per_fn.items(), key=lambda kv: -kv[1][0]):
        st = sorted(per_fn_stall[fn].items(), key=lambda kv: -kv[1])[:4]
        print("%-28s %7.2f%% %7.2f%%   %s" % (fn, 100.0 * ex / max(tot_ex, 1), 100.0 * smp / max(tot_smp, 1),
                                              ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(tot_smp, 1)) for h, v in st)))
    print("\nopcode mix (warp-instructions):")
    for op, c in sorted(opcount.items(), key=lambda kv: -kv[1])[:24]:
        print("  %-10s %6.2f%%" % (op, 100.0 * c / max(tot_ex, 1)))
    print("\ntop %d source lines by stall samples:" % top)
    for (f, l), (ex, smp) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        st = sorted(per_line_stall[(f, l)].items(), key=lambda kv: -kv[1])[:3]
        print("  %s:%-5d inst %5.2f%% samp %5.2f%%  %s" % (f, l, 100.0 * ex / max(tot_ex, 1), 100.0 * smp / max(tot_smp, 1),
                                                            ", ".join("%s %d" % (h[6:], v) for h, v in st)))


if __name__ == "__main__":
    main()
