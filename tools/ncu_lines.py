#!/usr/bin/env python
"""Per-source-line profile of one kernel: joins the SASS page of an .ncu-rep (instructions executed, stall samples per
SASS instruction) with the line table of the kernel's cubin (nvdisasm -g), read here without a GPU.
usage: python tools/ncu_lines.py REP.ncu-rep KERNEL_SUBSTR [LIB.so] [top N]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                             "isaac_rover_2.0_b200", "librover_b200.so")
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    kre = sys.argv[5] if len(sys.argv) > 5 else kern          # kernel-name regex for ncu (demangled), kern matches the cubin symbol
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    lines = None
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin") or "-" in f.split(".")[0]:
            continue
        out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kern not in out:
            continue
        cur, sect, lines = None, False, []
        for ln in out.splitlines():
            if ln.startswith("//--------------------- .text."):
                sect = kern in ln
                continue
            if ln.startswith("//--------------------- ") and not ln.startswith("//--------------------- .text."):
                sect = False
            if not sect:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                lines.append(cur)
        if lines:
            break
    if not lines:
        sys.exit("kernel not found in " + lib)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[1]
    ie, isamp = h.index("Instructions Executed"), h.index("# Samples")
    data = [(int(r[ie]), int(r[isamp])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
    nl = len(lines)
    launches = max(1, round(len(data) / nl))
    print("# %s: %d SASS instructions in the cubin, %d rows in the report (%d launches)" % (kern, nl, len(data), launches))
    agg = {}
    for i, (e, s) in enumerate(data):
        key = lines[i % nl] or ("?", 0)
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += e
        a[1] += s
        a[2] += 1
    ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print("# total: %.1f M warp instructions, %d samples per launch" % (ti / launches / 1e6, ts // launches))
    print("%-24s %6s %8s %8s %6s" % ("file:line", "sass", "inst %", "samp %", ""))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-24s %6d %8.2f %8.2f" % ("%s:%d" % key, a[2] // launches, 100.0 * a[0] / ti, 100.0 * a[1] / max(ts, 1)))
    # by file / 25-line bucket
    print("\n# by 20-line bucket")
    b = {}
    for (f, l), a in agg.items():
        k = (f, l // 20 * 20)
        x = b.setdefault(k, [0, 0, 0])
        x[0] += a[0]; x[1] += a[1]; x[2] += a[2]
    for key, a in sorted(b.items()):
        if a[0] > 0.004 * ti or a[1] > 0.004 * ts:
            print("%-24s %6d %8.2f %8.2f" % ("%s:%d+" % key, a[2] // launches, 100.0 * a[0] / ti, 100.0 * a[1] / max(ts, 1)))


if __name__ == "__main__":
    main()
