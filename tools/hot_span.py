"""Address span of the shadow kernel's dispatcher loop in the built library (no GPU): how much code the loop's range holds.
usage: python tools/hot_span.py   (after a build)"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "isaac_rover_2.0_b200", "csrc", "raycast_shadow.cu")).read().split("\n")


def find(s):
    for i, l in enumerate(src):
        if s in l:
            return i + 1
    raise SystemExit("marker not found: " + s)


loop, p4 = find("    while (true) {"), find("// ---- phase 4")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "isaac_rover_2.0_b200", "librover_b200.so")], cwd=tmp, capture_output=True)
out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, "raycast_shadow.sm_100a.cubin")], capture_output=True, text=True).stdout
kern = sys.argv[1] if len(sys.argv) > 1 else "hm_shadow_kernelILb0ELi1664"
sect, cur, recs = False, None, []
for ln in out.splitlines():
    if ln.startswith("//--------------------- .text."):
        sect = kern in ln
        continue
    if not sect:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", ln)
    if m and cur:
        recs.append((int(m.group(1), 16), cur[0], cur[1]))
hot = [r for r in recs if r[1] == "raycast_shadow.cu" and loop <= r[2] < p4]
lo, hi = min(r[0] for r in hot), max(r[0] for r in hot)
inside = [r for r in recs if lo <= r[0] <= hi]
print("kernel %s: %d instructions (%.1f KB); dispatcher loop spans %.1f KB holding %d instructions" % (
    kern, len(recs), len(recs) * 16 / 1024, (hi - lo + 16) / 1024, len(inside)))
agg = {}
for r in inside:
    agg[r[1]] = agg.get(r[1], 0) + 1
print("  by file:", sorted(agg.items(), key=lambda kv: -kv[1]))
