#!/usr/bin/env python
"""Stage-wise roll-up of tools/ncu_lines.py output for the shadow kernel (lines of raycast_shadow.cu grouped by phase / stage).
usage: python tools/ncu_lines.py REP KERNEL LIB 400 REGEX > lines.txt; python tools/ncu_stages.py lines.txt"""
import bisect
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for ln in open(sys.argv[1]):
    m = re.match(r'(\S+):(\d+)\s+(\d+)\s+([\d.]+)\s+([\d.]+)\s*$', ln)
    if m and not m.group(1).endswith('+'):
        rows.append((m.group(1), int(m.group(2)), int(m.group(3)), float(m.group(4)), float(m.group(5))))
src = open(os.path.join(ROOT, "isaac_rover_2.0_b200", "csrc", "raycast_shadow.cu")).read().splitlines()


def find(s):
    for i, l in enumerate(src):
        if s in l:
            return i + 1
    return 10 ** 9


marks = [('phase 0-1 (transform, cells)', find('// ---- phase 0')), ('phase 2 (sort)', find('// ---- phase 2')), ('items', find('// ---- superblock items')),
         ('phase 3 set-up', find('// ---- phase 3')), ('dispatcher', find('const uint32_t n1 = t1 - h1')), ('A1 (stage 1 loop)', find('case A1: {')),
         ('A2_START (stage 2)', find('case A2_START: {')), ('A2_EMIT (tasks)', find('case A2_EMIT: {')), ('A3L_START (prism test)', find('case A3L_START: {')),
         ('A3L_RUN (expand)', find('case A3L_RUN: {')), ('A3B (literal)', find('case A3B: {')),
         ('phase 4 (epilogue call)', find('// ---- phase 4')), ('host', find('}  // namespace'))]
s1a, s1b = find('__device__ __forceinline__ bool stage1('), find('__device__ __forceinline__ void stage2(')
s2b = find('__device__ __forceinline__ bool is_steep(')
starts = [m[1] for m in marks]
agg = {}
for f, l, sass, inst, samp in rows:
    if f == 'raycast_shadow.cu':
        if l < marks[0][1]:
            key = 'fn stage1' if s1a <= l < s1b else 'fn stage2' if s1b <= l < s2b else 'small helpers (hf, off16, tri_f, cell_coord_f)'
        else:
            key = marks[bisect.bisect_right(starts, l) - 1][0]
    else:
        key = f + ' (inlined)'
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += sass
    a[1] += inst
    a[2] += samp
print("%-50s %6s %8s %8s" % ("stage", "sass", "inst %", "samp %"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-50s %6d %8.2f %8.2f" % (k, *v))
