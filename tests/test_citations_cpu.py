"""Every reference citation (file.py:line[-line]) in the public header, the kernels and the Python mirrors points at an existing
file of the reference and at lines inside it.  Needs /root/reference (skipped on the GPU box)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ROVER_REFERENCE_ROOT", "/root/reference")
CITE = re.compile(r"(?<![\w/.])((?:\.\./)?(?:[\w]+/)*[\w]+\.(?:py|yaml|sh)):(\d+)(?:-(\d+))?")


def _index():
    files = {}
    for dirpath, _, names in os.walk(os.path.join(REF, "omniisaacgymenvs")):
        for n in names:
            if n.endswith((".py", ".yaml", ".sh")):
                files.setdefault(n, []).append(os.path.join(dirpath, n))
    for n in ("setup.py",):
        if os.path.exists(os.path.join(REF, n)):
            files.setdefault(n, []).append(os.path.join(REF, n))
    return files


def _sources():
    out = [os.path.join(ROOT, "include", "rover_b200.h"), os.path.join(ROOT, "DESIGN.md"), os.path.join(ROOT, "INTEGRATION.md")]
    for sub in ("isaac_rover_2.0_b200", os.path.join("isaac_rover_2.0_b200", "csrc"), "oracle"):
        d = os.path.join(ROOT, sub)
        out += [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".py", ".cu", ".cuh", ".c"))]
    return out


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "omniisaacgymenvs")), reason="reference tree not present")
def test_reference_citations_resolve():
    files = _index()
    own = {os.path.basename(p) for p in _sources()} | {"bench.py", "make_golden.py", "make_policy_golden.py", "make_reset_golden.py"}
    n_checked, bad = 0, []
    for src in _sources():
        text = open(src).read()
        for m in CITE.finditer(text):
            path, lo, hi = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            base = os.path.basename(path)
            cands = files.get(base)
            if not cands:
                if base in own:
                    continue                      # a citation of this repository's own file
                bad.append("%s: %s (no such file in the reference)" % (os.path.relpath(src, ROOT), m.group(0)))
                continue
            tail = path.replace("../", "")
            narrowed = [c for c in cands if c.endswith("/" + tail)] or cands
            n_lines = max(sum(1 for _ in open(c, errors="replace")) for c in narrowed)
            n_checked += 1
            if not (1 <= lo <= hi <= n_lines):
                bad.append("%s: %s (file has %d lines)" % (os.path.relpath(src, ROOT), m.group(0), n_lines))
    assert n_checked > 150, n_checked
    assert not bad, "\n".join(bad)
