"""The culling stages of the production ray-cast kernel (csrc/raycast_shadow.cu) must be CONSERVATIVE: no (ray, triangle)
pair that passes the packed-fp16 pre-filter of ray_casting.py:34-59 may be dropped.  tests/shadow_proto.py re-states
stages 1-2 (same formulas and constants, numpy fp32) and checks them against a brute-force fp16 evaluation."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("args", [
    ["--envs", "3", "--length", "200", "--nv", "708", "--seed", "1", "--wild", "0.34"],      # fp16 grid 0.125 m
    ["--envs", "3", "--length", "40", "--nv", "142", "--seed", "2", "--wild", "0.34"],
])
def test_shadow_bounds_are_conservative(args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "shadow_proto.py")] + args, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "violations 0\n" in r.stdout and "'s1_viol': 0.0" in r.stdout and "'gball_viol': 0.0" in r.stdout, r.stdout[-2000:]


def test_constants_match_kernel():
    src = "".join(open(os.path.join(ROOT, "isaac_rover_2.0_b200", "csrc", f)).read() for f in ("shadow_bounds.cuh", "raycast_shadow.cu"))
    for needle in ("GAMMA = 0.00390625f", "ALPHA = 1.9073486328125e-06f", "EPS0 = 0.1057f", "L_CAP = 64.0f", "OVF = 16000.0f",
                   "0x2E68u", "0x3C6Bu"):
        assert needle in src, needle
