"""Generates tests/golden/spawn_direct_golden.pt: RoverTask.avoid_pos_rock_collision (rover.py:649-661) and the nearest stone edge
of check_goal_collision (rover.py:536-538) run by the UNMODIFIED reference on inputs small enough (<= 25 positions, <= 25 stones)
for torch.cdist to take its DIRECT path -- the only formulation whose result does not depend on a GEMM's summation order, hence
the one on which spawn / goal validity can be demanded bit for bit.

Run in the CPU container only:   python tests/golden/make_spawn_golden.py
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import ref_import  # noqa: E402
import rover_oracle as O  # noqa: E402


def main():
    ns = ref_import.load("cpu")
    RT = ns.RoverTask
    gen = torch.Generator().manual_seed(2024)
    cases = []
    for n_pos, n_stones, extent in ((24, 5, 8.0), (25, 25, 20.0), (1, 3, 6.0), (7, 12, 10.0)):
        s6 = torch.rand(n_stones, 6, generator=gen)
        s6[:, 0:2] *= extent
        s6[:, 3:6] = 0.2 + 2.8 * s6[:, 3:6]
        stone7 = O.read_stone_info(s6.numpy())
        pos = torch.rand(n_pos, 3, generator=gen) * extent
        # half of the positions start inside a stone's clearance so that the loop has work to do
        k = torch.randint(0, n_stones, (n_pos,), generator=gen)
        near = stone7[k, 0:2] + (torch.rand(n_pos, 2, generator=gen) - 0.5) * 1.5
        pos[::2, 0:2] = near[::2]
        fake = types.SimpleNamespace(stone_info=stone7)
        ref_pos = RT.avoid_pos_rock_collision(fake, pos.clone())
        d = torch.cdist(pos[:, 0:2], stone7[:, 0:2], p=2.0)
        d[:] = d[:] - stone7[:, 6]
        cases.append(dict(in_pos=pos, stone7=stone7, ref_pos=ref_pos, ref_nearest=torch.min(d, dim=1)[0]))
        print(n_pos, n_stones, "moved rows", int((ref_pos[:, 0] != pos[:, 0]).sum()), "max steps",
              float(((ref_pos[:, 0] - pos[:, 0]) / 0.05).max()))
    out = os.path.join(HERE, "spawn_direct_golden.pt")
    torch.save({"note": "reference outputs, torch %s CPU, cdist direct path" % torch.__version__, "cases": cases}, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
