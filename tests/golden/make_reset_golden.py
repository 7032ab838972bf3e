"""Generates tests/golden/reset_golden.pt with the UNMODIFIED reference (run in the container that has /root/reference):
RoverTask.random_goals / check_goal_collision / get_pos_height (rover.py:533-564, 588-608) fed with known uniform numbers
(torch.rand is replaced for the call by a function returning them), so that the goal arithmetic of oracle/reset_oracle.py and
of rvb_reset_targets is pinned to the reference's.      python tests/golden/make_reset_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import ref_import          # noqa: E402
import reset_oracle as RO  # noqa: E402


def main():
    ns = ref_import.load("cpu")
    RT = ns.RoverTask
    g = torch.Generator().manual_seed(7)
    M, S, H = 24, 20, 400            # <= 25 rows: torch.cdist takes its direct path (the formulation rvb_reset_targets uses)
    initial = torch.rand(M, 3, generator=g) * 6 + 2
    stones6 = torch.zeros(S, 6)
    stones6[:, :2] = torch.rand(S, 2, generator=g) * 26 - 8
    stones6[:, 3:5] = torch.rand(S, 2, generator=g) * 2.8 + 0.2
    stones = torch.cat((stones6, (torch.maximum(stones6[:, 3], stones6[:, 4]) / 4).unsqueeze(1)), 1)      # terrain_utils.py:416-424
    hm = torch.rand(H, H, generator=g)
    u = torch.from_numpy(RO.uniform(42, 5, np.arange(M) + 1000, 0))
    fake = types.SimpleNamespace(_device="cpu", target_positions=torch.zeros(M, 3), initial_pos=initial, stone_info=stones)
    ids = torch.arange(M)
    real_rand = torch.rand
    torch.rand = lambda n, device=None: u.clone()
    try:
        RT.random_goals(fake, ids, radius=8)
    finally:
        torch.rand = real_rand
    e2, cnt = RT.check_goal_collision(fake, ids)
    flags = (torch.cdist(fake.target_positions[:, 0:2], stones[:, 0:2], p=2.0) - stones[:, 6]).min(1)[0] <= 1.0
    hz = RT.get_pos_height(fake, hm, fake.target_positions[:, 0:2], 0.05, 1, torch.tensor([-2.0, -3.0]))
    torch.save(dict(seed=42, epoch=5, env_offset=1000, u=u, initial=initial, stones=stones, heightmap=hm, hscale=0.05, vscale=1,
                    shift=torch.tensor([-2.0, -3.0]), ref_target_xy=fake.target_positions[:, :2].clone(), ref_invalid=flags,
                    ref_invalid_count=cnt, ref_height=hz), os.path.join(HERE, "reset_golden.pt"))
    print("invalid goals:", cnt, "of", M)


if __name__ == "__main__":
    main()
