"""Launches the policy epilogue (actor + critic, one grid) at 4096 and 65,536 envs: the target of the ncu capture
`ncu --set full --clock-control none --import-source on -k policy_forward_kernel --launch-skip 2 -c 2 ... python tools/policy_probe.py`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isaac_rover_b200 as R  # noqa: E402

net = R.model.NetworkInfo([256, 160, 128], [80, 60], [80, 60], [80, 60], "leakyrelu")
info = R.model.ObserverationInfo(4, 634, 1112, 0)
torch.manual_seed(0)
actor = R.model.StochasticActorHeightmap(1750, 2, net, info, device="cuda:0")
critic = R.model.DeterministicHeightmap(1750, 2, net, info, device="cuda:0")
small = torch.rand(4096, 1750, device="cuda")
big = torch.rand(65536, 1750, device="cuda")
for _ in range(3):
    R.model.compute_pair(actor, critic, small)
for _ in range(2):
    R.model.compute_pair(actor, critic, big)
torch.cuda.synchronize()
print("ok")
