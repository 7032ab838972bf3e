import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.getcwd())
import isaac_rover_b200 as R
torch.manual_seed(0)
net = R.model.NetworkInfo([256, 160, 128], [80, 60], [80, 60], [80, 60], "leakyrelu")
info = R.model.ObserverationInfo(4, 634, 1112, 0)
actor = R.model.StochasticActorHeightmap(1750, 2, net, info, device="cuda:0")
critic = R.model.DeterministicHeightmap(1750, 2, net, info, device="cuda:0")
N = 8192
obs = torch.rand(N, 1750, device="cuda")
obs[:, 4:] = (obs[:, 4:] * 5.5).half().float()
m, v = R.model.compute_pair(actor, critic, obs)
def eager(sd, tanh):
    def chain(x, prefix, n):
        for i in range(n):
            x = F.leaky_relu(F.linear(x, sd["%s.%d.layer.0.weight" % (prefix, i)], sd["%s.%d.layer.0.bias" % (prefix, i)]))
        return x
    x = torch.cat((obs[:, 0:4], chain(obs[:, 4:638], "encoder0.encoder", 2), chain(obs[:, 638:1750], "encoder1.encoder", 2)), dim=1)
    x = F.linear(chain(x, "network", 3), sd["network.3.weight"], sd["network.3.bias"])
    return torch.tanh(x) if tanh else x
sa = {k: t.cuda() for k, t in actor.state_dict().items()}
sc = {k: t.cuda() for k, t in critic.state_dict().items()}
ea, ec = eager(sa, True), eager(sc, False)
print("mean sample", m[:3].tolist(), "eager", ea[:3].tolist())
print("value sample", v[:3].flatten().tolist(), "eager", ec[:3].flatten().tolist())
print("max diff mean %.3e value %.3e; |mean| max %.3f, value range %.3f..%.3f" % ((m - ea).abs().max().item(), (v - ec).abs().max().item(), m.abs().max().item(), v.min().item(), v.max().item()))
