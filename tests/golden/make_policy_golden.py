"""Generates tests/golden/policy_golden.pt by running the UNMODIFIED reference models (learning/model.py) on the CPU.

Run in the build container (needs /root/reference):  python tests/golden/make_policy_golden.py
skrl / gym are absent: oracle/ref_import.py stubs them; the stub `Model` carries no state, so `num_actions` (which skrl's
Model.__init__ would set from the action space) is supplied as a class attribute.
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def reference_models(num_obs=1750, num_sparse=634, num_dense=1112, activation="leakyrelu", seed=0):
    ref_import.install()
    from omniisaacgymenvs.learning import model as M

    class Actor(M.StochasticActorHeightmap):
        num_actions = 2

    class Critic(M.DeterministicHeightmap):
        num_actions = 2

    torch.manual_seed(seed)
    obs_space = types.SimpleNamespace(shape=(num_obs,))
    act_space = types.SimpleNamespace(shape=(2,))
    net = M.NetworkInfo([256, 160, 128], [80, 60], [80, 60], [80, 60], activation)      # train.py:95
    info = M.ObserverationInfo(4, num_sparse, num_dense, 0)
    return Actor(obs_space, act_space, net, info, device="cpu"), Critic(obs_space, act_space, net, info, device="cpu")


def state_dict_of(m):
    sd = {}
    for name in ("encoder0", "encoder1", "network"):
        for k, v in getattr(m, name).state_dict().items():
            sd["%s.%s" % (name, k)] = v.detach().clone()
    if isinstance(getattr(m, "log_std_parameter", None), torch.Tensor):
        sd["log_std_parameter"] = m.log_std_parameter.detach().clone()
    return sd


def synthetic_obs(n, num_obs, seed):
    g = torch.Generator().manual_seed(seed)
    obs = torch.empty(n, num_obs)
    obs[:, 0] = torch.rand(n, generator=g) * 1.2                    # target distance / 9
    obs[:, 1] = torch.rand(n, generator=g) * 2 - 1                  # heading / pi
    obs[:, 2:4] = torch.rand(n, 2, generator=g) * 2 - 1             # last actions
    h = (torch.rand(n, num_obs - 4, generator=g) * 0.6 + 0.1).half().float() / 2      # fp16 ray distances / 2 (rover.py:324-325)
    miss = torch.rand(n, num_obs - 4, generator=g) < 0.02
    obs[:, 4:] = torch.where(miss, torch.tensor(5.5), h)            # 11.0 / 2: the miss sentinel
    return obs


def main():
    out = {}
    for tag, act in (("leakyrelu", "leakyrelu"),):       # the configured activation (cfg/trainSKRL/RoverPPOSKRL.yaml:5,9)
        actor, critic = reference_models(activation=act, seed=7)
        obs = synthetic_obs(70, 1750, seed=11)                      # 70 = two full 32-env tiles + a ragged one
        with torch.no_grad():
            mean, log_std = actor.compute(obs, None, "policy")
            value = critic.compute(obs, None, "value")
        out[tag] = {"activation": act, "obs": obs, "actor_sd": state_dict_of(actor), "critic_sd": state_dict_of(critic),
                    "mean": mean.clone(), "log_std": log_std.detach().clone(), "value": value.clone()}
    out["torch_version"] = str(torch.__version__)
    path = os.path.join(ROOT, "tests", "golden", "policy_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
