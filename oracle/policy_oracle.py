"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's policy / value forward pass.

Follows learning/model.py: Layer = nn.Linear + activation (model.py:86-116), Encoder = a chain of Layers (model.py:118-150),
StochasticActorHeightmap.compute (model.py:183-192) and DeterministicHeightmap.compute (model.py:231-241).  Parameters come
as a dict with the reference's state-dict keys.  Pinned: tests/test_policy_cpu.py runs the UNMODIFIED reference classes
(oracle/ref_import.py) on the same parameters and checks this restatement bit for bit in fp32; the committed fixture
tests/golden/policy_golden.pt holds parameters, observations and the reference's outputs (tests/golden/make_policy_golden.py).
Never imported by the product.
"""
import torch
import torch.nn.functional as F

ACTIVATIONS = {                                    # model.py:104-111
    "elu": F.elu, "relu": F.relu, "leakyrelu": F.leaky_relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh, "relu6": F.relu6,
}


def _chain(x, sd, prefix, n_layers, act):
    for i in range(n_layers):                      # Encoder.forward / the network loop, model.py:146-149,190-191
        x = act(F.linear(x, sd["%s.%d.layer.0.weight" % (prefix, i)], sd["%s.%d.layer.0.bias" % (prefix, i)]))
    return x


def forward(sd, states, num_proprioception, num_sparse, num_dense, activation="leakyrelu", actor=True, dtype=torch.float32,
            n_enc=2, n_mlp=3):
    """states [N, >= p+S+D] -> tanh(mean) [N, A] (actor) or value [N, 1] (critic), computed in `dtype`."""
    sd = {k: v.to(dtype) for k, v in sd.items() if k != "log_std_parameter"}
    states = states.to(dtype)
    act = ACTIVATIONS[activation]
    p = num_proprioception
    sparse = states[:, p:p + num_sparse]                                    # model.py:184
    dense = states[:, p + num_sparse:p + num_sparse + num_dense]            # model.py:185
    x0 = _chain(sparse, sd, "encoder0.encoder", n_enc, act)                 # model.py:186
    x1 = _chain(dense, sd, "encoder1.encoder", n_enc, act)                  # model.py:187
    x = torch.cat((states[:, 0:p], x0), dim=1)                              # model.py:188
    x = torch.cat((x, x1), dim=1)                                           # model.py:189
    x = _chain(x, sd, "network", n_mlp, act)
    x = F.linear(x, sd["network.%d.weight" % n_mlp], sd["network.%d.bias" % n_mlp])   # model.py:176 / 229
    return torch.tanh(x) if actor else x                                    # model.py:177
