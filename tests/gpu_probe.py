"""Development probe (GPU box): which torch-CUDA semantics differ from torch-CPU on this path."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import rover_oracle as O  # noqa

dev = "cuda"
print("torch", torch.__version__, torch.cuda.get_device_name(0))
a = torch.tensor([0.7998046875, 0.80029296875, 0.7993], dtype=torch.float16)
print("half<0.8 cpu", (a < 0.8).tolist(), "cuda", (a.cuda() < 0.8).tolist())
x = (1.0 + 1.0009765625) / 2 + 1e-12
t = torch.full((64,), x, dtype=torch.float64)
print("f64->f16 cpu", t.to(torch.float16)[0].item(), "cuda", t.cuda().to(torch.float16)[0].item(), "(direct RN would be 1.0009765625)")
g = torch.Generator().manual_seed(0)
# division by python scalar
v = (torch.rand(1 << 20, generator=g) * 200).float()
for c in (0.1, 9.0, 0.2, 3000.0, 3.141592653589793, 0.025):
    cpu = v / c
    gpu = (v.cuda() / c).cpu()
    mul = v * (torch.tensor(1.0) / torch.tensor(c, dtype=torch.float32))
    print("div by %g: cuda==cpu %.6f  cuda==mul-by-recip %.6f" % (c, (cpu == gpu).float().mean(), (gpu == mul).float().mean()))
h = (torch.rand(1 << 16, generator=g) * 20).to(torch.float16)
print("half/2 cuda==cpu", torch.equal((h / 2), (h.cuda() / 2).cpu()))
# cross / normalize in half
u = (torch.rand(1 << 18, 3, generator=g) * 4 - 2).to(torch.float16)
w = (torch.rand(1 << 18, 3, generator=g) * 4 - 2).to(torch.float16)
print("half cross cuda==cpu", torch.equal(u.cross(w, dim=1), u.cuda().cross(w.cuda(), dim=1).cpu()))
nc, ng = torch.nn.functional.normalize(u), torch.nn.functional.normalize(u.cuda()).cpu()
print("half normalize cuda==cpu frac", (nc.view(torch.int16) == ng.view(torch.int16)).float().mean().item())
n1 = torch.linalg.vector_norm(u, 2, dim=1)
n2 = torch.linalg.vector_norm(u.cuda(), 2, dim=1).cpu()
uf = u.float()
for name, ss in (("(x2+y2)+z2", (uf[:, 0] * uf[:, 0] + uf[:, 1] * uf[:, 1]) + uf[:, 2] * uf[:, 2]),
                 ("x2+(y2+z2)", uf[:, 0] * uf[:, 0] + (uf[:, 1] * uf[:, 1] + uf[:, 2] * uf[:, 2])),
                 ("(x2+z2)+y2", (uf[:, 0] * uf[:, 0] + uf[:, 2] * uf[:, 2]) + uf[:, 1] * uf[:, 1])):
    r = ss.sqrt().to(torch.float16)
    print("  norm order %s: ==cpu %.6f ==cuda %.6f" % (name, (r == n1).float().mean(), (r == n2).float().mean()))
r = uf.double().pow(2).sum(1).sqrt().to(torch.float16)
print("  norm f64 exact: ==cpu %.6f ==cuda %.6f" % ((r == n1).float().mean(), (r == n2).float().mean()))
# ray_distance cpu vs cuda on identical fp16 inputs
s = (torch.rand(1 << 18, 3, generator=g) * 4 - 2).to(torch.float16)
d = (torch.rand(1 << 18, 3, generator=g) * 2 - 1).to(torch.float16)
tr = (torch.rand(1 << 18, 3, 3, generator=g) * 4 - 2).to(torch.float16)
kc, pc = O.ray_distance(s, d, tr)
kg, pg = O.ray_distance(s.cuda(), d.cuda(), tr.cuda())
print("oracle ray_distance cuda==cpu: k %.6f pt %.6f" % ((kc.view(torch.int16) == kg.cpu().view(torch.int16)).float().mean(),
                                                          (pc.view(torch.int16) == pg.cpu().view(torch.int16)).float().mean()))
# trig
ang = (torch.rand(1 << 20, generator=g) * 6.4 - 3.2).float()
for f in (torch.sin, torch.cos, torch.asin, torch.atan):
    aa = ang if f not in (torch.asin,) else ang / 3.2
    print(f.__name__, "cuda==cpu %.6f" % (f(aa) == f(aa.cuda()).cpu()).float().mean())
y = (torch.rand(1 << 20, generator=g) * 2 - 1).float()
print("atan2 cuda==cpu %.6f" % (torch.atan2(y, ang) == torch.atan2(y.cuda(), ang.cuda()).cpu()).float().mean())
# linalg.norm of 2-vectors vs sqrt(x*x+y*y)
tv = (torch.rand(1 << 20, 2, generator=g) * 16 - 8).float()
nn_c, nn_g = torch.linalg.norm(tv, dim=1), torch.linalg.norm(tv.cuda(), dim=1).cpu()
plain = (tv[:, 0] * tv[:, 0] + tv[:, 1] * tv[:, 1]).sqrt()
print("norm2: cuda==cpu %.6f plain==cpu %.6f plain==cuda %.6f" % ((nn_c == nn_g).float().mean(), (plain == nn_c).float().mean(), (plain == nn_g).float().mean()))
# cdist
xy = (torch.rand(64, 2, generator=g) * 200).float()
st = (torch.rand(300, 2, generator=g) * 200).float()
cc, cg = torch.cdist(xy, st), torch.cdist(xy.cuda(), st.cuda()).cpu()
x1 = torch.cat((xy * -2, xy.pow(2).sum(1, keepdim=True), torch.ones(64, 1)), 1)
x2 = torch.cat((st, torch.ones(300, 1), st.pow(2).sum(1, keepdim=True)), 1)
acc = x1[:, None, 0] * x2[None, :, 0]
for k in range(1, 4):
    acc = torch.addcmul(acc.double(), x1[:, None, k].double(), x2[None, :, k].double()).float()   # fma chain emulation
emu = acc.clamp_min(0).sqrt()
direct = ((xy[:, None, :] - st[None, :, :]).pow(2).sum(2)).sqrt()
print("cdist: cuda==cpu %.4f  fma-chain==cpu %.4f  fma-chain==cuda %.4f  max|cuda-cpu| %.3g max|cuda-direct| %.3g" % (
    (cc == cg).float().mean(), (emu == cc).float().mean(), (emu == cg).float().mean(), (cc - cg).abs().max(), (cg - direct).abs().max()))
print("round half even cuda", torch.round(torch.tensor([0.5, 1.5, 2.5], device=dev)).tolist())
