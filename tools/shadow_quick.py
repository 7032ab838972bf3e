"""Quick A/B of the production ray-cast: bit-equality with the tiled kernel + timings at 4096 and 32768 envs + debug counters.
python tools/shadow_quick.py [ENVVAR=value ...]   (each ENVVAR=value pair is timed as an extra configuration)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isaac_rover_b200 as R      # noqa: E402


def timed(fn, reps=30, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2], t[0]


w = R.synth.make_world(length=200.0, nv=708, K=200, n_stones=2000, seed=42, build_index=None)
w.map_indices = R.build_knn_index(w.triangles, w.vertices, w.G, w.res, w.K, device="cuda:0")
cam = R.Camera("cuda:0", torch.tensor([0, 0, 0.0]), assets=(w.map_indices, w.triangles, w.vertices))
print("layer %.2f GB" % (cam.layer.bytes() / 1e9))
configs = [()] + [tuple(a.split("=", 1)) for a in sys.argv[1:]]
for N in (4096, 32768):
    st = {k: v.cuda() for k, v in R.synth.make_env_state(w, N, seed=100).items()}
    eul = R.tensor_quat_to_eul(st["quat"])
    cam.variant = 3
    obs_ref = torch.zeros((N, 1750), device="cuda")
    ref, _, _ = cam.get_depths(st["pos"], eul, want_pt=False, obs=obs_ref)
    ref = ref.clone()
    cam.variant = 0
    for rnd in range(2):
        for cfg in configs:
            if cfg:
                os.environ[cfg[0]] = cfg[1]
            obs = torch.zeros((N, 1750), device="cuda")
            med, mn = timed(lambda: cam.get_depths(st["pos"], eul, want_pt=False, obs=obs), reps=30 if N == 4096 else 10)
            obs.zero_()
            d, _, _ = cam.get_depths(st["pos"], eul, want_pt=False, obs=obs)
            ok_obs = bool(torch.equal(obs, obs_ref)) if obs_ref is not None else None
            print("N %6d %-28s median %.3f ms  min %.3f  (%.2f M envs/s)  equal to the tiled kernel: dist %s obs %s" % (
                N, "=".join(cfg) if cfg else "default", med, mn, N / med / 1e3, bool(torch.equal(d.view(torch.int16), ref.view(torch.int16))), ok_obs))
            if cfg:
                os.environ.pop(cfg[0])
    if N == 4096:
        sys.stdout.flush()
        os.environ["RVB_SHADOW_DBG"] = "1"
        cam.get_depths(st["pos"], eul, want_pt=False)
        torch.cuda.synchronize()
        os.environ.pop("RVB_SHADOW_DBG")
