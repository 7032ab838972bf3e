"""GPU micro-benchmark of the rock-collision kernel alone (Rock_Detection.get_collisions + check_collision), 4096 / 65536 envs."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isaac_rover_b200 as R
w = R.synth.make_world(length=200.0, nv=708, K=200, n_stones=2000, seed=42, build_index=None)
w.rock_indices = R.build_knn_index(w.rock_triangles, w.rock_vertices, w.G, w.res, w.K, device="cuda:0")
rock = R.Rock_Detection("cuda:0", torch.tensor([0, 0, 0.0]), assets=(w.rock_indices, w.rock_triangles, w.rock_vertices))
for N in (4096, 65536):
    st = {k: v.cuda() for k, v in R.synth.make_env_state(w, N, seed=100).items()}
    eul = R.tensor_quat_to_eul(st["quat"])
    f = lambda: rock.get_collisions(st["pos"], eul, st["joints"])
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in ev:
        a.record(); f(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    wd, bd = f()
    print("N=%d: rock collision median %.4f ms, min %.4f ms; checksum %d" % (N, t[10], t[0], int(wd.view(torch.int16).long().sum() + bd.view(torch.int16).long().sum())))
