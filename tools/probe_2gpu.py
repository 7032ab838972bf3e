"""torchrun probe: concurrent pinned D2H bandwidth per rank, and the host-pipeline loop with / without the stats all-reduce."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isaac_rover_b200 as R
from isaac_rover_b200 import synth
rank, world, local = R.dist.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.system("nvidia-smi topo -m 2>&1 | head -6") if rank == 0 else None
n = int(28.7e6)
d = torch.empty(n, dtype=torch.uint8, device=dev)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for it in range(2):
    R.dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("rank %d concurrent D2H: %.1f GB/s" % (rank, n * 20 / dt / 1e9), flush=True)
N = 4096
w = synth.make_world(length=200.0, nv=708, K=200, n_stones=2000, seed=42, build_index=None)
w.map_indices = R.build_knn_index(w.triangles, w.vertices, w.G, w.res, w.K, device=dev)
w.rock_indices = R.build_knn_index(w.rock_triangles, w.rock_vertices, w.G, w.res, w.K, device=dev)
states = [synth.make_env_state(w, N, seed=100 + s, env_offset=rank * N) for s in range(3)]
task = synth.make_task(w, states[0], device=str(dev), level=2, num_envs_total=N * world)
pipe = R.HostPipeline(task)
hs = [{k: v.pin_memory() for k, v in s.items() if k in ("pos", "quat", "joints", "actions")} for s in states]
def loop(steps, reduce, read):
    prev = None
    t0 = time.perf_counter()
    for i in range(steps):
        x = hs[i % 3]
        k = pipe.submit(x["pos"], x["quat"], x["joints"], x["actions"])
        if reduce:
            R.dist.reduce_stats(task.stats)
        if prev is not None and read:
            o, r, z = pipe.result(prev)
        prev = k
    pipe.result(prev)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3
for reduce in (True, False):
    for read in (True, False):
        loop(5, reduce, read)
        R.dist.barrier(); torch.cuda.synchronize()
        ms = loop(30, reduce, read)
        print("rank %d e2e loop reduce=%s read=%s: %.3f ms/step" % (rank, reduce, read, ms), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
