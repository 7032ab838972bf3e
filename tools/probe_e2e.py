"""Where does the host-pipeline step time go?  (1 GPU)"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isaac_rover_b200 as R
from isaac_rover_b200 import synth
dev = torch.device("cuda", 0)
N = 4096
w = synth.make_world(length=200.0, nv=708, K=200, n_stones=2000, seed=42, build_index=None)
w.map_indices = R.build_knn_index(w.triangles, w.vertices, w.G, w.res, w.K, device=dev)
w.rock_indices = R.build_knn_index(w.rock_triangles, w.rock_vertices, w.G, w.res, w.K, device=dev)
states = [synth.make_env_state(w, N, seed=100 + s) for s in range(3)]
task = synth.make_task(w, states[0], device=str(dev), level=2, num_envs_total=N)
dst = [{k: v.to(dev) for k, v in s.items()} for s in states]
# device-resident loop
def dev_loop(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    t0 = time.perf_counter()
    for i in range(steps):
        s = dst[i % 3]
        task._rover.pos, task._rover.quat, task._rover.joints = s["pos"], s["quat"], s["joints"]
        task.hot_step(s["actions"])
    t_host = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, t_host / steps * 1e3
dev_loop(5)
print("device loop: %.3f ms/step on the GPU, host enqueue %.3f ms/step" % dev_loop(100))
pipe = R.HostPipeline(task)
hs = [{k: v.pin_memory() for k, v in s.items() if k in ("pos", "quat", "joints", "actions")} for s in states]
def loop(steps, read=True, d2h=True):
    prev = None
    sub = 0.0
    t0 = time.perf_counter()
    for i in range(steps):
        x = hs[i % 3]
        a = time.perf_counter()
        k = pipe.submit(x["pos"], x["quat"], x["joints"], x["actions"])
        sub += time.perf_counter() - a
        if prev is not None and read:
            pipe.result(prev)
        prev = k
    pipe.result(prev)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3, sub / steps * 1e3
for read in (True, False):
    loop(5, read)
    ms, sub = loop(100, read)
    print("host pipeline read=%s: %.3f ms/step, of which inside submit() %.3f ms" % (read, ms, sub))
# synchronous step
for _ in range(3):
    pipe.step(hs[0]["pos"], hs[0]["quat"], hs[0]["joints"], hs[0]["actions"])
t0 = time.perf_counter()
for i in range(50):
    x = hs[i % 3]
    pipe.step(x["pos"], x["quat"], x["joints"], x["actions"])
print("synchronous step(): %.3f ms" % ((time.perf_counter() - t0) / 50 * 1e3))
# D2H alone while a kernel loop runs
def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
x = hs[0]
def h2d():
    for src, stage, dstt in ((x["pos"], pipe.h_pos[0], pipe.d_pos[0]), (x["quat"], pipe.h_quat[0], pipe.d_quat[0]), (x["joints"], pipe.h_joints[0], pipe.d_joints[0]), (x["actions"], pipe.h_actions[0], pipe.d_actions[0])):
        stage.copy_(src); dstt.copy_(stage, non_blocking=True)
    torch.cuda.synchronize()
print("h2d + sync: %.3f ms" % t(h2d))
def comp():
    task.hot_step(pipe.d_actions[0]); torch.cuda.synchronize()
print("hot_step + sync: %.3f ms" % t(comp))
def d2h():
    with torch.cuda.stream(pipe.copy_stream):
        pipe.h_obs[0].copy_(pipe.d_obs[0], non_blocking=True)
        pipe.h_rew[0].copy_(pipe.d_rew[0], non_blocking=True)
        pipe.h_reset[0].copy_(pipe.d_reset[0], non_blocking=True)
    torch.cuda.synchronize()
print("d2h + sync: %.3f ms" % t(d2h))
print("pinned?", pipe.h_obs[0].is_pinned(), pipe.h_obs[0].shape, pipe.d_obs[0].is_contiguous())
def ev():
    e = torch.cuda.Event(); e.record(); e.synchronize()
print("event record+sync: %.3f ms" % t(ev))
import ctypes
def step_sync():
    k = pipe.submit(x["pos"], x["quat"], x["joints"], x["actions"]); a = time.perf_counter(); pipe.result(k); return time.perf_counter() - a
step_sync(); print("result() wait inside sync step: %.3f ms" % (sum(step_sync() for _ in range(10)) / 10 * 1e3))
