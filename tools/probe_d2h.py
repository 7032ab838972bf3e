"""Host-link probe: D2H / H2D bandwidth of pinned buffers, before and after binding the process to the GPU's NUMA node."""
import os
import subprocess
import time

import torch

print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
print("cpus allowed:", len(os.sched_getaffinity(0)), "of", os.cpu_count())
dev = torch.device("cuda:0")
x = torch.empty(917 * 1024 * 1024, dtype=torch.uint8, device=dev)


def bw(tag):
    h = torch.empty(x.numel(), dtype=torch.uint8).pin_memory()
    for direction in ("d2h", "h2d"):
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if direction == "d2h":
                h.copy_(x, non_blocking=True)
            else:
                x.copy_(h, non_blocking=True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        print("%s %s: %.1f GB/s (best of 5, %d MB)" % (tag, direction, x.numel() / min(ts) / 1e9, x.numel() >> 20))
    del h


bw("default affinity")
try:
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    nv.nvmlDeviceSetCpuAffinity(h)
    print("cpus allowed after nvmlDeviceSetCpuAffinity:", sorted(os.sched_getaffinity(0))[:8], "...", len(os.sched_getaffinity(0)))
    torch._C._host_emptyCache() if hasattr(torch._C, "_host_emptyCache") else None
    bw("GPU-local affinity")
except Exception as e:
    print("affinity probe failed:", e)
