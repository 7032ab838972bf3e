"""Multi-GPU plumbing: one process per GPU, envs sharded by contiguous index blocks, terrain replicated,
one all-reduce(sum) of the 16-entry episode-statistics vector per step.  No data-path
collective: every env is independent in every function of the hot path."""
import os

import torch
import torch.distributed as dist


def env_shard(total_envs, rank, world_size):
    """Contiguous block [lo, hi) of env ids owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total_envs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*).
    Returns (rank, world_size, local_rank).  Single process when WORLD_SIZE is unset or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs NVML lists as local to its GPU (nvmlDeviceSetCpuAffinity), BEFORE any pinned host buffer
    is allocated: the pages of a pinned buffer live on the NUMA node of the thread that allocates them, and a GPU that writes
    its 0.9 GB observation block per step into the other socket's memory shares the inter-socket link with every other rank
    doing the same.  Returns the number of CPUs in the new affinity mask, or None if NVML is not available."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        try:
            uuid = torch.cuda.get_device_properties(local_rank).uuid
            h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
        except Exception:
            h = nv.nvmlDeviceGetHandleByIndex(local_rank)
        nv.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def reduce_stats(stats, async_op=False):
    """In-place all-reduce(sum) of a per-rank statistics vector (f64 [16]) on the current stream (NCCL) --
    the only exchange of the step.  No-op in a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(stats, op=dist.ReduceOp.SUM, async_op=async_op)
    return None


class StatsReducer:
    """The per-step statistics all-reduce taken off the critical path: the step's 16 sums are copied into one of `depth`
    buffers and all-reduced asynchronously (NCCL's own stream), so the next step's kernels do not wait for the slowest rank
    of this one.  submit() returns a slot; result(slot) makes the current stream wait for that reduction and returns the
    reduced vector.  Single process: the copy alone."""

    def __init__(self, depth=2):
        self.depth, self.bufs, self.works, self.i = depth, [None] * depth, [None] * depth, 0

    def submit(self, stats):
        k = self.i % self.depth
        self.i += 1
        if self.works[k] is not None:
            self.works[k].wait()
            self.works[k] = None
        if self.bufs[k] is None:
            self.bufs[k] = torch.empty_like(stats)
        self.bufs[k].copy_(stats)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.works[k] = dist.all_reduce(self.bufs[k], op=dist.ReduceOp.SUM, async_op=True)
        return k

    def result(self, k):
        if self.works[k] is not None:
            self.works[k].wait()
            self.works[k] = None
        return self.bufs[k]


def max_over_ranks(value, device):
    """max of a python float over ranks (timing)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(value)


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
