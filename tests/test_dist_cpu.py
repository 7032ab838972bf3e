"""Host-side multi-process logic on CPU (gloo, world_size 2): env sharding, the per-step statistics all-reduce and the
max-over-ranks timing helper used by bench.py.  The data path has no collective."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_envs, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import isaac_rover_b200 as R
    r, w, _ = R.dist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = R.dist.env_shard(total_envs, r, w)
    # per-rank statistics of "its" envs: envs, reward sum, resets (stand-ins with a closed form)
    ids = torch.arange(lo, hi, dtype=torch.float64)
    stats = torch.zeros(16, dtype=torch.float64)
    stats[0], stats[1], stats[8] = hi - lo, ids.sum(), (ids % 3 == 0).sum()
    # the asynchronous reducer: three steps through two slots, the source vector changes between submissions
    red = R.dist.StatsReducer(depth=2)
    slots = []
    for step in range(3):
        slots.append(red.submit(stats * (step + 1)))
    async_out = [red.result(slots[-1]).tolist(), red.result(slots[-2]).tolist()]
    R.dist.reduce_stats(stats)
    t = R.dist.max_over_ranks(1.0 + rank, torch.device("cpu"))
    R.dist.barrier()
    q.put((rank, lo, hi, stats.tolist(), t, async_out))
    dist.destroy_process_group()


def test_env_shard_partitions_exactly():
    sys.path.insert(0, ROOT)
    import isaac_rover_b200 as R
    for total, world in ((1048576, 8), (4096, 1), (65536, 3), (7, 4), (0, 2)):
        blocks = [R.dist.env_shard(total, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b[1] - b[0] for b in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_stats_reduction_gloo():
    total, world, port = 1001, 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = torch.arange(total, dtype=torch.float64)
    for rank, lo, hi, stats, t, async_out in out:
        assert stats[0] == total and stats[1] == ids.sum().item() and stats[8] == (ids % 3 == 0).sum().item()
        assert async_out[0][0] == 3 * total and async_out[0][1] == 3 * ids.sum().item()        # step 3 (slot reused)
        assert async_out[1][0] == 2 * total and async_out[1][1] == 2 * ids.sum().item()        # step 2
        assert t == 2.0                                        # max over ranks of (1 + rank)
    assert out[0][1] == 0 and out[0][2] == out[1][1] and out[1][2] == total


def test_single_process_helpers_are_noops():
    sys.path.insert(0, ROOT)
    import isaac_rover_b200 as R
    s = torch.arange(16, dtype=torch.float64)
    assert R.dist.reduce_stats(s) is None and torch.equal(s, torch.arange(16, dtype=torch.float64))
    assert R.dist.max_over_ranks(3.5, torch.device("cpu")) == 3.5
    R.dist.barrier()
    red = R.dist.StatsReducer()
    assert torch.equal(red.result(red.submit(s)), s)


def _hooks_worker(rank, world, port, total_envs, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import numpy as np
    import hooks_oracle as HO
    import reset_oracle as RO
    import isaac_rover_b200 as R
    r, w, _ = R.dist.init_from_env(backend="gloo")
    lo, hi = R.dist.env_shard(total_envs, r, w)
    base = np.full((hi - lo, 40), 0.25, np.float32)
    # what each rank's device kernels compute for its env block: env-indexed Philox draws with env_offset = first global env id
    hooked = HO.obs_hooks(base, 4, 0.45, 0.1, 0.02, None, seed=42, epoch=7, env_offset=lo)
    goal_u = RO.uniform(42, 7, np.arange(lo, hi), 0)
    parts = [None] * w
    dist.all_gather_object(parts, (lo, hooked, goal_u))
    dist.barrier()
    if r == 0:
        q.put(parts)
    dist.destroy_process_group()


def test_two_rank_env_indexed_draws_equal_unsharded_gloo():
    """'results are independent of W when env-indexed Philox counters are used' -- the observation hooks (8f-4) and the
    goal draws (8f-2) of two env shards, gathered, equal the unsharded run bit for bit."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import numpy as np
    import hooks_oracle as HO
    import reset_oracle as RO
    total, world, port = 77, 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_hooks_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = sorted(q.get(timeout=120), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = HO.obs_hooks(np.full((total, 40), 0.25, np.float32), 4, 0.45, 0.1, 0.02, None, seed=42, epoch=7)
    assert np.array_equal(np.concatenate([h for _, h, _ in parts]), full)
    assert np.array_equal(np.concatenate([u for _, _, u in parts]), RO.uniform(42, 7, np.arange(total), 0))
