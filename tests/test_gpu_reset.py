"""GPU parity of rvb_reset_targets (device-side reset path) against oracle/reset_oracle.py, through the C ABI
and through RoverTask.reset_targets_device."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reset_oracle as RO          # noqa: E402
from test_reset_cpu import _world  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import isaac_rover_b200
    return isaac_rover_b200


def _run(R, reset, off, seed, epoch, initial, stones, hm, target, progress, max_attempts=64, sem=0):
    lib = R._lib.load()
    d = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)      # noqa: E731
    reset_d, init_d, st_d, hm_d, tg_d, pr_d = d(reset), d(initial), d(stones), d(hm), d(target), d(progress, torch.int64)
    cnt = torch.zeros(3, dtype=torch.int32, device="cuda")
    R._lib.check(lib.rvb_reset_targets(R._lib.ptr(reset_d), reset.shape[0], off, seed, epoch, R._lib.ptr(init_d), 8.0, R._lib.ptr(st_d),
                                       stones.shape[0], 1.0, max_attempts, R._lib.ptr(hm_d), hm.shape[0], hm.shape[1], 0.25, 1.0, 0.0, 0.0,
                                       R._lib.ptr(tg_d), R._lib.ptr(pr_d), R._lib.ptr(reset_d), R._lib.ptr(cnt), sem,
                                       R._lib.stream_of(tg_d)))
    torch.cuda.synchronize()
    return tg_d.cpu().numpy(), pr_d.cpu().numpy(), reset_d.cpu().numpy(), cnt.cpu().numpy()


def test_reset_targets_matches_oracle(R):
    initial, stones, hm, reset, target, progress = _world(seed=5, N=3000, S=500)
    t, p, r, c = _run(R, reset, 0, 42, 9, initial, stones, hm, target, progress)
    def cuda_cos_sin(alpha):          # the CUDA math library's cosf / sinf (what the kernel calls), through torch
        a = torch.from_numpy(np.ascontiguousarray(alpha)).cuda()
        return torch.cos(a).cpu().numpy(), torch.sin(a).cpu().numpy()
    to, po, ro, co, attempts = RO.reset_targets(reset, 0, 42, 9, initial, 8.0, stones, 1.0, 64, hm, 0.25, 1.0, (0.0, 0.0), target, progress,
                                                cos_sin=cuda_cos_sin)
    ids = np.nonzero(reset)[0]
    assert np.array_equal(p, po) and np.array_equal(r, ro) and (r == 0).all()
    assert c[0] == co[0] == ids.size and c[2] == 0
    # with the device's own cosf / sinf in the oracle every goal, every height and the number of draws are bit-identical
    assert np.array_equal(t, to)
    assert int(c[1]) == int(co[1])
    keep = reset == 0
    assert np.array_equal(t[keep], target[keep])
    # every goal the device produced clears the stones (library's own validator, direct formulation) and sits on the circle
    lib = R._lib.load()
    xy = torch.from_numpy(t[ids, :2].copy()).cuda()
    flag = torch.zeros(ids.size, dtype=torch.int64, device="cuda")
    st_d = torch.from_numpy(stones).cuda()
    R._lib.check(lib.rvb_stone_validate(R._lib.ptr(xy), 2, ids.size, R._lib.ptr(st_d), stones.shape[0], 1.0, 1, None, R._lib.ptr(flag),
                                        None, R._lib.stream_of(xy)))
    assert int(flag.sum()) == 0
    d = np.hypot(t[ids, 0].astype(np.float64) - initial[ids, 0], t[ids, 1].astype(np.float64) - initial[ids, 1])
    assert np.allclose(d, 8.0, atol=1e-4)


def test_reset_targets_sharding_and_edges(R):
    initial, stones, hm, reset, target, progress = _world(seed=6, N=1001, S=40)
    full = _run(R, reset, 0, 7, 3, initial, stones, hm, target, progress)[0]
    h = 400
    a = _run(R, reset[:h], 0, 7, 3, initial[:h], stones, hm, target[:h], progress[:h])[0]
    b = _run(R, reset[h:], h, 7, 3, initial[h:], stones, hm, target[h:], progress[h:])[0]
    assert np.array_equal(np.concatenate((a, b)), full)                       # env shards: bit-identical to the unsharded run
    assert np.array_equal(_run(R, reset, 0, 7, 3, initial, stones, hm, target, progress)[0], full)       # deterministic
    # nothing to reset: buffers untouched, counters zero
    t, p, r, c = _run(R, np.zeros_like(reset), 0, 7, 3, initial, stones, hm, target, progress)
    assert np.array_equal(t, target) and np.array_equal(p, progress) and (c == 0).all()
    # a goal that can never clear (stone of radius 100 on top of everything): the loop stops after max_attempts and says so
    big = stones.copy()
    big[0] = (30, 30, 0, 400, 400, 1, 100)
    t, p, r, c = _run(R, reset, 0, 7, 3, initial, big, hm, target, progress, max_attempts=5)
    n = int((reset != 0).sum())
    assert c[0] == n and c[1] == 5 * n and c[2] == n and (r == 0).all()
    # N = 0 and bad arguments
    lib = R._lib.load()
    assert lib.rvb_reset_targets(None, 0, 0, 1, 1, None, 8.0, None, 1, 1.0, 1, None, 1, 1, 0.1, 1.0, 0.0, 0.0, None, None, None, None, 0, None) == 0
    assert lib.rvb_reset_targets(None, 5, 0, 1, 1, None, 8.0, None, 1, 1.0, 1, None, 1, 1, 0.1, 1.0, 0.0, 0.0, None, None, None, None, 0, None) == -1


def test_task_pre_physics_step_device(R):
    w = R.synth.make_world(length=20.0, nv=72, K=64, n_stones=40, seed=42, build_index=None)
    w.map_indices = R.build_knn_index(w.triangles, w.vertices, w.G, w.res, w.K, device="cuda:0").cpu()
    kr = min(w.K, w.rock_triangles.shape[0])
    w.rock_indices = R.build_knn_index(w.rock_triangles, w.rock_vertices, w.G, w.res, kr, device="cuda:0").cpu()
    st = R.synth.make_env_state(w, 256, seed=9)
    task = R.synth.make_task(w, st, device="cuda:0", level=2)
    task.reset_buf[:] = 0
    task.reset_buf[::3] = 1
    before = task.target_positions.clone()
    task.progress_buf[:] = 17
    task.pre_physics_step_device(st["actions"].cuda())
    torch.cuda.synchronize()
    ids = torch.arange(256, device="cuda")[::3]
    keep = torch.ones(256, dtype=torch.bool, device="cuda")
    keep[::3] = False
    assert int(task.reset_buf.sum()) == 0 and (task.progress_buf[ids] == 0).all() and (task.progress_buf[keep] == 17).all()
    assert torch.equal(task.target_positions[keep], before[keep]) and not torch.equal(task.target_positions[ids], before[ids])
    c = task.reset_counters.cpu()
    assert c[0] == ids.numel() and c[2] == 0
    near, flag, _ = task.nearest_stone_edge(task.target_positions[ids][:, :2].contiguous(), 1.0)
    assert int(flag.sum()) <= 1          # cdist's matmul formulation (the reference's check at this size) may flip a borderline goal
    h = task.get_pos_height(task.heightmap, task.target_positions[ids][:, :2].contiguous(), task.horizontal_scale, task.vertical_scale,
                            task.shift[0:2])
    assert torch.equal(h, task.target_positions[ids][:, 2])
    assert task.joint_position_targets is not None          # the action half ran
