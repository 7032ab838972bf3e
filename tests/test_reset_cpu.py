"""CPU tests of the device-side reset path's oracle (oracle/reset_oracle.py): Philox known answers (Random123), the goal
arithmetic against the golden vectors the reference's own random_goals / check_goal_collision / get_pos_height produced, and
the properties of the per-env rejection loop."""
import os

import numpy as np
import torch

import reset_oracle as RO

HERE = os.path.dirname(os.path.abspath(__file__))


def test_philox_known_answers():
    for c, k, want in RO.PHILOX_KAT:
        assert [int(x) for x in RO.philox4x32_10(c, k)] == list(want)
    # vectorised == scalar
    c = np.arange(8, dtype=np.uint64)
    out = RO.philox4x32_10((c, c + 1, 0, 7), (42, 0))
    for i in range(8):
        assert [int(x) for x in RO.philox4x32_10((i, i + 1, 0, 7), (42, 0))] == [int(o[i]) for o in out]


def test_goal_arithmetic_matches_reference_golden():
    g = torch.load(os.path.join(HERE, "golden", "reset_golden.pt"))
    M = g["initial"].shape[0]
    u = RO.uniform(g["seed"], g["epoch"], np.arange(M) + g["env_offset"], 0)
    assert np.array_equal(u, g["u"].numpy()) and (u >= 0).all() and (u < 1).all()
    x, y = RO.goal_from_uniform(u, g["initial"].numpy(), 8.0)
    ref = g["ref_target_xy"].numpy()
    assert np.allclose(x, ref[:, 0], rtol=1e-6, atol=1e-6) and np.allclose(y, ref[:, 1], rtol=1e-6, atol=1e-6)
    near = RO.nearest_edge(ref[:, 0].copy(), ref[:, 1].copy(), g["stones"].numpy())
    assert np.array_equal(near <= np.float32(1.0), g["ref_invalid"].numpy()) and int((near <= 1.0).sum()) == g["ref_invalid_count"]
    h = RO.height(g["heightmap"].numpy(), ref[:, 0].copy(), ref[:, 1].copy(), g["hscale"], g["vscale"], g["shift"].numpy(),
                  cuda_semantics=False)
    assert np.array_equal(h, g["ref_height"].numpy())


def _world(seed=3, N=200, S=300):
    rng = np.random.default_rng(seed)
    initial = np.concatenate((rng.uniform(10, 50, (N, 2)), np.zeros((N, 1))), 1).astype(np.float32)
    stones = np.zeros((S, 7), np.float32)
    stones[:, :2] = rng.uniform(0, 60, (S, 2))
    stones[:, 3:5] = rng.uniform(0.2, 3.0, (S, 2))
    stones[:, 6] = np.maximum(stones[:, 3], stones[:, 4]) / 4
    hm = rng.uniform(-1, 1, (240, 240)).astype(np.float32)
    reset = (rng.uniform(size=N) < 0.4).astype(np.int64)
    target = rng.uniform(0, 1, (N, 3)).astype(np.float32)
    progress = rng.integers(0, 3000, N)
    return initial, stones, hm, reset, target, progress


def test_rejection_loop_properties():
    initial, stones, hm, reset, target, progress = _world()
    t, p, r, counters, attempts = RO.reset_targets(reset, 0, 42, 9, initial, 8.0, stones, 1.0, 64, hm, 0.25, 1.0, (0.0, 0.0), target,
                                                   progress)
    ids = np.nonzero(reset)[0]
    keep = np.nonzero(reset == 0)[0]
    assert counters[0] == ids.size and counters[2] == 0 and counters[1] == attempts.sum() >= ids.size
    assert np.array_equal(t[keep], target[keep]) and np.array_equal(p[keep], progress[keep])           # untouched envs
    assert (p[ids] == 0).all() and (r == 0).all()
    assert (RO.nearest_edge(t[ids, 0], t[ids, 1], stones) > 1.0).all()                                  # every goal clears the stones
    d = np.hypot(t[ids, 0].astype(np.float64) - initial[ids, 0], t[ids, 1].astype(np.float64) - initial[ids, 1])
    assert np.allclose(d, 8.0, atol=1e-4)                                                               # on the r = 8 circle
    assert attempts.max() > 1                                                                           # the loop did retry
    # deterministic, and independent of how the envs are sharded (env_offset)
    t2 = RO.reset_targets(reset, 0, 42, 9, initial, 8.0, stones, 1.0, 64, hm, 0.25, 1.0, (0.0, 0.0), target, progress)[0]
    assert np.array_equal(t, t2)
    h = initial.shape[0] // 2
    ta = RO.reset_targets(reset[:h], 0, 42, 9, initial[:h], 8.0, stones, 1.0, 64, hm, 0.25, 1.0, (0.0, 0.0), target[:h], progress[:h])[0]
    tb = RO.reset_targets(reset[h:], h, 42, 9, initial[h:], 8.0, stones, 1.0, 64, hm, 0.25, 1.0, (0.0, 0.0), target[h:], progress[h:])[0]
    assert np.array_equal(np.concatenate((ta, tb)), t)
    # another epoch draws other goals
    t3 = RO.reset_targets(reset, 0, 42, 10, initial, 8.0, stones, 1.0, 64, hm, 0.25, 1.0, (0.0, 0.0), target, progress)[0]
    assert not np.array_equal(t3[ids, :2], t[ids, :2])
    # uniform angles: mean direction of many draws ~ 0
    u = RO.uniform(1, 2, np.arange(20000), 0)
    assert abs(u.mean() - 0.5) < 0.01 and abs(np.cos(2 * np.pi * u).mean()) < 0.02
