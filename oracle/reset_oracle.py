"""TEST INFRASTRUCTURE ONLY (imported by tests/ alone; the product path never touches oracle/).

CPU restatement (numpy) of the device-side reset path rvb_reset_targets (include/rover_b200.h): the reference's
reset_idx book-keeping (rover.py:451-452), generate_goals / random_goals / check_goal_collision (rover.py:533-564) run per
env, and set_targets' height look-up (rover.py:582-583, get_pos_height :588-608), with the reference's torch.rand replaced
by a counter-based generator -- Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11;
Random123 v1.14 known-answer vectors below) -- because a draw order "one number per resetting env, in env order, per retry"
cannot be reproduced without a host round trip.  Parity for this path: PINNED for the generator (Random123 KATs) and for
the goal arithmetic against the reference's own random_goals / check_goal_collision fed with the same uniform numbers
(tests/test_reset_cpu.py imports the reference); the draw ORDER is this repository's definition, not the reference's.
"""
import math

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF

# Random123 kat_vectors: philox4x32 10 rounds  (counter[4], key[2]) -> output[4]
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(c, k):
    """c: 4 uint32 arrays (or ints), k: 2 -> 4 uint32 arrays.  Vectorised over numpy arrays."""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) & MASK for x in c]
    k0, k1 = [np.asarray(x, dtype=np.uint64) & MASK for x in k]
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return [x.astype(np.uint32) for x in (c0, c1, c2, c3)]


def uniform(seed, epoch, gid, attempt):
    """The k-th draw of env gid at call `epoch`: u in [0, 1), 24 bits (what torch.rand yields for f32)."""
    gid = np.asarray(gid, dtype=np.uint64)
    r0 = philox4x32_10((gid & MASK, gid >> np.uint64(32), np.uint64(attempt), np.uint64(epoch & MASK)),
                       (np.uint64(seed & MASK), np.uint64(((seed >> 32) ^ (epoch >> 32)) & MASK)))[0]
    return ((r0 >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def goal_from_uniform(u, initial_xy, radius, cos_sin=None):
    """random_goals (rover.py:554-564) in f32: alpha = 2 pi u; target = radius * (cos, sin) + 0 + initial.
    cos_sin: optional f32 array -> (cos, sin) f32 arrays; default numpy's (tests pass the CUDA math library's, whose last ulp
    differs from numpy's, to demand bit-identical goals from the kernel)."""
    f = np.float32
    alpha = (f(2 * math.pi) * u).astype(f)
    c, s_ = cos_sin(alpha) if cos_sin is not None else (np.cos(alpha), np.sin(alpha))
    x = ((f(radius) * np.asarray(c).astype(f)).astype(f) + f(0)) + initial_xy[:, 0].astype(f)
    y = ((f(radius) * np.asarray(s_).astype(f)).astype(f) + f(0)) + initial_xy[:, 1].astype(f)
    return x.astype(f), y.astype(f)


def nearest_edge(x, y, stones):
    """check_goal_collision's distance (rover.py:536-538), direct formula, f32."""
    f = np.float32
    dx = (x[:, None] - stones[None, :, 0].astype(f)).astype(f)
    dy = (y[:, None] - stones[None, :, 1].astype(f)).astype(f)
    d = np.sqrt((dx * dx + dy * dy).astype(f)).astype(f) - stones[None, :, 6].astype(f)
    return d.min(axis=1)


def height(hm, x, y, hscale, vscale, shift, cuda_semantics=True):
    f = np.float32
    if cuda_semantics:
        inv = f(1.0) / f(hscale)
        u, v = (x - f(shift[0])) * inv, (y - f(shift[1])) * inv
    else:
        u, v = (x - f(shift[0])) / f(hscale), (y - f(shift[1])) / f(hscale)
    hi = f(hm.shape[0] - 1)
    i = np.rint(np.clip(u.astype(f), f(0), hi)).astype(np.int64)
    j = np.minimum(np.rint(np.clip(v.astype(f), f(0), hi)).astype(np.int64), hm.shape[1] - 1)
    return (hm[i, j].astype(f) * f(vscale)).astype(f)


def reset_targets(reset, env_offset, seed, epoch, initial_pos, radius, stones, thr, max_attempts, hm, hscale, vscale, shift,
                  target, progress, cuda_semantics=True, cos_sin=None):
    """-> (target, progress, reset_out, counters[3], attempts per env)"""
    reset = np.asarray(reset)
    target, progress = target.copy(), progress.copy()
    ids = np.nonzero(reset != 0)[0]
    attempts = np.zeros(reset.shape[0], dtype=np.int64)
    failed = 0
    todo = ids.copy()
    k = 0
    x = np.zeros(0, np.float32)
    while todo.size and k < max_attempts:
        u = uniform(seed, epoch, todo + env_offset, k)
        x, y = goal_from_uniform(u, initial_pos[todo], radius, cos_sin)
        target[todo, 0], target[todo, 1] = x, y
        attempts[todo] = k + 1
        bad = nearest_edge(x, y, stones) <= np.float32(thr)
        todo = todo[bad]
        k += 1
    failed = int(todo.size)
    if ids.size:
        target[ids, 2] = height(hm, target[ids, 0], target[ids, 1], hscale, vscale, shift, cuda_semantics)
    progress[ids] = 0
    reset_out = reset.copy()
    reset_out[ids] = 0
    return target, progress, reset_out, np.array([ids.size, int(attempts.sum()), failed], dtype=np.int32), attempts
