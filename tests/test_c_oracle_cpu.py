"""The plain-C restatement of the reference's ray / triangle test (oracle/ray_oracle.c, gcc, no torch in the arithmetic) against
the golden vectors the reference itself produced and against the torch-based oracle: two independent statements of the fp16
arithmetic that every CUDA ray-cast kernel is held to bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

import rover_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


@pytest.fixture(scope="module")
def clib():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    lib = C.CDLL(os.path.join(ORACLE_DIR, "libray_oracle.so"))
    p, i64 = C.c_void_p, C.c_int64
    lib.rvo_ray_distance.argtypes = [p, p, p, i64, p, p]
    lib.rvo_ray_distance.restype = None
    lib.rvo_cast_min.argtypes = [p, p, i64, p, i64, p, p, p, p]
    lib.rvo_cast_min.restype = None
    lib.rvo_get_depths_env.argtypes = [p, p, p, i64, p, i64, i64, p, p, C.c_float, C.c_float, C.c_float, p, p, p]
    lib.rvo_get_depths_env.restype = None
    lib.rvo_cast_rays_env.argtypes = [p, p, i64, p, i64, i64, p, p, C.c_float, C.c_float, C.c_float, p, p]
    lib.rvo_cast_rays_env.restype = None
    lib.rvo_rock_rays_env.argtypes = [p, p, p, p, p]
    lib.rvo_rock_rays_env.restype = None
    lib.rvo_check_collision.argtypes = [p, i64, p, i64]
    lib.rvo_check_collision.restype = i64
    lib.rvo_ackermann.argtypes = [p, p, i64, p, p]
    lib.rvo_ackermann.restype = None
    lib.rvo_joint_targets.argtypes = [p, p, i64, p, p]
    lib.rvo_joint_targets.restype = None
    lib.rvo_reward_reset.argtypes = [p] * 8 + [i64, p, p, p, C.c_int, i64, i64, p, p, p]
    lib.rvo_reward_reset.restype = None
    lib.rvo_heightmap_pattern.argtypes = [p, i64, p, p, p, p]
    lib.rvo_heightmap_pattern.restype = i64
    lib.rvo_read_stone_info.argtypes = [p, i64, p]
    lib.rvo_read_stone_info.restype = None
    lib.rvo_nearest_stone_edge.argtypes = [p, i64, i64, p, i64, p]
    lib.rvo_nearest_stone_edge.restype = None
    lib.rvo_avoid_pos_rock_collision.argtypes = [p, i64, p, i64, i64]
    lib.rvo_avoid_pos_rock_collision.restype = i64
    lib.rvo_quat_to_euler.argtypes = [p, i64, p]
    lib.rvo_obs_proprio.argtypes = [p, p, p, p, p, i64, p, p]
    lib.rvo_obs_heightmap.argtypes = [p, p, i64, p]
    lib.rvo_pos_height.argtypes = [p, i64, i64, p, i64, i64, C.c_float, C.c_float, C.c_float, C.c_float, p]
    for fn in (lib.rvo_quat_to_euler, lib.rvo_obs_proprio, lib.rvo_obs_heightmap, lib.rvo_pos_height):
        fn.restype = None
    return lib


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "rover_golden.pt"))


def _u16(t):
    return np.ascontiguousarray(t.to(torch.float16).contiguous().view(torch.int16).numpy().view(np.uint16))


def c_ray_distance(lib, src, dirs, tri):
    s, d, t = _u16(src), _u16(dirs), _u16(tri)
    n = s.shape[0]
    k = np.empty(n, np.uint16)
    pt = np.empty((n, 3), np.uint16)
    lib.rvo_ray_distance(s.ctypes.data, d.ctypes.data, t.ctypes.data, n, k.ctypes.data, pt.ctypes.data)
    return k, pt


def same_halves(a_u16, b_half):
    """Bit equality, any NaN equal to any NaN (payloads are not part of the contract)."""
    b = _u16(b_half).reshape(a_u16.shape)
    nan_a, nan_b = (a_u16 & 0x7FFF) > 0x7C00, (b & 0x7FFF) > 0x7C00
    return bool(np.array_equal(nan_a, nan_b) and np.array_equal(a_u16[~nan_a], b[~nan_b]))


def test_golden_ray_distance_vectors(clib, golden):
    for tag in ("rd", "rr"):
        k, pt = c_ray_distance(clib, golden["in_%s_src" % tag], golden["in_%s_dir" % tag], golden["in_%s_tri" % tag])
        assert same_halves(k, golden["ref_%s_k" % tag]) and same_halves(pt, golden["ref_%s_pt" % tag]), tag
    k, _ = c_ray_distance(clib, golden["in_rd_src"], golden["in_rd_dir"], golden["in_rd_tri"])
    assert [float(x) for x in k[:5].view(np.float16)] == [1.0, -1.0, 11.0, 11.0, 11.0]


def test_random_pairs_equal_torch_oracle(clib):
    g = torch.Generator().manual_seed(7)
    n = 200_000
    tri = (torch.rand(n, 3, 3, generator=g) * 4 - 2).half()
    src = (torch.rand(n, 3, generator=g) * 4 - 2).half()
    dirs = (torch.rand(n, 3, generator=g) * 2 - 1).half()
    # degenerate cases: zero direction, zero-area triangle, ray in the triangle's plane, huge values (fp16 overflow -> inf / nan)
    dirs[:100] = 0
    tri[100:200, 1] = tri[100:200, 0]
    dirs[200:300, 2] = 0
    tri[200:300, :, 2] = 0
    src[300:400] *= 300
    tri[400:500] *= 300
    k_ref, pt_ref = O.ray_distance(src, dirs, tri)
    k, pt = c_ray_distance(clib, src, dirs, tri)
    assert same_halves(k, k_ref) and same_halves(pt, pt_ref)
    hits = (k_ref.float() < 11.0).float().mean().item()
    assert 0.01 < hits < 0.9                                      # the sample exercises both outcomes


def test_golden_get_depths_through_c(clib, golden):
    """camera.py:84-120 for the first envs of the golden world: K-list gather, ray_distance per pair, torch.min(dim) value + index."""
    w = golden["world"]
    n_env = 2
    pos, euler = golden["in_pos"][:n_env], golden["ref_euler"][:n_env]
    src, d = O.depth_transform(pos, euler, golden["ref_pattern"])
    idx = O.permute_index(w["map_indices"].to(torch.int32))
    G, K = idx.shape[0], idx.shape[2]
    cx, cy = O.cell_lookup(src[:, :, 0:2], torch.zeros(2), 0.1, G)
    tri = np.ascontiguousarray(w["triangles"].to(torch.int32).numpy())
    ver = _u16(w["vertices"])
    for e in range(n_env):
        ids = np.ascontiguousarray(idx[cx[e], cy[e]].numpy().astype(np.int32))          # [P,K]
        P = ids.shape[0]
        s = _u16(src[e])
        dd = _u16(d[e].unsqueeze(0).expand(P, 3))
        dist, slot = np.empty(P, np.uint16), np.empty(P, np.int32)
        clib.rvo_cast_min(s.ctypes.data, dd.ctypes.data, P, ids.ctypes.data, K, tri.ctypes.data, ver.ctypes.data,
                          dist.ctypes.data, slot.ctypes.data)
        assert same_halves(dist, golden["ref_dist"][e])
        assert np.array_equal(slot, golden["oracle_slot"][e].numpy().astype(np.int32))


def test_golden_get_depths_whole_chain_in_c(clib, golden):
    """Camera.get_depths end to end in C for every env of the golden world (pose transform in fp64, fp16 cast, cell lookup, K-list
    gather, ray_distance, min): sources, distances and hit slots equal the reference's outputs bit for bit.  The sin/cos values
    are the reference run's own (golden["trig"]), so no libm enters the comparison."""
    w = golden["world"]
    pat = np.ascontiguousarray(golden["ref_pattern"].numpy().astype(np.float64))
    P = pat.shape[0]
    m = np.ascontiguousarray(w["map_indices"].to(torch.int32).numpy())
    K, G = m.shape[0], m.shape[1]
    tri = np.ascontiguousarray(w["triangles"].to(torch.int32).numpy())
    ver = _u16(w["vertices"])
    for e in range(golden["in_pos"].shape[0]):
        pos = np.ascontiguousarray(golden["in_pos"][e].numpy().astype(np.float32))
        trig = np.ascontiguousarray(golden["trig"][e].numpy().astype(np.float32))
        src, dist, slot = np.empty((P, 3), np.uint16), np.empty(P, np.uint16), np.empty(P, np.int32)
        clib.rvo_get_depths_env(pos.ctypes.data, trig.ctypes.data, pat.ctypes.data, P, m.ctypes.data, G, K, tri.ctypes.data,
                                ver.ctypes.data, 0.0, 0.0, 0.1, src.ctypes.data, dist.ctypes.data, slot.ctypes.data)
        assert same_halves(src, golden["ref_sources"][e]), e
        assert same_halves(dist, golden["ref_dist"][e]), e
        assert np.array_equal(slot, golden["oracle_slot"][e].numpy().astype(np.int32)), e


def _c_ackermann(lib, lin, ang):
    l, a = np.ascontiguousarray(lin.numpy().astype(np.float32)), np.ascontiguousarray(ang.numpy().astype(np.float32))
    n = l.shape[0]
    steer, vel = np.empty((n, 6), np.float32), np.empty((n, 6), np.float32)
    lib.rvo_ackermann(l.ctypes.data, a.ctypes.data, n, steer.ctypes.data, vel.ctypes.data)
    return steer, vel


def test_golden_ackermann_in_c(clib, golden):
    """kinematics.py:14-67 in C: wheel velocities bit for bit (IEEE arithmetic only), steering angles to an ulp of atan2."""
    for lin, ang, ref_s, ref_v in ((golden["in_ka_lin"], golden["in_ka_ang"], golden["ref_ka_steer"], golden["ref_ka_vel"]),
                                   (golden["in_actions"][:, 0], golden["in_actions"][:, 1], golden["ref_steer"], golden["ref_vel"])):
        steer, vel = _c_ackermann(clib, lin, ang)
        rv = ref_v.numpy()
        assert np.array_equal(np.isnan(vel), np.isnan(rv)) and np.array_equal(vel[~np.isnan(rv)], rv[~np.isnan(rv)])
        assert np.allclose(steer, ref_s.numpy(), rtol=0, atol=2.4e-7, equal_nan=True)
    # every branch is in the sample: turn on the spot, straight (ang = 0 -> P = inf), reverse, |P| just above / below the bound
    lin, ang = golden["in_ka_lin"], golden["in_ka_ang"]
    P = torch.copysign(lin / ang, -ang)
    assert (P.abs() <= 0.45).any() and torch.isinf(P).any() and (lin < 0).any() and ((lin == 0) & (ang != 0)).any()
    steer, vel = _c_ackermann(clib, lin, ang)
    pos, vt = np.empty((len(lin), 4), np.float32), np.empty((len(lin), 6), np.float32)
    clib.rvo_joint_targets(steer.ctypes.data, vel.ctypes.data, len(lin), pos.ctypes.data, vt.ctypes.data)
    p_ref, v_ref = O.joint_targets(torch.from_numpy(steer), torch.from_numpy(vel))
    assert np.array_equal(pos, p_ref.numpy()) and np.array_equal(vt, v_ref.numpy(), equal_nan=True)


def test_golden_reward_and_reset_in_c(clib, golden):
    """calculate_metrics + is_done (rover.py:460-531,610-647) in C, both curriculum levels: rewards, reward terms and reset masks
    bit for bit against the reference's outputs."""
    g = golden
    f = lambda t: np.ascontiguousarray(t.numpy().astype(np.float32))
    pos, target, heading = f(g["in_pos"]), f(g["in_target"]), f(g["ref_heading"])
    lin, ang = f(g["in_actions"][:, 0]), f(g["in_actions"][:, 1])
    lin_p, ang_p = f(g["in_prev_actions"][:, 0]), f(g["in_prev_actions"][:, 1])
    joints, eul = f(g["in_joints"]), f(g["ref_euler"])
    progress = np.ascontiguousarray(g["in_progress"].numpy().astype(np.int64))
    coll = np.ascontiguousarray(g["ref_rock_collision"].numpy().astype(np.int64))
    n = pos.shape[0]
    for level, ref_rew, ref_reset in ((2, g["ref_rew"], g["ref_reset"]), (1, g["ref_rew_level1"], g["ref_reset_level1"])):
        rew, reset, ex = np.empty(n, np.float32), np.empty(n, np.int64), np.empty((n, 5), np.float32)
        clib.rvo_reward_reset(pos.ctypes.data, target.ctypes.data, heading.ctypes.data, lin.ctypes.data, lin_p.ctypes.data,
                              ang.ctypes.data, ang_p.ctypes.data, joints.ctypes.data, joints.shape[1], progress.ctypes.data,
                              coll.ctypes.data if level >= 2 else None, eul.ctypes.data, level, 3000, n, rew.ctypes.data,
                              reset.ctypes.data, ex.ctypes.data)
        assert np.array_equal(rew, ref_rew.numpy()), level
        assert np.array_equal(reset, ref_reset.numpy()), level
        if level == 2:
            e = g["ref_extras"]
            for col, key in enumerate(("pos_reward", "heading_contraint_penalty", "motion_contraint_penalty", "goal_angle_penalty",
                                       "uprightness_penalty")):
                assert np.array_equal(ex[:, col], e[key].numpy().astype(np.float32)), key
    assert g["ref_reset"].sum() > 0 and (g["ref_reset"] == 0).sum() > 0


def test_golden_observation_in_c(clib, golden):
    """get_observations (rover.py:272-336) in C: quat -> euler and the proprioceptive columns to an ulp of libm's atan2 / asin /
    sin / cos, the 1746 heightmap columns (fp16 halving of the ray distances, gathered by the pattern's index vectors) bit for bit."""
    g = golden
    f = lambda t: np.ascontiguousarray(t.numpy().astype(np.float32))
    n = g["in_quat"].shape[0]
    q, e = f(g["in_quat"]), np.empty((n, 3), np.float32)
    clib.rvo_quat_to_euler(q.ctypes.data, n, e.ctypes.data)
    assert np.allclose(e, g["ref_euler"].numpy(), rtol=0, atol=5e-7)
    pos, target, eul = f(g["in_pos"]), f(g["in_target"]), f(g["ref_euler"])
    lin, ang = f(g["in_actions"][:, 0]), f(g["in_actions"][:, 1])
    obs4, heading = np.empty((n, 4), np.float32), np.empty(n, np.float32)
    clib.rvo_obs_proprio(pos.ctypes.data, eul.ctypes.data, target.ctypes.data, lin.ctypes.data, ang.ctypes.data, n,
                         obs4.ctypes.data, heading.ctypes.data)
    assert np.allclose(heading, g["ref_heading"].numpy(), rtol=0, atol=5e-7)
    assert np.allclose(obs4, g["ref_obs"][:, :4].numpy(), rtol=1e-6, atol=2e-7)
    assert np.array_equal(obs4[:, 2:], g["ref_obs"][:, 2:4].numpy())
    idx = np.ascontiguousarray(torch.cat((g["ref_coarse_idx"], g["ref_fine_idx"])).numpy().astype(np.int64))
    assert idx.shape[0] == 634 + 1112                                   # teacher_loader.py:47-48
    for env in range(n):
        d = _u16(g["ref_dist"][env])
        out = np.empty(idx.shape[0], np.float32)
        clib.rvo_obs_heightmap(d.ctypes.data, idx.ctypes.data, idx.shape[0], out.ctypes.data)
        assert np.array_equal(out, g["ref_obs"][env, 4:].numpy()), env


def test_golden_pos_height_in_c(clib, golden):
    g = golden
    hm = np.ascontiguousarray(g["world"]["heightmap"].numpy().astype(np.float32))
    xy = np.ascontiguousarray(g["ref_spawn_pos"][:, 0:2].numpy().astype(np.float32))
    out = np.empty(xy.shape[0], np.float32)
    clib.rvo_pos_height(hm.ctypes.data, hm.shape[0], hm.shape[1], xy.ctypes.data, 2, xy.shape[0], float(g["world"]["hm_res"]), 1.0,
                        0.0, 0.0, out.ctypes.data)
    assert np.array_equal(out, g["ref_spawn_height"].numpy())


def test_golden_rock_cast_and_collision_in_c(clib, golden):
    """Rock_Detection.get_collisions' cast (rock_detect.py:65-115) + check_collision (rover.py:663-668) in C, fed with the 26 rays per
    env the reference produced: wheel / body distances and collision flags bit for bit."""
    g, w = golden, golden["world"]
    m = np.ascontiguousarray(w["rock_indices"].to(torch.int32).numpy())
    K, G = m.shape[0], m.shape[1]
    tri = np.ascontiguousarray(w["rock_triangles"].to(torch.int32).numpy())
    ver = _u16(w["rock_vertices"])
    ref = torch.cat((g["ref_wheel"], g["ref_body"]), 1)
    flags = []
    for e in range(ref.shape[0]):
        s, d = _u16(g["ref_rock_sources"][e]), _u16(g["ref_rock_dirs"][e])
        R = s.shape[0]
        dist, slot = np.empty(R, np.uint16), np.empty(R, np.int32)
        clib.rvo_cast_rays_env(s.ctypes.data, d.ctypes.data, R, m.ctypes.data, G, K, tri.ctypes.data, ver.ctypes.data, 0.0, 0.0, 0.1,
                               dist.ctypes.data, slot.ctypes.data)
        assert R == 26 and same_halves(dist, ref[e]), e
        wd, bd = np.ascontiguousarray(dist[:24]), np.ascontiguousarray(dist[24:])
        flags.append(clib.rvo_check_collision(wd.ctypes.data, 24, bd.ctypes.data, 2))
    assert flags == g["ref_rock_collision"].tolist()
    assert 0 < sum(flags) < len(flags)


def test_golden_rock_rays_in_c(clib, golden):
    """_get_wheel_rays / _get_body_rays (rock_detect.py:160-371) in C.  The joint-angle sin / cos are libm's (Sleef's in torch), so
    the fp32 values may differ in the last ulp before the fp16 cast: every ray within one fp16 ulp, almost all bit-identical."""
    g = golden
    n = g["in_pos"].shape[0]
    same = total = 0
    for e in range(n):
        pos = np.ascontiguousarray(g["in_pos"][e].numpy().astype(np.float32))
        trig = np.ascontiguousarray(g["trig"][e].numpy().astype(np.float32))
        joints = np.ascontiguousarray(g["in_joints"][e].numpy().astype(np.float32))
        src, dirs = np.empty((26, 3), np.uint16), np.empty((26, 3), np.uint16)
        clib.rvo_rock_rays_env(pos.ctypes.data, trig.ctypes.data, joints.ctypes.data, src.ctypes.data, dirs.ctypes.data)
        rs, rd = g["ref_rock_sources"][e], g["ref_rock_dirs"][e]
        assert np.abs(src.view(np.float16).astype(np.float32) - rs.float().numpy()).max() <= 2e-3
        assert np.abs(dirs.view(np.float16).astype(np.float32) - rd.float().numpy()).max() <= 1e-3
        same += int((src == _u16(rs)).sum() + (dirs == _u16(rd)).sum())
        total += src.size + dirs.size
    assert same / total > 0.98, same / total


def test_golden_stones_in_c(clib, golden):
    """read_stone_info, check_goal_collision's distance, avoid_pos_rock_collision (terrain_utils.py:416-424, rover.py:533-542,649-661)."""
    g = golden
    s6 = np.ascontiguousarray(g["world"]["stone_info6"].numpy().astype(np.float64))
    S = s6.shape[0]
    s7 = np.empty((S, 7), np.float32)
    clib.rvo_read_stone_info(s6.ctypes.data, S, s7.ctypes.data)
    assert np.array_equal(s7, g["ref_stone7"].numpy())
    xy = np.ascontiguousarray(g["in_target"][:, 0:2].numpy().astype(np.float32))
    near = np.empty(xy.shape[0], np.float32)
    clib.rvo_nearest_stone_edge(xy.ctypes.data, 2, xy.shape[0], s7.ctypes.data, S, near.ctypes.data)
    assert xy.shape[0] <= 25 and np.array_equal(near, g["ref_goal_nearest"].numpy())          # cdist's direct path: bit for bit
    assert int((near <= 1.0).sum()) == g["ref_goal_count"]
    many = np.ascontiguousarray(g["in_spawn_pos"].numpy().astype(np.float32))
    near = np.empty(many.shape[0], np.float32)
    clib.rvo_nearest_stone_edge(many.ctypes.data, 3, many.shape[0], s7.ctypes.data, S, near.ctypes.data)
    assert many.shape[0] > 25 and np.allclose(near, g["ref_many_nearest"].numpy(), rtol=0, atol=2e-4)   # cdist's matmul path
    moved = many.copy()
    sweeps = clib.rvo_avoid_pos_rock_collision(moved.ctypes.data, moved.shape[0], s7.ctypes.data, S, 100000)
    assert sweeps > 1
    ref = g["ref_spawn_pos"].numpy()
    assert np.array_equal(moved[:, 1:], ref[:, 1:])                        # only x moves
    # the fixed point is reached in steps of 0.05: equal unless a matmul-path rounding flipped one `<= 1.4` decision
    assert np.abs(moved[:, 0] - ref[:, 0]).max() <= 0.05 + 1e-6 and (moved[:, 0] == ref[:, 0]).mean() >= 0.95


def test_golden_pattern_in_c(clib, golden):
    """Heightmap (heightmap_distribution.py:11-204) in C: the 1634 points and both index vectors equal the reference's bit for bit."""
    cap = 4096
    pts = np.empty((cap, 3), np.float64)
    ci, fi = np.empty(cap, np.int64), np.empty(cap, np.int64)
    nc, nf = C.c_int64(), C.c_int64()
    n = clib.rvo_heightmap_pattern(pts.ctypes.data, cap, ci.ctypes.data, C.addressof(nc), fi.ctypes.data, C.addressof(nf))
    assert (n, nc.value, nf.value) == (1634, 634, 1112)                 # teacher_loader.py:47-48
    assert np.array_equal(pts[:n], golden["ref_pattern"].numpy())
    assert np.array_equal(ci[:nc.value], golden["ref_coarse_idx"].numpy()) and np.array_equal(fi[:nf.value], golden["ref_fine_idx"].numpy())


def test_c_and_torch_oracles_agree_on_random_task_inputs(clib):
    """The two independent restatements (C and torch) on 20 000 random envs -- far more branch combinations than the 8 golden envs:
    Ackermann (incl. 0/0, x/0, tiny angular rates), reward terms and reset masks for both curriculum levels, height lookups."""
    g = torch.Generator().manual_seed(123)
    n = 20_000
    rnd = lambda *shape: torch.rand(*shape, generator=g)
    lin, ang = rnd(n) * 2 - 1, rnd(n) * 2 - 1
    lin[:200], ang[200:400] = 0.0, 0.0
    ang[400:600] = (rnd(200) - 0.5) * 1e-6
    lin[600:700], ang[600:700] = 0.0, 0.0
    steer_t, vel_t = O.ackermann(lin, ang)
    steer, vel = _c_ackermann(clib, lin, ang)
    # torch-CPU's float32 sqrt is not correctly rounded (e.g. sqrt(4.087024211883545) -> 2.0216388702 where IEEE / sqrtf / CUDA give
    # 2.0216391087, the nearer float): the wheel distance, hence the velocity, is one ulp off for ~0.3 % of the inputs
    vt = vel_t.numpy()
    assert np.allclose(vel, vt, rtol=2.5e-7, atol=0, equal_nan=True)
    assert (vel != vt)[~np.isnan(vt)].mean() < 0.01
    assert np.allclose(steer, steer_t.numpy(), rtol=0, atol=2.4e-7, equal_nan=True)

    f = lambda t: np.ascontiguousarray(t.numpy().astype(np.float32))
    pos = torch.cat((rnd(n, 2) * 40, rnd(n, 1)), 1)
    target = pos.clone()
    ang_t, rad = rnd(n) * 6.2831853, torch.where(rnd(n) < 0.1, rnd(n) * 0.3, rnd(n) * 13)      # near-goal and too-far cases
    target[:, 0] += rad * torch.cos(ang_t)
    target[:, 1] += rad * torch.sin(ang_t)
    heading = rnd(n) * 6.4 - 3.2
    lin_p, ang_p = rnd(n) * 2 - 1, rnd(n) * 2 - 1
    same = rnd(n) < 0.2
    lin_p, ang_p = torch.where(same, lin, lin_p), torch.where(same, ang + 0.01, ang_p)           # |delta| around the 0.05 threshold
    joints = rnd(n, 13) * 0.6 - 0.3
    progress = torch.randint(0, 3002, (n,), generator=g)
    coll = (rnd(n) < 0.2).long()
    rot = torch.where(rnd(n, 3) < 0.05, rnd(n, 3) * 2.6 - 1.3, rnd(n, 3) * 0.4 - 0.2)
    for level in (1, 2):
        rew_t, ex_t = O.metrics(pos, target, heading, lin, lin_p, ang, ang_p, joints, progress, coll if level >= 2 else None, level)
        reset_t = O.is_done(pos, target, rot, progress, coll if level >= 2 else None, level)
        rew, reset, ex = np.empty(n, np.float32), np.empty(n, np.int64), np.empty((n, 5), np.float32)
        a = [f(pos), f(target), f(heading), f(lin), f(lin_p), f(ang), f(ang_p), f(joints)]
        pr, co, ro = np.ascontiguousarray(progress.numpy()), np.ascontiguousarray(coll.numpy()), f(rot)
        clib.rvo_reward_reset(*[x.ctypes.data for x in a], 13, pr.ctypes.data, co.ctypes.data if level >= 2 else None, ro.ctypes.data,
                              level, 3000, n, rew.ctypes.data, reset.ctypes.data, ex.ctypes.data)
        # the target distance goes through the same sqrt: position reward / total reward to an ulp, everything else bit for bit
        assert np.allclose(rew, rew_t.numpy(), rtol=1e-6, atol=1e-10), level
        assert (rew != rew_t.numpy()).mean() < 0.02
        assert (reset != reset_t.numpy()).sum() <= 1, level                  # a distance within an ulp of 0.18 / 11 could flip
        for col, key in enumerate(("pos_reward", "heading_contraint_penalty", "motion_contraint_penalty", "goal_angle_penalty",
                                   "uprightness_penalty")):
            ref = ex_t[key].numpy().astype(np.float32)
            if key == "pos_reward":
                assert np.allclose(ex[:, col], ref, rtol=1e-6, atol=0), (level, key)
            else:
                assert np.array_equal(ex[:, col], ref), (level, key)
        assert 0.05 < reset_t.float().mean() < 0.95

    hm = rnd(300, 300)
    xy = rnd(n, 2) * 9 - 1                                                    # includes points off both edges of the 7.5 m grid
    h_t = O.pos_height(hm, xy, 0.025, 1, torch.tensor([0.0, 0.0]))
    out = np.empty(n, np.float32)
    hmn, xyn = f(hm), f(xy)
    clib.rvo_pos_height(hmn.ctypes.data, 300, 300, xyn.ctypes.data, 2, n, 0.025, 1.0, 0.0, 0.0, out.ctypes.data)
    assert np.array_equal(out, h_t.numpy())


def test_c_get_depths_equals_torch_oracle_on_a_fresh_world(clib):
    """Camera.get_depths in C against the torch oracle on a fresh synthetic world with 24 envs (1 % wildly tilted, border poses):
    sources, distances and hit slots bit for bit; same for the rock cast of the oracle's wheel / body rays."""
    import isaac_rover_b200  # noqa: F401  (CPU-side synthetic world helpers only)
    from isaac_rover_b200 import synth
    world = synth.make_world(length=10.0, nv=36, K=48, n_stones=6, seed=77)
    st = synth.make_env_state(world, 24, seed=9, margin=1.0)
    eul = O.quat_to_euler(st["quat"])
    eul[0, 0], eul[1, 1], eul[2, 0] = 1.25, -1.2, 3.0                        # steep, steep, upside down
    pat, _, _ = O.heightmap_pattern()
    ref = O.get_depths(st["pos"], eul, pat, world.map_indices, world.triangles, world.vertices, torch.zeros(3))
    trig = torch.cat(O._neg_trig(eul), 1)
    patn = np.ascontiguousarray(pat.numpy())
    P = patn.shape[0]
    m = np.ascontiguousarray(world.map_indices.to(torch.int32).numpy())
    K, G = m.shape[0], m.shape[1]
    tri = np.ascontiguousarray(world.triangles.to(torch.int32).numpy())
    ver = _u16(world.vertices)
    for e in range(st["pos"].shape[0]):
        pos = np.ascontiguousarray(st["pos"][e].numpy().astype(np.float32))
        tg = np.ascontiguousarray(trig[e].numpy().astype(np.float32))
        src, dist, slot = np.empty((P, 3), np.uint16), np.empty(P, np.uint16), np.empty(P, np.int32)
        clib.rvo_get_depths_env(pos.ctypes.data, tg.ctypes.data, patn.ctypes.data, P, m.ctypes.data, G, K, tri.ctypes.data,
                                ver.ctypes.data, 0.0, 0.0, 0.1, src.ctypes.data, dist.ctypes.data, slot.ctypes.data)
        assert same_halves(src, ref["sources"][e]) and same_halves(dist, ref["dist"][e]), e
        assert np.array_equal(slot, ref["slot"][e].numpy().astype(np.int32)), e
    hits = (ref["dist"].float() < 11.0).float().mean().item()
    assert 0.3 < hits < 1.0
    rk = O.get_collisions(st["pos"], eul, st["joints"], world.rock_indices, world.rock_triangles, world.rock_vertices, torch.zeros(3))
    rm = np.ascontiguousarray(world.rock_indices.to(torch.int32).numpy())
    rtri = np.ascontiguousarray(world.rock_triangles.to(torch.int32).numpy())
    rver = _u16(world.rock_vertices)
    want = torch.cat((rk["wheel"], rk["body"]), 1)
    for e in range(st["pos"].shape[0]):
        s, d = _u16(rk["sources"][e]), _u16(rk["dirs"][e])
        dist, slot = np.empty(26, np.uint16), np.empty(26, np.int32)
        clib.rvo_cast_rays_env(s.ctypes.data, d.ctypes.data, 26, rm.ctypes.data, rm.shape[1], rm.shape[0], rtri.ctypes.data,
                               rver.ctypes.data, 0.0, 0.0, 0.1, dist.ctypes.data, slot.ctypes.data)
        assert same_halves(dist, want[e]), e
