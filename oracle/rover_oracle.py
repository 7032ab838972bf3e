"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the rover per-step hot path.

A restatement, in eager torch ops, of the arithmetic the reference performs on the path
named by BASELINE.json:north_star.  The reference's arithmetic *is* a sequence of torch
ops (Half element-wise ops = fp32 op + one rounding, type promotion f16-f32 -> f32 and
f64*f32 -> f64, torch.round = half-to-even, torch.min(dim) = first index on ties, cdist's
matmul path above 25 rows); torch (2.11.0 here; the reference pins no version, setup.py:12-22)
is therefore the third-party dependency whose semantics define the answer, and this oracle
calls the same ops in the same order.  Every function cites the reference lines it follows.

Pinned: `tests/test_oracle_cpu.py` runs this file against the unmodified reference
imported from /root/reference (bit-exact on CPU) and `tests/golden/*.pt` were produced by the
reference itself (`tests/golden/make_golden.py`).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product (isaac_rover_2.0_b200/) never does.

All functions are device-agnostic: on the GPU box the same code run with device='cuda'
gives the torch-CUDA behaviour of the reference (which hard-wires 'cuda:0', rover.py:90).
"""
import math

import numpy as np
import torch

F16 = torch.float16
PI_F32 = 3.1415927410125732        # tensor_quat_to_euler.py:4  (acos(0)*2 in fp32)


# --------------------------------------------------------------------------------------
# a1  Heightmap pattern   (heightmap_distribution.py:11-204)
# --------------------------------------------------------------------------------------
def _line_test(pt, lines):
    """heightmap_distribution.py:153-193 -- half-plane tests with the reference's quirks
    (the 'left' test uses '<' like 'right' for slanted lines)."""
    x, y = pt
    ok = True
    for p0, p1, side in lines:
        dx, dy = p0[0] - p1[0], p0[1] - p1[1]
        if dx == 0:
            if x < p0[0] and side == 'right':
                ok = False
            if x > p0[0] and side == 'left':
                ok = False
            continue
        a = dy / dx
        b = p0[1] - a * p0[0]
        if a == 0:
            if y > b and side == 'below':
                ok = False
            if y < b and side == 'over':
                ok = False
            continue
        if y < a * x + b and side == 'over':
            ok = False
        if y > a * x + b and side == 'below':
            ok = False
        if x < (y - b) / a and side in ('right', 'left'):
            ok = False
    return ok


def heightmap_pattern():
    """Returns (points f64 [P,3] in rover frame (x fwd, y left after the swap), coarse_idx, fine_idx).
    heightmap_distribution.py:16-30 (constants), :45-59 (coarse fan), :62-78 (fine box, exact-value
    dedup), :100 (round 4), :105 (x<->y swap)."""
    coarse = [[[1.220, 0.118], [4.4455, 3.150], 'over'], [[-1.220, 0.118], [-4.4455, 3.150], 'over'],
              [[1.220, 0.118], [-1.220, 0.118], 'over']]
    fine = [[[1.0, 0.118], [1.0, 0.119], 'left'], [[-1.0, 0.118], [-1.0, 0.119], 'right'],
            [[1.0, 0.118], [-1.0, 0.118], 'over'], [[1.0, 1.400], [-1.0, 1.400], 'below']]
    z = -0.26878
    pts = []
    y = -10
    while y < 10:                      # accumulated floats, exactly as the reference loops do
        x = -10
        while x < 10:
            x += 0.15
            if _line_test([x, y], coarse) and math.sqrt(x * x + y * y) < 3.5:
                pts.append([x, y, z])
        y += 0.15
    coarse_idx = list(range(len(pts)))                 # :57-59 re-tests the same predicate
    seen = set((p[0], p[1]) for p in pts)
    y = -10
    while y < 10:
        x = -10
        while x < 10:
            x += 0.05
            if _line_test([x, y], fine) and (x, y) not in seen:
                pts.append([x, y, z])
                seen.add((x, y))
        y += 0.05
    fine_idx = [i for i, p in enumerate(pts) if _line_test(p[0:2], fine)]
    arr = np.round(np.array(pts, dtype=np.float64), 4)
    arr = arr[:, [1, 0, 2]]
    return torch.from_numpy(arr.copy()), torch.tensor(coarse_idx), torch.tensor(fine_idx)


# --------------------------------------------------------------------------------------
# a2  point transform   (camera.py:165-212)
# --------------------------------------------------------------------------------------
def _body_rotate(x, y, z, sx, cx, sy, cy, sz, cz, tx, ty, tz):
    """camera.py:197-199 / rock_detect.py:305-307,356-358 -- identical expression trees."""
    xp = tx + sz * (y * cx + z * sx) + cz * (x * cy - sy * (z * cx - y * sx))
    yp = ty + cz * (y * cx + z * sx) - sz * (x * cy - sy * (z * cx - y * sx))
    zp = tz + x * sy + cy * (z * cx - y * sx)
    return xp, yp, zp


def _neg_trig(euler):
    """sin/cos of the NEGATED euler angles, fp32, as [N,1] columns (camera.py:184-189)."""
    out = []
    for a in range(3):
        ang = -euler[:, a]
        out += [torch.sin(ang).unsqueeze(1), torch.cos(ang).unsqueeze(1)]
    return out   # sx, cx, sy, cy, sz, cz


def depth_transform(pos, euler, pattern):
    """-> (sources f16 [N,P,3], dir f16 [N,3]).  pattern f64 [P,3]; pos/euler f32.
    The extra point (0,0,-1) is transformed with the others and minus the translation gives the
    ray direction (camera.py:179-181,202-207); fp64 math on fp32 trig (promotion), cast f16 (:212)."""
    dev = pos.device
    extra = torch.tensor([[0, 0, -1]], device=dev, dtype=pattern.dtype)
    p = torch.cat((pattern.to(dev), extra), 0)
    x, y, z = p[:, 0].unsqueeze(0), p[:, 1].unsqueeze(0), p[:, 2].unsqueeze(0)
    sx, cx, sy, cy, sz, cz = _neg_trig(euler)
    tx, ty, tz = pos[:, 0:1], pos[:, 1:2], pos[:, 2:3]
    xp, yp, zp = _body_rotate(x, y, z, sx, cx, sy, cy, sz, cz, tx, ty, tz)
    P = pattern.shape[0]
    d = torch.stack((xp[:, P] - pos[:, 0], yp[:, P] - pos[:, 1], zp[:, P] - pos[:, 2]), 1)
    src = torch.stack((xp[:, :P], yp[:, :P], zp[:, :P]), 2)
    return src.to(F16), d.to(F16)


# --------------------------------------------------------------------------------------
# a3  cell lookup   (camera.py:233-264, rock_detect.py:373-401)
# --------------------------------------------------------------------------------------
def cell_lookup(xy16, shift_xy, res, G):
    """xy16 f16 [...,2], shift f32 [2] -> long cell ids (cx, cy).  f16 - f32 -> f32 (:241),
    clamp to [0, G-1] (:243), round half-even (:245)."""
    s = (xy16 - shift_xy) / res
    s = torch.round(torch.clamp(s, min=0, max=G - 1))
    return s[..., 0].long(), s[..., 1].long()


# --------------------------------------------------------------------------------------
# a5  ray / triangle   (ray_casting.py:3-66)
# --------------------------------------------------------------------------------------
def _cross(u, v):
    """Tensor.cross on [n,3] Half: each component is (a*b) - (c*d) with every op rounded to fp16."""
    return torch.stack((u[:, 1] * v[:, 2] - u[:, 2] * v[:, 1],
                        u[:, 2] * v[:, 0] - u[:, 0] * v[:, 2],
                        u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]), 1)


def _dot3(u, v):
    return u[:, 0] * v[:, 0] + u[:, 1] * v[:, 1] + u[:, 2] * v[:, 2]     # left to right (:41)


def ray_distance(src, dirs, tri):
    """n rays vs n triangles, all fp16.  src,dirs [n,3]; tri [n,3,3] (vertex, xyz) -> (k [n], pt [n,3]).
    eps-inflated barycentric test, miss sentinel 11.0, guards compare det against the *shifted*
    constants (:46,51,56), no k>=0 test (:59)."""
    n = src.shape[0]
    dev = src.device
    lo = torch.zeros(n, device=dev, dtype=F16) - 0.1            # -0.0999755859375
    hi = torch.ones(n, device=dev, dtype=F16) + 0.1             #  1.099609375
    miss = hi * 10.0                                            #  11.0
    d = -torch.nn.functional.normalize(dirs)                    # :31  (fp32-accumulated norm, eps 1e-12)
    a = tri[:, 2]
    b = tri[:, 1] - a
    c = tri[:, 0] - a
    g = src - a
    bxc = _cross(b, c)
    det = _dot3(bxc, d)
    nn = _dot3(_cross(g, c), d) / det
    nn = torch.where(det == lo, miss, nn)
    mm = _dot3(_cross(b, g), d) / det
    mm = torch.where(det == hi, miss, mm)
    kk = _dot3(bxc, g) / det
    kk = torch.where(det == hi, miss, kk)
    k = torch.where((nn >= lo) & (mm >= lo) & (nn + mm <= hi), kk, miss)
    pt = src - d * k.unsqueeze(1)                               # :63
    return k, pt


def _cast_against_index(src16, dir16, map_idx_gkk, triangles, vertices, shift_xy, res, ray_chunk):
    """Shared by a4 and a8: per ray look up the K-list of its cell, gather the K triangles, run
    ray_distance on every (ray, candidate) pair, min over K (first index on ties).
    src16 [N,R,3], dir16 [N,R,3] fp16.  -> dist f16 [N,R], slot i64 [N,R], tri_id i64 [N,R], pt f16 [N,R,3]."""
    N, R, _ = src16.shape
    G, K = map_idx_gkk.shape[0], map_idx_gkk.shape[2]
    cx, cy = cell_lookup(src16[:, :, 0:2], shift_xy, res, G)
    dist = torch.empty((N, R), dtype=F16, device=src16.device)
    slot = torch.empty((N, R), dtype=torch.long, device=src16.device)
    tid = torch.empty((N, R), dtype=torch.long, device=src16.device)
    pts = torch.empty((N, R, 3), dtype=F16, device=src16.device)
    for r0 in range(0, R, ray_chunk):
        r1 = min(r0 + ray_chunk, R)
        ids = map_idx_gkk[cx[:, r0:r1], cy[:, r0:r1]].long()              # [N,r,K]
        tri = vertices[triangles[ids].long()]                             # [N,r,K,3,3]
        r = r1 - r0
        s = src16[:, r0:r1].unsqueeze(2).expand(N, r, K, 3).reshape(-1, 3)
        d = dir16[:, r0:r1].unsqueeze(2).expand(N, r, K, 3).reshape(-1, 3)
        k, pt = ray_distance(s, d, tri.reshape(-1, 3, 3))
        k = k.reshape(N, r, K)
        mn = torch.min(k, 2)
        dist[:, r0:r1] = mn.values
        slot[:, r0:r1] = mn.indices
        tid[:, r0:r1] = ids.gather(2, mn.indices.unsqueeze(2)).squeeze(2)
        pts[:, r0:r1] = pt.reshape(N, r, K, 3).gather(2, mn.indices[:, :, None, None].expand(N, r, 1, 3)).squeeze(2)
    return dist, slot, tid, pts


def permute_index(map_indices_kgg):
    """[K,G,G] -> [G,G,K] view (camera.py:157-158)."""
    return map_indices_kgg.swapaxes(0, 1).swapaxes(1, 2)


# --------------------------------------------------------------------------------------
# a4  Camera.get_depths   (camera.py:60-145)
# --------------------------------------------------------------------------------------
def get_depths(pos, euler, pattern, map_indices_kgg, triangles, vertices, shift, res=0.1,
               env_chunk=16, ray_chunk=408):
    """-> dict(dist f16 [N,P], slot, tri, pt f16 [N,P,3], sources f16 [N,P,3]).
    The reference's 4(+1)-way ray partition (:77-84) only bounds memory; results are per ray."""
    idx = permute_index(map_indices_kgg)
    outs = []
    for e0 in range(0, pos.shape[0], env_chunk):
        e = slice(e0, e0 + env_chunk)
        src, d = depth_transform(pos[e], euler[e], pattern)
        dirs = d.unsqueeze(1).expand(-1, src.shape[1], -1)
        dist, slot, tid, pt = _cast_against_index(src, dirs, idx, triangles, vertices, shift[0:2], res, ray_chunk)
        outs.append((dist, slot, tid, pt, src))
    cat = [torch.cat(x, 0) for x in zip(*outs)]
    return dict(dist=cat[0], slot=cat[1], tri=cat[2], pt=cat[3], sources=cat[4])


# --------------------------------------------------------------------------------------
# a6-a8  rock rays   (rock_detect.py:160-371, 52-149)
# --------------------------------------------------------------------------------------
_WHEEL_RAYS = [[0.215 / 2, 0.130 / 2, 0.1], [0.215 / 2, -0.130 / 2, 0.1], [-0.215 / 2, 0.130 / 2, 0.1],
               [-0.215 / 2, -0.130 / 2, 0.1], [0, 0, -1]]                          # :193-197 (last = direction)
_WHEEL_POS0 = [[0.286, 0.385, -0.197], [0.286, -0.385, -0.197], [-0.146, 0.447, -0.197],
               [-0.146, -0.447, -0.197], [-0.440, 0.385, -0.197], [-0.440, -0.385, -0.197]]   # :201-206
_WHEEL_POS1 = [[0.153, 0, 0.03], [0.153, 0, 0.03], [0.153, 0, 0.03], [0.153, -0, 0.03],
               [0, 0, 0.03], [0, 0, 0.03]]                                         # :210-215


def wheel_rays(pos, euler, joints):
    """-> (sources f16 [N,24,3], dirs f16 [N,24,3]).  30 = 6 wheels x (4 points + 1 direction entry);
    direction entries get zero translation at every stage (:227,232,293)."""
    dev = pos.device
    N = pos.shape[0]
    rays = torch.tensor(_WHEEL_RAYS, device=dev).repeat(6, 1)                       # [30,3]
    isdir = (torch.arange(30, device=dev) % 5 == 4)
    t0 = torch.tensor(_WHEEL_POS0, device=dev).repeat_interleave(5, 0)
    t1 = torch.tensor(_WHEEL_POS1, device=dev).repeat_interleave(5, 0)
    t0[isdir] = 0
    t1[isdir] = 0
    x, y, z = rays[:, 0].unsqueeze(0), rays[:, 1].unsqueeze(0), rays[:, 2].unsqueeze(0)
    j = joints
    zero = torch.zeros_like(j[:, 0])
    steer = torch.stack((j[:, 4], j[:, 6], zero, zero, -j[:, 7], j[:, 8]), 1).repeat_interleave(5, 1)   # :248
    ss, cs = torch.sin(-steer), torch.cos(-steer)
    x1 = t0[:, 0] + x * cs + y * ss                                                  # :256-258
    y1 = t0[:, 1] + y * cs - x * ss
    z1 = t0[:, 2] + z
    susy = torch.stack((-j[:, 0], j[:, 1], -j[:, 0], j[:, 1], zero, zero), 1).repeat_interleave(5, 1)   # :263
    susx = torch.stack((zero, zero, zero, zero, -j[:, 2], -j[:, 2]), 1).repeat_interleave(5, 1)        # :264
    sux, cux, suy, cuy = torch.sin(susx), torch.cos(susx), torch.sin(susy), torch.cos(susy)
    x2 = t1[:, 0] + x1 * cuy - suy * (z1 * cux - y1 * sux)                           # :275-277
    y2 = t1[:, 1] + y1 * cux + z1 * sux
    z2 = t1[:, 2] + x1 * suy + cuy * (z1 * cux - y1 * sux)
    sx, cx, sy, cy, sz, cz = _neg_trig(euler)
    zcol = torch.zeros((N, 30), device=dev)
    tx = torch.where(isdir.unsqueeze(0), zcol, pos[:, 0:1].expand(N, 30))           # translation or exact +0 (:292-302)
    ty = torch.where(isdir.unsqueeze(0), zcol, pos[:, 1:2].expand(N, 30))
    tz = torch.where(isdir.unsqueeze(0), zcol, pos[:, 2:3].expand(N, 30))
    xp, yp, zp = _body_rotate(x2, y2, z2, sx, cx, sy, cy, sz, cz, tx, ty, tz)
    allp = torch.stack((xp, yp, zp), 2)                                              # [N,30,3]
    dirs = allp[:, 4::5, :].repeat_interleave(4, 1)                                  # :314
    src = allp.reshape(-1, 5, 3)[:, :4].reshape(N, 24, 3)                            # :317
    return src.to(F16), dirs.to(F16)


def body_rays(pos, euler):
    """-> (sources f16 [N,2,3], dirs f16 [N,2,3]); direction = local (0,1,0) (:338-340)."""
    dev = pos.device
    p = torch.tensor([[0.340, 0, -0.01], [-0.485, 0, -0.01], [0, 1, 0]], device=dev)
    x, y, z = p[:, 0].unsqueeze(0), p[:, 1].unsqueeze(0), p[:, 2].unsqueeze(0)
    sx, cx, sy, cy, sz, cz = _neg_trig(euler)
    xp, yp, zp = _body_rotate(x, y, z, sx, cx, sy, cy, sz, cz, pos[:, 0:1], pos[:, 1:2], pos[:, 2:3])
    d = torch.stack((xp[:, 2] - pos[:, 0], yp[:, 2] - pos[:, 1], zp[:, 2] - pos[:, 2]), 1)
    src = torch.stack((xp[:, :2], yp[:, :2], zp[:, :2]), 2)
    return src.to(F16), d.unsqueeze(1).repeat(1, 2, 1).to(F16)


def get_collisions(pos, euler, joints, rock_indices_kgg, triangles, vertices, shift, res=0.1, env_chunk=512):
    """-> dict(wheel f16 [N,24], body f16 [N,2], slot, tri) (rock_detect.py:52-149)."""
    idx = permute_index(rock_indices_kgg)
    outs = []
    for e0 in range(0, pos.shape[0], env_chunk):
        e = slice(e0, e0 + env_chunk)
        s1, d1 = wheel_rays(pos[e], euler[e], joints[e])
        s2, d2 = body_rays(pos[e], euler[e])
        src, dirs = torch.cat((s1, s2), 1), torch.cat((d1, d2), 1)
        dist, slot, tid, _ = _cast_against_index(src, dirs, idx, triangles, vertices, shift[0:2], res, 26)
        outs.append((dist, slot, tid, src, dirs))
    dist, slot, tid, src, dirs = [torch.cat(x, 0) for x in zip(*outs)]
    return dict(wheel=dist[:, :24], body=dist[:, 24:], slot=slot, tri=tid, sources=src, dirs=dirs)


def check_collision(wheel, body):
    """rover.py:663-668 -> i64 [N]."""
    w = torch.min(wheel, dim=1)[0]
    b = torch.min(body, dim=1)[0]
    one = torch.ones(wheel.shape[0], dtype=torch.long, device=wheel.device)
    col = torch.where(torch.abs(w) < 0.8, one, torch.zeros_like(one))
    return torch.where(torch.abs(b) < 0.45, one, col)


# --------------------------------------------------------------------------------------
# a10  quaternion -> euler   (tensor_quat_to_euler.py:6-31)
# --------------------------------------------------------------------------------------
def quat_to_euler(q):
    dev = q.device
    n = q.shape[0]
    one = torch.ones(n, device=dev)
    zero = torch.zeros(n, device=dev)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    e = torch.zeros((n, 3), device=dev)
    e[:, 0] = torch.atan2(2 * (w * x + y * z), one - (2 * (x * x + y * y)))
    sinp = 2 * (w * y - z * x)
    e[:, 1] = torch.where(torch.sign(sinp - one) >= zero, torch.copysign((one * PI_F32) / 2, sinp), torch.asin(sinp))
    e[:, 2] = torch.atan2(2 * (w * z + x * y), one - (2 * (y * y + z * z)))
    return e


# --------------------------------------------------------------------------------------
# a13/a14  Ackermann + joint-target mapping   (kinematics.py:14-67, rover.py:396-409)
# --------------------------------------------------------------------------------------
_WHEEL_XY = [[-0.385, 0.438], [0.385, 0.438], [-0.447, 0.0], [0.447, 0.0], [-0.385, -0.411], [0.385, -0.411]]


def ackermann(lin, ang):
    """-> (steer [N,6], vel [N,6]) fp32, wheel order FL,FR,ML,MR,RL,RR."""
    dev = lin.device
    wxy = torch.tensor(_WHEEL_XY, device=dev)                       # [6,2]
    P = torch.copysign(lin / ang, -ang)                             # :34-35
    zero = torch.zeros_like(P)
    P = torch.where(torch.abs(P) > 0.45, P, zero)                   # :38
    lin = torch.where(P != 0, lin, zero)                            # :39
    dxy = torch.stack((P.unsqueeze(0) - wxy[:, 0:1], zero.unsqueeze(0) - wxy[:, 1:2]), 2)   # [6,N,2] (:42-43)
    dist = dxy.pow(2).sum(2).sqrt().t()                             # [N,6]
    side = torch.tensor([-1.0, 1.0, -1.0, 1.0, -1.0, 1.0], device=dev)
    w_lin = torch.copysign(ang, lin).unsqueeze(1).expand(-1, 6)     # :49
    w_turn = ang.unsqueeze(1) * side                                # :51
    lin6 = lin.unsqueeze(1).expand(-1, 6)
    omega = torch.where(lin6 != 0, w_lin, w_turn)                   # :52
    vel = dist * omega
    vel = torch.where(dist > 1000, lin6, vel)                       # :58
    vel = vel / 0.2                                                 # :61
    steer = torch.atan2(wxy[:, 1].unsqueeze(0).expand(lin.shape[0], -1), wxy[:, 0].unsqueeze(0) - P.unsqueeze(1))   # :63
    steer = torch.where(steer < -3.14 / 2, steer + math.pi, steer)  # :64
    steer = torch.where(steer > 3.14 / 2, steer - math.pi, steer)   # :65
    return steer, vel


def joint_targets(steer, vel):
    """rover.py:400-409: positions (FR,RR,FL,RL), velocities (FR,CR,RR,FL,CL,RL)."""
    return steer[:, [1, 5, 0, 4]].contiguous(), vel[:, [1, 3, 5, 0, 2, 4]].contiguous()


# --------------------------------------------------------------------------------------
# a11  observation assembly   (rover.py:272-336)
# --------------------------------------------------------------------------------------
def heading_and_dist(pos, euler, target):
    yaw = euler[:, 2]
    dx, dy = torch.cos(yaw), torch.sin(yaw)
    tv = target[:, 0:2] - pos[:, 0:2]
    heading = -torch.atan2(tv[:, 0] * dy - tv[:, 1] * dx, tv[:, 0] * dx + tv[:, 1] * dy)     # :283
    return heading, torch.linalg.norm(tv, dim=1)


def observations(pos, quat, target, lin_now, ang_now, dist16, coarse_idx, fine_idx):
    """-> (obs f32 [N,4+ns+nd], euler, heading)."""
    euler = quat_to_euler(quat)
    heading, tnorm = heading_and_dist(pos, euler, target)
    ns, nd = coarse_idx.shape[0], fine_idx.shape[0]
    obs = torch.zeros((pos.shape[0], 4 + ns + nd), device=pos.device)
    obs[:, 0] = tnorm / 9
    obs[:, 1] = heading / math.pi
    obs[:, 2] = lin_now
    obs[:, 3] = ang_now
    obs[:, 4:4 + ns] = dist16[:, coarse_idx.to(dist16.device)] / 2          # fp16 halving, then stored as f32 (:324)
    obs[:, 4 + ns:] = dist16[:, fine_idx.to(dist16.device)] / 2
    return obs, euler, heading


# --------------------------------------------------------------------------------------
# a15/a16  reward terms and resets   (rover.py:460-531, 610-647)
# --------------------------------------------------------------------------------------
DEFAULT_REW = dict(pos_reward=1.0, terminalReward=0, collision_reward=0.3, heading_contraint_reward=0.05,
                   motion_contraint_reward=-0.01, goal_angle_reward=0.3, boogie_contraint_reward=0.5)   # Rover.yaml:37-46


def metrics(pos, target, heading, lin, lin_prev, ang, ang_prev, joints, progress, rock_collision, level,
            rew=DEFAULT_REW, max_len=3000, num_envs=None):
    """-> (rew_buf f32 [N], extras dict).  `rock_collision` i64 [N] or None."""
    N = pos.shape[0]
    num_envs = N if num_envs is None else num_envs
    izero = torch.zeros(N, dtype=torch.long, device=pos.device)
    ione = torch.ones(N, dtype=torch.long, device=pos.device)
    fzero = izero.float()
    td = torch.sqrt(torch.square(target[:, 0:2] - pos[:, 0:2]).sum(-1))                      # :482
    heading_pen = torch.where(lin < 0, -ione, izero) * rew["heading_contraint_reward"]       # :486
    boogie = (torch.abs(joints[:, 0]) + torch.abs(joints[:, 1]) + torch.abs(joints[:, 2])) * rew["boogie_contraint_reward"]
    goal_pen = torch.where(torch.abs(heading) > 2, -torch.abs(heading * 0.3 * rew["goal_angle_reward"]), fzero)   # :495
    dl = torch.abs(lin * 3 - 3 * lin_prev)
    da = torch.abs(ang * 3 - 3 * ang_prev)
    p1 = torch.where(dl > 0.05, torch.square(dl), fzero)
    p2 = torch.where(da > 0.05, torch.square(da), fzero)
    motion = torch.pow(p1, 2) * rew["motion_contraint_reward"]
    motion = motion + (torch.pow(p2, 2)) * rew["motion_contraint_reward"]                    # :500-502
    pos_rew = (1.0 / (1.0 + (0.33 * 0.33 * td * td))) * rew["pos_reward"]                    # :505
    pos_rew = torch.where(td <= 0.18, 1.03 * (max_len - progress), pos_rew.float())          # :506
    reward = pos_rew + heading_pen + motion + goal_pen                                        # :512
    if level >= 2:
        tracker = torch.where(rock_collision == 1, ione * num_envs, izero)                   # :517
        reward = torch.where(rock_collision == 1, reward - 300, reward)
    else:
        tracker = izero
    reward = reward / 3000
    extras = dict(pos_reward=pos_rew, collision_penalty=tracker, uprightness_penalty=boogie,
                  heading_contraint_penalty=heading_pen, motion_contraint_penalty=motion,
                  goal_angle_penalty=goal_pen, torque_penalty_driving=lin, torque_penalty_steering=ang)
    return reward.float(), extras


def is_done(pos, target, rover_rot, progress, rock_collision, level, max_len=3000):
    """-> reset i64 [N].  `rover_rot` is the PRE-physics euler (rover.py:343,615-616)."""
    one = torch.ones_like(progress)
    r = torch.where(progress >= max_len, one, torch.zeros_like(progress))
    r = torch.where(torch.abs(rover_rot[:, 0]) >= 0.78 * 1.5, one, r)
    r = torch.where(torch.abs(rover_rot[:, 1]) >= 0.78 * 1.5, one, r)
    td = torch.sqrt(torch.square(target[:, 0:2] - pos[:, 0:2]).sum(-1))
    r = torch.where(td >= 11, one, r)
    r = torch.where(td <= 0.18, one, r)
    if level >= 2:
        r = torch.where(rock_collision == 1, one, r)
    return r


# --------------------------------------------------------------------------------------
# a17-a20  stones and height grid   (terrain_utils.py:416-424, rover.py:533-542,588-608,649-661)
# --------------------------------------------------------------------------------------
def read_stone_info(arr):
    """npy [S,6] -> f32 [S,7]; col 6 = max(col3, col4)/4 computed in the file's dtype, then cast."""
    n = np.asarray(arr)
    rs = np.zeros((len(n), 1))
    for i in range(len(n)):
        rs[i][0] = max(n[i][3], n[i][4]) / 4
    return torch.from_numpy(np.append(n, rs, axis=1)).float()


def nearest_stone_edge(xy, stone7):
    """min over stones of (cdist - radius)  (rover.py:536-538 / :655-657).  torch.cdist picks its
    matmul formulation by itself above 25 rows; calling it keeps that switch."""
    d = torch.cdist(xy, stone7[:, 0:2], p=2.0)
    d[:] = d[:] - stone7[:, 6]
    return torch.min(d, dim=1)[0]


def goal_invalid(target_xy, stone7):
    return (nearest_stone_edge(target_xy, stone7) <= 1.0).long()


def avoid_pos_rock_collision(pos, stone7, max_iter=100000):
    """Fixed-point loop: x += 0.05 for every env within 1.4 m of a stone edge (rover.py:649-661)."""
    cur = pos.clone()
    old = torch.zeros_like(cur)
    it = 0
    while not torch.equal(cur, old):
        old = cur.clone()
        near = nearest_stone_edge(cur[:, 0:2], stone7)
        cur[:, 0] = torch.where(near <= 1.4, torch.add(cur[:, 0], 0.05), cur[:, 0])
        it += 1
        if it > max_iter:
            raise RuntimeError("spawn validation did not converge")
    return cur


def pos_height(heightmap, xy, hscale, vscale, shift_xy):
    """rover.py:588-608."""
    s = torch.round(torch.clamp((xy - shift_xy) / hscale, min=0, max=heightmap.size()[0] - 1))
    return heightmap[s[:, 0].long(), s[:, 1].long()] * vscale


# --------------------------------------------------------------------------------------
# whole env-step (used as the CPU baseline / reference arm by bench.py and by smoke())
# --------------------------------------------------------------------------------------
def full_step(assets, st, level=2, env_chunk=16):
    """One env-step of the hot path with PhysX excluded, on whatever device the tensors live on:
    pre_physics_step's action half (rover.py:343,366-409) then post_physics_step (rl_task.py:239-259).
    assets: dict(pattern, coarse_idx, fine_idx, map_indices, triangles, vertices, rock_indices, rock_triangles,
    rock_vertices, shift); st: dict(pos, quat, joints, actions, prev_actions, target, progress)."""
    a = assets
    rover_rot = quat_to_euler(st["quat"])                                   # rover.py:343
    lin, ang = st["actions"][:, 0], st["actions"][:, 1]
    steer, vel = ackermann(lin, ang)
    pos_t, vel_t = joint_targets(steer, vel)
    progress = st["progress"] + 1                                           # rl_task.py:250
    euler = quat_to_euler(st["quat"])
    dep = get_depths(st["pos"], euler, a["pattern"], a["map_indices"], a["triangles"], a["vertices"], a["shift"],
                     env_chunk=env_chunk)
    obs, _, heading = observations(st["pos"], st["quat"], st["target"], lin, ang, dep["dist"], a["coarse_idx"], a["fine_idx"])
    col = get_collisions(st["pos"], euler, st["joints"], a["rock_indices"], a["rock_triangles"], a["rock_vertices"], a["shift"])
    rock = check_collision(col["wheel"], col["body"]) if level >= 2 else None
    rew, extras = metrics(st["pos"], st["target"], heading, lin, st["prev_actions"][:, 0], ang, st["prev_actions"][:, 1],
                          st["joints"], progress, rock, level)
    reset = is_done(st["pos"], st["target"], rover_rot, progress, rock, level)
    return dict(obs=obs, rew=rew, reset=reset, extras=extras, rock_collision=rock, dist=dep["dist"], tri=dep["tri"],
                wheel=col["wheel"], body=col["body"], pos_targets=pos_t, vel_targets=vel_t, heading=heading, euler=euler)
