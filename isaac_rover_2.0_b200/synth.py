"""Seeded synthetic worlds and per-env state for the rover hot path.

The reference's terrain assets (map.ply, big_stones.ply, knn_terrain/*.pt, knn_rocks/*.pt,
stone_info.npy, heightmap_tensor.pt) are git-LFS blobs that are not available, so every
input of the hot path is synthesised here in the *formats* the reference consumes
(camera.py:154-161, rock_detect.py:151-158, terrain_utils.py:416-424, rover.py:210-212):

  vertices   fp16 [V,3]     triangles int32 [T,3]     map_indices int32 [K,G,G]
  stone_info f32  [S,6] (x,y,z,dx,dy,dz)              heightmap   f32  [H,H] @ hm_res

The mesh follows the heightfield->trimesh layout of terrain_utils.py:349-369 (two triangles
per quad, (i0,i3,i1) and (i0,i2,i3)).  Everything is deterministic given `seed`.
"""
import math
from dataclasses import dataclass, field

import numpy as np
import torch


@dataclass
class World:
    length: float                 # side of the square map [m]
    res: float                    # index cell size [m] (reference: 0.1)
    G: int                        # cells per side
    K: int                        # candidates per cell (reference: 200)
    vertices: torch.Tensor        # fp16 [V,3]
    triangles: torch.Tensor       # int32 [T,3]
    rock_vertices: torch.Tensor   # fp16 [Vr,3]
    rock_triangles: torch.Tensor  # int32 [Tr,3]
    stone_info: torch.Tensor      # f32 [S,6]
    heightmap: torch.Tensor       # f32 [H,H]
    hm_res: float
    heightfield: np.ndarray = field(repr=False, default=None)   # f32 [nv,nv] (generator detail)
    hf_spacing: float = 0.0
    map_indices: torch.Tensor = None        # int32 [K,G,G]   (terrain layer)
    rock_indices: torch.Tensor = None       # int32 [K,G,G]   (big_rock_layer)


def make_heightfield(length, nv, n_stones, seed):
    """Rolling sinusoids + uniform roughness + gaussian rock bumps (idea: terrain_generation.py:18-65,104-153)."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(0.0, length, nv, dtype=np.float64)
    xx, yy = np.meshgrid(xs, xs, indexing="ij")
    hf = np.zeros((nv, nv), dtype=np.float64)
    for _ in range(4):
        wl = rng.uniform(6.0, 40.0)
        ang = rng.uniform(0, 2 * math.pi)
        amp = rng.uniform(0.05, 0.35)
        ph = rng.uniform(0, 2 * math.pi)
        hf += amp * np.sin((xx * math.cos(ang) + yy * math.sin(ang)) * (2 * math.pi / wl) + ph)
    hf += rng.uniform(-0.02, 0.02, size=hf.shape)
    # stones: (x, y, z, dx, dy, dz)
    stones = np.zeros((n_stones, 6), dtype=np.float64)
    stones[:, 0] = rng.uniform(2.0, length - 2.0, n_stones)
    stones[:, 1] = rng.uniform(2.0, length - 2.0, n_stones)
    stones[:, 3] = rng.uniform(0.2, 3.0, n_stones)
    stones[:, 4] = stones[:, 3] * rng.uniform(0.7, 1.3, n_stones)
    stones[:, 5] = np.minimum(stones[:, 3], stones[:, 4]) * rng.uniform(0.3, 0.7, n_stones)
    sp = length / (nv - 1)
    for s in range(n_stones):
        cx, cy, dx, dy, dz = stones[s, 0], stones[s, 1], stones[s, 3], stones[s, 4], stones[s, 5]
        r = int(math.ceil(1.5 * max(dx, dy) / sp)) + 1
        i0, j0 = int(round(cx / sp)), int(round(cy / sp))
        ia, ib = max(i0 - r, 0), min(i0 + r + 1, nv)
        ja, jb = max(j0 - r, 0), min(j0 + r + 1, nv)
        px = xx[ia:ib, ja:jb] - cx
        py = yy[ia:ib, ja:jb] - cy
        hf[ia:ib, ja:jb] += dz * np.exp(-((px / (0.35 * dx)) ** 2 + (py / (0.35 * dy)) ** 2))
        stones[s, 2] = hf[min(max(i0, 0), nv - 1), min(max(j0, 0), nv - 1)]
    return hf.astype(np.float32), stones.astype(np.float32)


def heightfield_to_mesh(hf, length):
    """Vertices f32 [nv*nv,3] and triangles int32 [2(nv-1)^2,3]; row-major vertex ids, x = row."""
    nv = hf.shape[0]
    xs = np.linspace(0.0, length, nv, dtype=np.float64)
    v = np.empty((nv * nv, 3), dtype=np.float32)
    v[:, 0] = np.repeat(xs, nv)
    v[:, 1] = np.tile(xs, nv)
    v[:, 2] = hf.reshape(-1)
    i, j = np.meshgrid(np.arange(nv - 1), np.arange(nv - 1), indexing="ij")
    i0 = (i * nv + j).reshape(-1)
    i1, i2 = i0 + 1, i0 + nv
    i3 = i2 + 1
    t = np.empty((i0.size * 2, 3), dtype=np.int32)
    t[0::2] = np.stack((i0, i3, i1), 1)
    t[1::2] = np.stack((i0, i2, i3), 1)
    return v, t


def rock_layer(vertices, triangles, stones, min_size=1.5, nv=None):
    """Sub-mesh made of the triangles under the big stones (the reference's big_stones.ply role).
    With `nv` (heightfield mesh of nv x nv vertices) only the quads around each stone are examined."""
    keep = np.zeros(len(triangles), dtype=bool)
    if nv is not None:
        sp = float(vertices[1, 1] - vertices[0, 1])
        for s in stones:
            if max(s[3], s[4]) < min_size:
                continue
            ri, rj = 0.6 * s[3] / sp + 2, 0.6 * s[4] / sp + 2
            ia, ib = int(max(s[0] / sp - ri, 0)), int(min(s[0] / sp + ri + 1, nv - 1))
            ja, jb = int(max(s[1] / sp - rj, 0)), int(min(s[1] / sp + rj + 1, nv - 1))
            if ia >= ib or ja >= jb:
                continue
            ii, jj = np.meshgrid(np.arange(ia, ib), np.arange(ja, jb), indexing="ij")
            tid = (2 * (ii * (nv - 1) + jj)).reshape(-1)
            tid = np.concatenate((tid, tid + 1))
            cen = vertices[triangles[tid].astype(np.int64)].mean(1)[:, :2]
            d = ((cen[:, 0] - s[0]) / (0.6 * s[3])) ** 2 + ((cen[:, 1] - s[1]) / (0.6 * s[4])) ** 2
            keep[tid[d <= 1.0]] = True
    else:
        cen = vertices[triangles.astype(np.int64)].mean(1)[:, :2]
        for s in stones:
            if max(s[3], s[4]) < min_size:
                continue
            d = ((cen[:, 0] - s[0]) / (0.6 * s[3])) ** 2 + ((cen[:, 1] - s[1]) / (0.6 * s[4])) ** 2
            keep |= d <= 1.0
    tri = triangles[keep]
    if len(tri) < 2:                        # keep the layer non-empty so every cell has candidates
        tri = triangles[:2]
    used, inv = np.unique(tri.reshape(-1), return_inverse=True)
    return vertices[used].copy(), inv.reshape(-1, 3).astype(np.int32)


def knn_index_bruteforce(vertices16, triangles, G, res, K, cell_chunk=2048):
    """[K,G,G] int32 nearest-centroid index, restating rover_utils.py:52-118 on the CPU.

    fp16 2-D centroids, fp16 cell coordinates (cell (i,j) at (i*res, j*res)), fp16 difference,
    L2 norm accumulated in fp32 and rounded to fp16, K smallest.  `topk` leaves the order of
    equal distances unspecified; this builder (and the CUDA one) fixes it as (distance, id).
    Brute force over all T -- small worlds only (tests, golden fixtures)."""
    v = vertices16.double()
    t = triangles.long()
    cen = ((v[t[:, 0]] + v[t[:, 1]] + v[t[:, 2]]) / 3)[:, :2].to(torch.float16)       # [T,2]
    T = cen.shape[0]
    assert T >= K, "need at least K triangles"
    coord = torch.arange(0, G * res, res, dtype=torch.float16)[:G]
    out = torch.empty((K, G * G), dtype=torch.int32)
    ids = torch.arange(T, dtype=torch.int64)
    cells = torch.arange(G * G)
    for c0 in range(0, G * G, cell_chunk):
        c = cells[c0:c0 + cell_chunk]
        px = coord[c // G].unsqueeze(1)
        py = coord[c % G].unsqueeze(1)
        dx = (cen[:, 0].unsqueeze(0) - px).float()
        dy = (cen[:, 1].unsqueeze(0) - py).float()
        d16 = torch.sqrt(dx * dx + dy * dy).to(torch.float16)
        key = (d16.view(torch.int16).long() << 32) | ids.unsqueeze(0)                 # dist >= 0: bit order == value order
        sel = torch.topk(key, K, dim=1, largest=False, sorted=True).values & 0xFFFFFFFF
        out[:, c0:c0 + cell_chunk] = sel.t().to(torch.int32)
    return out.reshape(K, G, G)


def heightmap_grid(hf, length, hm_res):
    """f32 [H,H] nearest-sample height grid (the role of heightmap_tensor.pt, rover.py:210-212)."""
    nv = hf.shape[0]
    H = int(round(length / hm_res))
    idx = np.clip(np.rint(np.arange(H) * hm_res / (length / (nv - 1))).astype(np.int64), 0, nv - 1)
    return torch.from_numpy(hf[np.ix_(idx, idx)].copy())


def make_world(length=20.0, nv=72, K=200, res=0.1, n_stones=40, hm_res=0.025, seed=42,
               build_index="cpu"):
    """Build a synthetic world.  build_index: "cpu" (brute force, small worlds) or None
    (caller builds the index on the GPU with `terrain.build_knn_index`)."""
    hf, stones = make_heightfield(length, nv, n_stones, seed)
    v, t = heightfield_to_mesh(hf, length)
    rv, rt = rock_layer(v, t, stones, nv=nv)
    G = int(round(length / res))
    w = World(length=length, res=res, G=G, K=K,
              vertices=torch.from_numpy(v).to(torch.float16), triangles=torch.from_numpy(t),
              rock_vertices=torch.from_numpy(rv).to(torch.float16), rock_triangles=torch.from_numpy(rt),
              stone_info=torch.from_numpy(stones), heightmap=heightmap_grid(hf, length, hm_res),
              hm_res=hm_res, heightfield=hf, hf_spacing=length / (nv - 1))
    if build_index == "cpu":
        w.map_indices = knn_index_bruteforce(w.vertices, w.triangles, G, res, K)
        kr = min(K, w.rock_triangles.shape[0])
        ri = knn_index_bruteforce(w.rock_vertices, w.rock_triangles, G, res, kr)
        if kr < K:      # tiny rock layers: pad by repeating the farthest candidate (keeps [K,G,G])
            ri = torch.cat((ri, ri[-1:].expand(K - kr, -1, -1)), 0).contiguous()
        w.rock_indices = ri
    return w


def terrain_height(world, xy):
    """Nearest-vertex terrain height (f32) at xy [N,2] (generator-side helper for spawn z)."""
    nv = world.heightfield.shape[0]
    ij = np.clip(np.rint(xy.numpy().astype(np.float64) / world.hf_spacing).astype(np.int64), 0, nv - 1)
    return torch.from_numpy(world.heightfield[ij[:, 0], ij[:, 1]].copy())


def euler_to_quat_wxyz(roll, pitch, yaw):
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    return torch.stack((cr * cp * cy + sr * sp * sy,
                        sr * cp * cy - cr * sp * sy,
                        cr * sp * cy + sr * cp * sy,
                        cr * cp * sy - sr * sp * cy), 1).float()


def make_env_state(world, N, seed=42, margin=5.0, env_offset=0):
    """Per-env inputs.  Env i's values depend only on (seed, env_offset+i) blocks of a
    single generator stream, so a rank's shard [r*N/W,(r+1)*N/W) equals the same rows of the global state."""
    g = torch.Generator().manual_seed(seed)
    total = env_offset + N
    L = world.length
    m = min(margin, L * 0.25)

    def u(lo, hi, *shape):
        return torch.rand(*shape, generator=g) * (hi - lo) + lo

    xy = u(m, L - m, total, 2)
    z = terrain_height(world, xy) + 0.5 + u(0.0, 0.05, total)
    pos = torch.cat((xy, z.unsqueeze(1)), 1).float()
    roll, pitch = u(-0.2, 0.2, total), u(-0.2, 0.2, total)
    wild = torch.rand(total, generator=g) < 0.01
    roll = torch.where(wild, u(-1.3, 1.3, total), roll)
    pitch = torch.where(wild, u(-1.3, 1.3, total), pitch)
    yaw = u(-math.pi, math.pi, total)
    quat = euler_to_quat_wxyz(roll, pitch, yaw)
    joints = u(-0.3, 0.3, total, 13).float()
    actions = u(-1.0, 1.0, total, 2).float()
    prev_actions = (actions + u(-0.1, 0.1, total, 2)).float()
    # Ackermann branch rows: stand still, straight line, turn on the spot
    sp = torch.arange(total) % 97
    actions[sp == 0] = 0.0
    actions[sp == 1, 1] = 1e-6
    actions[sp == 2, 0] = 0.0
    alpha = u(0, 2 * math.pi, total)
    tr = torch.full((total,), 8.0)
    near = torch.rand(total, generator=g) < 0.005
    far = torch.rand(total, generator=g) < 0.005
    tr = torch.where(near, u(0.0, 0.3, total), tr)
    tr = torch.where(far, u(10.5, 12.0, total), tr)
    target = torch.zeros(total, 3)
    target[:, 0] = pos[:, 0] + tr * torch.cos(alpha)
    target[:, 1] = pos[:, 1] + tr * torch.sin(alpha)
    progress = torch.randint(0, 3001, (total,), generator=g)
    progress[torch.arange(total) % 53 == 0] = 2999
    progress[torch.arange(total) % 59 == 0] = 3000
    s = slice(env_offset, total)
    return dict(pos=pos[s].contiguous(), quat=quat[s].contiguous(), joints=joints[s].contiguous(),
                actions=actions[s].contiguous(), prev_actions=prev_actions[s].contiguous(),
                target=target[s].float().contiguous(), progress=progress[s].contiguous())


class SyntheticRoverView:
    """Stand-in for the ArticulationView the task talks to (rover.py:274-275,291,412-414): serves poses and
    joint positions from tensors and records the joint targets it is handed."""
    name = "rover_view"
    actuated_pos_indices = [6, 8, 4, 7]            # robots/articulations/views/rover_view.py:38-47
    actuated_vel_indices = [10, 5, 12, 9, 3, 11]

    def __init__(self, pos, quat, joints):
        self.pos, self.quat, self.joints = pos, quat, joints
        self.position_targets = None
        self.velocity_targets = None

    def get_world_poses(self):
        return self.pos, self.quat

    def get_joint_positions(self):
        return self.joints

    def set_joint_position_targets(self, positions, indices=None, joint_indices=None):
        self.position_targets = positions

    def set_joint_velocity_targets(self, velocities, indices=None, joint_indices=None):
        self.velocity_targets = velocities


def make_task(world, st, device="cuda:0", level=2, sem=0, num_envs_total=None, compact_terrain=False):
    """RoverTask mirror wired to a SyntheticRoverView holding the env state `st` (make_env_state)."""
    from .rover import RoverTask
    from .terrain_utils import stone_info_from_array
    dev = torch.device(device)
    view = SyntheticRoverView(st["pos"].to(dev), st["quat"].to(dev), st["joints"].to(dev))
    N = st["pos"].shape[0]
    task = RoverTask(view, N, (world.map_indices, world.triangles, world.vertices),
                     (world.rock_indices, world.rock_triangles, world.rock_vertices),
                     stone_info_from_array(world.stone_info.numpy(), device=dev), world.heightmap, device=device,
                     horizontal_scale=world.hm_res, sem=sem, num_envs_total=num_envs_total, compact_terrain=compact_terrain)
    task.curriculum_level = level
    task.target_positions = st["target"].to(dev).clone()
    task.progress_buf = st["progress"].to(dev).clone()
    task.initial_pos = st["pos"].to(dev).clone()
    task.linear_velocity.input_state(st["prev_actions"][:, 0].to(dev))
    task.angular_velocity.input_state(st["prev_actions"][:, 1].to(dev))
    return task
