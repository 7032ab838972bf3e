"""Generates tests/golden/rover_golden.pt by running the UNMODIFIED reference (imported from
/root/reference with the stub recipe of oracle/ref_import.py) on a small seeded synthetic world.

Run in the CPU container only:   python tests/golden/make_golden.py
Everything stored under "ref_*" is an output of the reference's own code (torch CPU); "in_*" are inputs;
"trig" holds the fp32 sin/cos of the negated euler angles as torch-CPU computed them, so a GPU test can
inject them and demand bit-exact results for everything that is not a libm call.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import isaac_rover_b200  # noqa: E402
from isaac_rover_b200 import synth  # noqa: E402
import ref_import  # noqa: E402
import ref_world  # noqa: E402
import rover_oracle as O  # noqa: E402


def main():
    torch.manual_seed(42)
    ns = ref_import.load("cpu")
    RT = ns.RoverTask
    world = synth.make_world(length=8.0, nv=30, K=48, n_stones=5, seed=7)
    N = 8
    st = synth.make_env_state(world, N, seed=11, margin=2.5)
    st["target"][0, 0:2] = st["pos"][0, 0:2] + 0.1          # goal reached
    st["target"][1, 0:2] = st["pos"][1, 0:2] + 9.0          # too far
    fake = ref_world.make_fake_task(ns, world, st, level=2)
    g = {"note": "reference outputs, torch %s CPU" % torch.__version__}
    g["world"] = dict(length=world.length, res=world.res, G=world.G, K=world.K, hm_res=world.hm_res,
                      vertices=world.vertices, triangles=world.triangles,
                      map_indices=world.map_indices.to(torch.int16),
                      rock_vertices=world.rock_vertices, rock_triangles=world.rock_triangles,
                      rock_indices=world.rock_indices.to(torch.int16),
                      stone_info6=world.stone_info, heightmap=world.heightmap)
    for k, v in st.items():
        g["in_" + k] = v
    eul = ns.tensor_quat_to_eul(st["quat"])
    g["ref_euler"] = eul
    g["trig"] = torch.stack([f(-eul[:, a]) for a in range(3) for f in (torch.sin, torch.cos)], 1)
    # heightmap pattern + Camera.get_depths
    hm = fake.Camera.heightmap
    g["ref_pattern"], g["ref_coarse_idx"], g["ref_fine_idx"] = hm.distribution, hm.coarse_idx, hm.fine_idx
    dist, pt, src = fake.Camera.get_depths(st["pos"], eul)
    g["ref_dist"], g["ref_pt"], g["ref_sources"] = dist, pt, src
    o = O.get_depths(st["pos"], eul, hm.distribution, world.map_indices, world.triangles, world.vertices,
                     torch.tensor([0, 0, 0.0]))
    assert torch.equal(o["dist"], dist) and torch.equal(o["pt"], pt) and torch.equal(o["sources"], src)
    g["oracle_slot"], g["oracle_tri"] = o["slot"].to(torch.int32), o["tri"].to(torch.int32)   # argmin (reference keeps it internal, camera.py:116)
    # Rock_Detection.get_collisions
    wheel, body = fake.Rock_detector.get_collisions(st["pos"], eul, st["joints"])
    g["ref_wheel"], g["ref_body"] = wheel, body
    ws, wd = fake.Rock_detector._get_wheel_rays(st["pos"], eul, st["joints"])
    bs, bd = fake.Rock_detector._get_body_rays(st["pos"], eul)
    g["ref_rock_sources"], g["ref_rock_dirs"] = torch.cat((ws, bs), 1), torch.cat((wd, bd), 1)
    # get_observations / calculate_metrics / is_done through the unbound RoverTask methods
    RT.get_observations(fake)
    g["ref_obs"], g["ref_heading"], g["ref_rock_collision"] = fake.obs_buf.clone(), fake.heading_diff.clone(), fake.rock_collison.clone()
    RT.calculate_metrics(fake)
    g["ref_rew"] = fake.rew_buf.clone()
    g["ref_extras"] = {k: v.clone() for k, v in fake.extras.items()}
    RT.is_done(fake)
    g["ref_reset"] = fake.reset_buf.clone()
    fake.curriculum_level = 1
    RT.calculate_metrics(fake)
    RT.is_done(fake)
    g["ref_rew_level1"], g["ref_reset_level1"] = fake.rew_buf.clone(), fake.reset_buf.clone()
    # Ackermann + joint target mapping
    steer, vel = ns.Ackermann(st["actions"][:, 0], st["actions"][:, 1], "cpu")
    g["ref_steer"], g["ref_vel"] = steer, vel
    ka_lin = torch.tensor([0.5, 0.0, 1.0, 0.0, -0.7, 0.3, 0.9])
    ka_ang = torch.tensor([0.2, -2.0, 1e-6, 0.0, 0.4, -3.0, 0.0])
    g["in_ka_lin"], g["in_ka_ang"] = ka_lin, ka_ang
    g["ref_ka_steer"], g["ref_ka_vel"] = ns.Ackermann(ka_lin, ka_ang, "cpu")
    # ray_distance known answers (unit triangle) + random pairs
    tri = torch.tensor([[[1, 0, 0], [0, 1, 0], [0, 0, 0]]], dtype=torch.float16).repeat(6, 1, 1)
    tri[3] = 0                                           # degenerate
    srcs = torch.tensor([[0.25, 0.25, 1], [0.25, 0.25, -1], [2, 2, 1], [0.2, 0.2, 1], [0.25, 0.25, 1], [-0.05, 0.3, 0.5]], dtype=torch.float16)
    dirs = torch.tensor([[0, 0, -1], [0, 0, -1], [0, 0, -1], [0, 0, -1], [1, 0, 0], [0, 0, -2]], dtype=torch.float16)
    g["in_rd_src"], g["in_rd_dir"], g["in_rd_tri"] = srcs, dirs, tri
    g["ref_rd_k"], g["ref_rd_pt"] = ns.ray_distance(srcs, dirs, tri)
    gen = torch.Generator().manual_seed(5)
    rs = (torch.rand(4096, 3, generator=gen) * 4 - 2).to(torch.float16)
    rd = (torch.rand(4096, 3, generator=gen) * 2 - 1).to(torch.float16)
    rt = (torch.rand(4096, 3, 3, generator=gen) * 4 - 2).to(torch.float16)
    rt[::7, :, 2] = rt[::7, 0:1, 2]                       # horizontal triangles
    rd[::5, 0:2] = 0
    g["in_rr_src"], g["in_rr_dir"], g["in_rr_tri"] = rs, rd, rt
    g["ref_rr_k"], g["ref_rr_pt"] = ns.ray_distance(rs, rd, rt)
    # stones: goal validity, spawn validation, height lookup
    fake.stone_info = O.read_stone_info(world.stone_info.numpy())
    g["ref_stone7"] = fake.stone_info
    fake.initial_pos = st["pos"].clone()
    ids = torch.arange(N)
    dr = torch.cdist(fake.target_positions[ids][:, 0:2], fake.stone_info[:, 0:2], p=2.0)
    g["ref_goal_nearest"] = torch.min(dr - fake.stone_info[:, 6], dim=1)[0]
    e2, cnt = RT.check_goal_collision(fake, ids)
    g["ref_goal_env_ids"], g["ref_goal_count"] = e2, cnt
    many = torch.rand(64, 3, generator=gen) * world.length
    g["in_spawn_pos"] = many.clone()
    g["ref_spawn_pos"] = RT.avoid_pos_rock_collision(fake, many.clone())
    g["ref_spawn_height"] = RT.get_pos_height(fake, world.heightmap, g["ref_spawn_pos"][:, 0:2], world.hm_res, 1, torch.tensor([0.0, 0.0]))
    dr = torch.cdist(many[:, 0:2], fake.stone_info[:, 0:2], p=2.0)
    g["ref_many_nearest"] = torch.min(dr - fake.stone_info[:, 6], dim=1)[0]
    out = os.path.join(HERE, "rover_golden.pt")
    torch.save(g, out)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")
    print("hit fraction", (dist != 11).float().mean().item(), "wheel hits", (wheel != 11).float().mean().item(),
          "collisions", fake.rock_collison.tolist(), "resets", g["ref_reset"].tolist(), "goal invalid", cnt)


if __name__ == "__main__":
    main()
