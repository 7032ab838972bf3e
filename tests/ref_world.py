"""Test helper: run the UNMODIFIED reference classes on a synthetic world (CPU container only)."""
import contextlib
import os
import tempfile
import types

import torch

import ref_import


@contextlib.contextmanager
def reference_assets(world):
    """Write the world in the reference's asset layout (camera.py:154-161, rock_detect.py:151-158) and chdir there."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        for sub, idx, tri, ver in (("knn_terrain", world.map_indices, world.triangles, world.vertices),
                                   ("knn_rocks", world.rock_indices, world.rock_triangles, world.rock_vertices)):
            d = os.path.join(tmp, "tasks/utils/terrain", sub)
            os.makedirs(d)
            torch.save(idx, os.path.join(d, "map_indices.pt"))
            torch.save(tri, os.path.join(d, "triangles.pt"))
            torch.save(ver, os.path.join(d, "vertices.pt"))
        os.chdir(tmp)
        try:
            yield tmp
        finally:
            os.chdir(old)


def make_fake_task(ns, world, st, level=2):
    """A SimpleNamespace standing in for `self` of RoverTask."""
    N = st["pos"].shape[0]
    shift = torch.tensor([0, 0, 0.0])
    with reference_assets(world):
        cam = ns.Camera("cpu", shift)
        rock = ns.Rock_Detection("cpu", shift)
    rover = types.SimpleNamespace(name="rover_view",
                                  get_world_poses=lambda: (st["pos"], st["quat"]),
                                  get_joint_positions=lambda: st["joints"])
    lin = ns.Memory(N, 1, 3, "cpu")
    ang = ns.Memory(N, 1, 3, "cpu")
    lin.input_state(st["prev_actions"][:, 0]); ang.input_state(st["prev_actions"][:, 1])
    lin.input_state(st["actions"][:, 0]); ang.input_state(st["actions"][:, 1])
    fake = types.SimpleNamespace(
        _rover=rover, _device="cpu", num_envs=N, _num_envs=N, Camera=cam, Rock_detector=rock,
        target_positions=st["target"].clone(), curriculum_level=level, save_teacher_data=False,
        obs_buf=torch.zeros((N, 4 + 634 + 1112)), rew_buf=torch.zeros(N), reset_buf=torch.ones(N, dtype=torch.long),
        progress_buf=st["progress"].clone(), extras={}, _num_proprioceptive=4,
        linear_velocity=lin, angular_velocity=ang, is_evaluation=False, max_episode_length=3000,
        rew_scales=dict(pos_reward=1.0, terminalReward=0, collision_reward=0.3, heading_contraint_reward=0.05,
                        motion_contraint_reward=-0.01, goal_angle_reward=0.3, boogie_contraint_reward=0.5),
        rover_rot=ns.tensor_quat_to_eul(st["quat"]))
    fake.check_collision = lambda w, b: ns.RoverTask.check_collision(fake, w, b)
    return fake
