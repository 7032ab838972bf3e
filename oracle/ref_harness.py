"""TEST / BASELINE INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference classes on a synthetic world.

Used by tests/ (pinning the oracle, generating golden vectors) and by bench.py's `cpu_baseline` / `--impl reference` /
`torch_eager_gpu_baseline` legs: the reference's own torch code (Camera.get_depths, Rock_Detection.get_collisions,
RoverTask.get_observations / calculate_metrics / is_done, Ackermann), imported through oracle/ref_import.py from
`/root/reference` (this container) or from the git-ignored copy `baseline/_ref/` that __graft_entry__.build() makes
(the copy travels to the GPU box; /root/reference does not).  Never imported by the product.
"""
import contextlib
import os
import tempfile
import types

import torch

import ref_import

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    """Where the reference's python tree is: $ROVER_REFERENCE_ROOT, else /root/reference, else baseline/_ref; None if absent."""
    for r in (os.environ.get("ROVER_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if r and os.path.isdir(os.path.join(r, "omniisaacgymenvs", "tasks")):
            return r
    return None


def load(device="cpu"):
    """ref_import.load() against whichever reference tree is present."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (neither /root/reference nor baseline/_ref)")
    ref_import.REF_ROOT = root
    return ref_import.load(device)


@contextlib.contextmanager
def reference_assets(world, directory=None, device="cpu"):
    """Write the world in the reference's asset layout (camera.py:154-161, rock_detect.py:151-158) and chdir there.  The
    reference never moves the loaded tensors: its asset files were saved from tensors on its device, so they are written from
    `device` here (torch.load restores them there)."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory(dir=directory) as tmp:
        for sub, idx, tri, ver in (("knn_terrain", world.map_indices, world.triangles, world.vertices),
                                   ("knn_rocks", world.rock_indices, world.rock_triangles, world.rock_vertices)):
            d = os.path.join(tmp, "tasks/utils/terrain", sub)
            os.makedirs(d)
            torch.save(idx.to(device), os.path.join(d, "map_indices.pt"))
            torch.save(tri.to(device), os.path.join(d, "triangles.pt"))
            torch.save(ver.to(device), os.path.join(d, "vertices.pt"))
        os.chdir(tmp)
        try:
            yield tmp
        finally:
            os.chdir(old)


def make_fake_task(ns, world, st, level=2, device="cpu"):
    """A SimpleNamespace standing in for `self` of RoverTask; `st` tensors must live on `device`."""
    N = st["pos"].shape[0]
    shift = torch.tensor([0, 0, 0.0], device=device)
    with reference_assets(world, device=device):
        cam = ns.Camera(device, shift)
        rock = ns.Rock_Detection(device, shift)
    rover = types.SimpleNamespace(name="rover_view", count=N,
                                  get_world_poses=lambda: (st["pos"], st["quat"]),
                                  get_joint_positions=lambda: st["joints"])
    lin = ns.Memory(N, 1, 3, device)
    ang = ns.Memory(N, 1, 3, device)
    lin.input_state(st["prev_actions"][:, 0]); ang.input_state(st["prev_actions"][:, 1])
    lin.input_state(st["actions"][:, 0]); ang.input_state(st["actions"][:, 1])
    fake = types.SimpleNamespace(
        _rover=rover, _device=device, num_envs=N, _num_envs=N, Camera=cam, Rock_detector=rock,
        target_positions=st["target"].clone(), curriculum_level=level, save_teacher_data=False,
        obs_buf=torch.zeros((N, 4 + 634 + 1112), device=device), rew_buf=torch.zeros(N, device=device),
        reset_buf=torch.ones(N, dtype=torch.long, device=device),
        progress_buf=st["progress"].clone(), extras={}, _num_proprioceptive=4,
        linear_velocity=lin, angular_velocity=ang, is_evaluation=False, max_episode_length=3000,
        rew_scales=dict(pos_reward=1.0, terminalReward=0, collision_reward=0.3, heading_contraint_reward=0.05,
                        motion_contraint_reward=-0.01, goal_angle_reward=0.3, boogie_contraint_reward=0.5),
        rover_rot=ns.tensor_quat_to_eul(st["quat"]))
    fake.check_collision = lambda w, b: ns.RoverTask.check_collision(fake, w, b)
    return fake


def reference_step(ns, fake, st):
    """One env-step of the hot path through the reference's own code, PhysX excluded: the action half of pre_physics_step
    (rover.py:343,379-409: histories, Ackermann, joint-target permutation) and post_physics_step (rl_task.py:239-259)."""
    dev = fake._device
    RT = ns.RoverTask
    fake.rover_rot = ns.tensor_quat_to_eul(st["quat"])                                  # rover.py:343
    act = st["actions"]
    fake.linear_velocity.input_state(act[:, 0])                                        # rover.py:379-380
    fake.angular_velocity.input_state(act[:, 1])
    steer, vel = ns.Ackermann(act[:, 0], act[:, 1], dev) if dev == "cpu" else ns.Ackermann(act[:, 0], act[:, 1])
    positions = torch.stack((steer[:, 1], steer[:, 5], steer[:, 0], steer[:, 4]), 1)   # rover.py:400-403
    velocities = torch.stack((vel[:, 1], vel[:, 3], vel[:, 5], vel[:, 0], vel[:, 2], vel[:, 4]), 1)
    fake.progress_buf[:] += 1                                                          # rl_task.py:250
    RT.get_observations(fake)
    RT.calculate_metrics(fake)
    RT.is_done(fake)
    return fake.obs_buf, fake.rew_buf, fake.reset_buf, positions, velocities
