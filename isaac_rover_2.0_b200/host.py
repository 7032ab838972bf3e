"""Host-buffer front end of the hot path: the call a simulator that keeps its state in HOST memory makes.

One `step()` = pinned host -> device copies of this step's poses / joints / actions, the device hot path
(RoverTask.hot_step), and device -> pinned host copies of obs_buf / rew_buf / reset_buf, all on the current
stream, with one synchronisation at the end.  bench.py's `e2e` number times exactly this call.
"""
import torch


class HostPipeline:
    def __init__(self, task):
        self.task = task
        N, dev = task.num_envs, torch.device(task._device)
        pin = dict(pin_memory=True)
        self.h_pos = torch.empty((N, 3), dtype=torch.float32, **pin)
        self.h_quat = torch.empty((N, 4), dtype=torch.float32, **pin)
        self.h_joints = torch.empty((N, 13), dtype=torch.float32, **pin)
        self.h_actions = torch.empty((N, 2), dtype=torch.float32, **pin)
        self.h_obs = torch.empty((N, task.num_observations), dtype=torch.float32, **pin)
        self.h_rew = torch.empty((N,), dtype=torch.float32, **pin)
        self.h_reset = torch.empty((N,), dtype=torch.int64, **pin)
        self.d_pos = torch.empty((N, 3), dtype=torch.float32, device=dev)
        self.d_quat = torch.empty((N, 4), dtype=torch.float32, device=dev)
        self.d_joints = torch.empty((N, 13), dtype=torch.float32, device=dev)
        self.d_actions = torch.empty((N, 2), dtype=torch.float32, device=dev)
        view = task._rover
        view.pos, view.quat, view.joints = self.d_pos, self.d_quat, self.d_joints
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in (self.h_pos, self.h_quat, self.h_joints, self.h_actions))
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in (self.h_obs, self.h_rew, self.h_reset))

    def step(self, pos, quat, joints, actions):
        """pos/quat/joints/actions: host tensors (any memory).  Returns pinned host (obs, rew, reset)."""
        for src, stage, dst in ((pos, self.h_pos, self.d_pos), (quat, self.h_quat, self.d_quat),
                                (joints, self.h_joints, self.d_joints), (actions, self.h_actions, self.d_actions)):
            if src.data_ptr() != stage.data_ptr():
                stage.copy_(src)
            dst.copy_(stage, non_blocking=True)
        t = self.task
        t.hot_step(self.d_actions)
        self.h_obs.copy_(t.obs_buf, non_blocking=True)
        self.h_rew.copy_(t.rew_buf, non_blocking=True)
        self.h_reset.copy_(t.reset_buf, non_blocking=True)
        torch.cuda.current_stream(t.obs_buf.device).synchronize()
        return self.h_obs, self.h_rew, self.h_reset
