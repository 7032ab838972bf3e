"""Observation hooks and the teacher-data recorder: the per-step epilogues of
RoverTask.get_observations the reference's authors toggled by hand (rover.py:298-317 recorder, :326-329 noise / dropout /
offset / remove_idx masking).  One kernel each behind the C ABI (csrc/hooks.cu); no CPU path."""
import math

import torch

from . import _lib


class ObsHooks:
    """The four commented-out lines of rover.py:326-329 as switches.  Reference values: noise_std = 0.20 ** 0.5,
    dropout_p = 0.1, offset = 0.02, remove_idx = torch.load('remove_idx.pt') (rover.py:88; indices into obs[:, 4:])."""

    def __init__(self, num_observations, noise_std=0.0, dropout_p=0.0, offset=0.0, remove_idx=None, num_proprioceptive=4, seed=42,
                 device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("rover_b200: ObsHooks needs a CUDA device (there is no CPU path)")
        if not (0.0 <= dropout_p < 1.0) or noise_std < 0:
            raise ValueError("dropout_p must be in [0, 1) and noise_std >= 0")
        self.num_observations = int(num_observations)
        self.col0 = int(num_proprioceptive)
        self.noise_std, self.dropout_p, self.offset, self.seed = float(noise_std), float(dropout_p), float(offset), int(seed)
        self.mask = None
        if remove_idx is not None:
            idx = torch.as_tensor(remove_idx, dtype=torch.long).flatten() + self.col0          # rover.py:329
            if idx.numel() and (idx.min() < 0 or idx.max() >= self.num_observations):
                raise IndexError("remove_idx out of range")
            m = torch.zeros(self.num_observations, dtype=torch.uint8)
            m[idx] = 1
            self.mask = m.to(self.device)
        self._lib = _lib.load()

    @classmethod
    def reference_values(cls, num_observations, remove_idx=None, **kw):
        return cls(num_observations, noise_std=math.sqrt(0.20), dropout_p=0.1, offset=0.02, remove_idx=remove_idx, **kw)

    def apply(self, obs, epoch, env_offset=0):
        """In place on obs f32 [N, >= num_observations]; `epoch` = the step counter (a new draw per step)."""
        _lib.require_cuda(obs)
        if obs.dtype != torch.float32 or obs.dim() != 2 or obs.stride(1) != 1 or obs.shape[1] < self.num_observations:
            raise RuntimeError("obs must be float32 [N, >=%d] with unit column stride" % self.num_observations)
        with torch.cuda.device(obs.device):
            _lib.check(self._lib.rvb_obs_hooks(_lib.ptr(obs), obs.stride(0), obs.shape[0], self.num_observations, self.col0,
                                               self.noise_std, self.dropout_p, self.offset, _lib.ptr(self.mask), self.seed,
                                               int(epoch), int(env_offset), _lib.stream_of(obs)))
        return obs


class TeacherRecorder:
    """rover.py:174-180,298-317: a [T, N, 1 + num_actions + num_observations] block filled one step per call and written to
    `teacher_dataset_<nr>.pt` in the reference's dict layout when full.  The block lives on the device (the reference keeps
    it in host memory and pays a device-to-host copy of obs_buf per step)."""

    def __init__(self, num_envs, num_observations, num_sparse, num_dense, num_proprioceptive=4, steps=5 * 30, device="cuda:0",
                 directory=".", save=True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("rover_b200: TeacherRecorder needs a CUDA device (there is no CPU path)")
        self.num_envs, self.num_observations = int(num_envs), int(num_observations)
        self.info = {"reset": 1, "actions": 2, "proprioceptive": num_proprioceptive, "sparse": num_sparse, "dense": num_dense}
        self.teacher_dataset = torch.empty((steps, num_envs, 3 + num_observations), device=self.device)     # rover.py:178
        self.curr_timestep = 0
        self.dataset_nr = 0
        self.directory, self.save = directory, save
        self.last_file = None
        self._lib = _lib.load()

    def record(self, reset_info, actions, obs):
        """One step: row = [reset_info, actions, obs] per env (rover.py:299-300,364,374-375).  Returns the path of the file
        written when this call filled the block, else None."""
        _lib.require_cuda(reset_info, actions, obs)
        reset_info = reset_info.to(torch.float32).contiguous()
        actions = actions.to(torch.float32)
        if actions.stride(1) != 1:
            actions = actions.contiguous()
        if obs.dtype != torch.float32 or obs.stride(1) != 1 or obs.shape != (self.num_envs, obs.shape[1]) or obs.shape[1] < self.num_observations:
            raise RuntimeError("obs must be float32 [%d, >=%d]" % (self.num_envs, self.num_observations))
        row = self.teacher_dataset[self.curr_timestep]
        with torch.cuda.device(obs.device):
            _lib.check(self._lib.rvb_teacher_record(_lib.ptr(reset_info), _lib.ptr(actions), actions.stride(0), _lib.ptr(obs),
                                                    obs.stride(0), self.num_envs, self.num_observations, _lib.ptr(row), row.stride(0),
                                                    _lib.stream_of(obs)))
        self.curr_timestep += 1
        if self.curr_timestep < self.teacher_dataset.shape[0]:
            return None
        self.curr_timestep = 0                                                    # rover.py:302-317
        path = None
        if self.save:
            import os
            path = os.path.join(self.directory, "teacher_dataset_" + str(self.dataset_nr) + ".pt")
            torch.save({"info": dict(self.info), "data": self.teacher_dataset.cpu()}, path)
        self.last_file = path
        self.dataset_nr += 1
        self.teacher_dataset = torch.empty_like(self.teacher_dataset)
        return path
