"""Host-buffer front end of the hot path: the call a simulator that keeps its state in HOST memory makes.

`submit()` enqueues one env step: pinned host -> device copies of this step's poses / joints / actions, the device hot path
(RoverTask.hot_step, one library call), and -- on a second stream -- device -> pinned host copies of obs_buf / rew_buf /
reset_buf.  `result(slot)` waits for that step's copies and returns the host tensors.  Two slots (double buffering): the
28.7 MB observation read-back of step i overlaps the kernels of step i+1, and the inputs of step i+1 are uploaded on a third
stream into the other slot's device buffers while step i still computes (measured: small uploads queued on the compute stream
start together with the previous step's read-back and wait behind it, +0.2 ms per step).  `step()` = submit + result (no overlap).
bench.py's `e2e` number times submit/result in a loop where every step's inputs are copied in and every step's results are
read on the host.
"""
import torch


class HostPipeline:
    def __init__(self, task, depth=2):
        self.task = task
        self.depth = depth
        N, dev = task.num_envs, torch.device(task._device)
        self.dev = dev
        pin = dict(pin_memory=True)
        mk = lambda shape, dt: [torch.empty(shape, dtype=dt, **pin) for _ in range(depth)]      # noqa: E731
        self.h_pos, self.h_quat = mk((N, 3), torch.float32), mk((N, 4), torch.float32)
        self.h_joints, self.h_actions = mk((N, 13), torch.float32), mk((N, 2), torch.float32)
        self.h_obs, self.h_rew, self.h_reset = mk((N, task.num_observations), torch.float32), mk((N,), torch.float32), mk((N,), torch.int64)
        dk = lambda shape: [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(depth)]      # noqa: E731
        self.d_pos, self.d_quat, self.d_joints, self.d_actions = dk((N, 3)), dk((N, 4)), dk((N, 13)), dk((N, 2))
        self.d_obs = [torch.zeros((N, task.num_observations), dtype=torch.float32, device=dev) for _ in range(depth)]
        self.d_rew = [torch.zeros((N,), dtype=torch.float32, device=dev) for _ in range(depth)]
        self.d_reset = [torch.zeros((N,), dtype=torch.int64, device=dev) for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(dev)
        self.up_stream = torch.cuda.Stream(dev)
        self.uploaded = [torch.cuda.Event() for _ in range(depth)]
        self.computed = [torch.cuda.Event() for _ in range(depth)]
        self.done = [None] * depth
        self.i = 0
        view = task._rover
        view.pos, view.quat, view.joints = self.d_pos[0], self.d_quat[0], self.d_joints[0]
        self.h2d_bytes = sum(t[0].numel() * t[0].element_size() for t in (self.h_pos, self.h_quat, self.h_joints, self.h_actions))
        self.d2h_bytes = sum(t[0].numel() * t[0].element_size() for t in (self.h_obs, self.h_rew, self.h_reset))

    def submit(self, pos, quat, joints, actions):
        """pos/quat/joints/actions: host tensors (any memory).  Enqueues the step and returns its slot."""
        k = self.i % self.depth
        self.i += 1
        if self.done[k] is not None:
            self.done[k].synchronize()          # slot k's previous read-back (and so its staging copies) has finished
        cur = torch.cuda.current_stream(self.dev)
        # slot k's device inputs were last read by the step whose read-back has just been waited for
        with torch.cuda.stream(self.up_stream):
            for src, stage, dst in ((pos, self.h_pos[k], self.d_pos[k]), (quat, self.h_quat[k], self.d_quat[k]),
                                    (joints, self.h_joints[k], self.d_joints[k]), (actions, self.h_actions[k], self.d_actions[k])):
                if src.data_ptr() != stage.data_ptr():
                    stage.copy_(src)
                dst.copy_(stage, non_blocking=True)
            self.uploaded[k].record(self.up_stream)
        cur.wait_event(self.uploaded[k])
        t = self.task
        view = t._rover
        view.pos, view.quat, view.joints = self.d_pos[k], self.d_quat[k], self.d_joints[k]
        t.obs_buf, t.rew_buf, t.reset_buf = self.d_obs[k], self.d_rew[k], self.d_reset[k]
        t.hot_step(self.d_actions[k])
        self.computed[k].record(cur)
        self.copy_stream.wait_event(self.computed[k])
        with torch.cuda.stream(self.copy_stream):
            self.h_obs[k].copy_(self.d_obs[k], non_blocking=True)
            self.h_rew[k].copy_(self.d_rew[k], non_blocking=True)
            self.h_reset[k].copy_(self.d_reset[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.done[k] = ev
        # the next step's kernels write d_obs[k'] of ANOTHER slot; slot k is rewritten only after done[k] (see above)
        return k

    def result(self, k):
        self.done[k].synchronize()
        return self.h_obs[k], self.h_rew[k], self.h_reset[k]

    def step(self, pos, quat, joints, actions):
        """Synchronous convenience: returns pinned host (obs, rew, reset) of this step."""
        return self.result(self.submit(pos, quat, joints, actions))
