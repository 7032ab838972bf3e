"""`Ackermann` -- mirror of tasks/utils/kinematics.py:14-67 (6-wheel Ackermann, turn-on-the-spot mode)."""
import torch

from . import _lib


def Ackermann(lin_vel, ang_vel, device='cuda:0', sem=_lib.SEM_TORCH_CUDA, want_targets=False):
    """-> (steering_angles [N,6], motor_velocities [N,6]) in wheel order FL,FR,ML,MR,RL,RR.
    lin_vel / ang_vel may be strided 1-D views (the reference passes actions[:,0], actions[:,1], rover.py:391).
    want_targets=True additionally returns the joint targets of rover.py:400-409."""
    _lib.require_cuda(lin_vel, ang_vel)
    lib = _lib.load()
    if lin_vel.dtype != torch.float32 or ang_vel.dtype != torch.float32:
        lin_vel, ang_vel = lin_vel.float(), ang_vel.float()
    if lin_vel.dim() != 1 or ang_vel.shape != lin_vel.shape:
        raise ValueError("Ackermann: lin_vel and ang_vel must be 1-D and of equal length")
    N, dev = lin_vel.shape[0], lin_vel.device
    steer = torch.empty((N, 6), dtype=torch.float32, device=dev)
    vel = torch.empty((N, 6), dtype=torch.float32, device=dev)
    pos_t = torch.empty((N, 4), dtype=torch.float32, device=dev) if want_targets else None
    vel_t = torch.empty((N, 6), dtype=torch.float32, device=dev) if want_targets else None
    ls = lin_vel.stride(0) if N > 1 else 1
    as_ = ang_vel.stride(0) if N > 1 else 1
    with torch.cuda.device(dev):
        _lib.check(lib.rvb_ackermann(_lib.ptr(lin_vel), ls, _lib.ptr(ang_vel), as_, N, _lib.ptr(steer), _lib.ptr(vel),
                                     _lib.ptr(pos_t), _lib.ptr(vel_t), sem, _lib.stream_of(lin_vel)))
    if want_targets:
        return steer, vel, pos_t, vel_t
    return steer, vel
