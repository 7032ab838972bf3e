"""`tensor_quat_to_eul` -- mirror of tasks/utils/math/tensor_quat_to_euler.py:6-31 (wxyz -> roll,pitch,yaw)."""
import torch

from . import _lib


def tensor_quat_to_eul(quats):
    _lib.require_cuda(quats)
    lib = _lib.load()
    q = quats.to(torch.float32).contiguous()
    if q.dim() != 2 or q.shape[1] != 4:
        raise ValueError("tensor_quat_to_eul: expected [N,4] (w,x,y,z)")
    e = torch.empty((q.shape[0], 3), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(lib.rvb_quat_to_euler(_lib.ptr(q), q.shape[0], _lib.ptr(e), _lib.stream_of(q)))
    return e
