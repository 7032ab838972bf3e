"""`ray_distance` -- mirror of tasks/utils/camera/ray_casting.py:3-66 (n rays vs n triangles, fp16)."""
import torch

from . import _lib


def ray_distance(sources, directions, triangles, device='cuda:0', dtype=torch.float16):
    _lib.require_cuda(sources, directions, triangles)
    if dtype != torch.float16:
        raise RuntimeError("rover_b200.ray_distance computes in float16 only (the reference's dtype)")
    lib = _lib.load()
    s = sources.to(torch.float16).contiguous()
    d = directions.to(torch.float16).contiguous()
    t = triangles.to(torch.float16).contiguous()
    n = s.shape[0]
    if d.shape != (n, 3) or t.shape != (n, 3, 3) or s.shape != (n, 3):
        raise ValueError("ray_distance: expected sources [n,3], directions [n,3], triangles [n,3,3]")
    k = torch.empty((n,), dtype=torch.float16, device=s.device)
    pt = torch.empty((n, 3), dtype=torch.float16, device=s.device)
    with torch.cuda.device(s.device):
        _lib.check(lib.rvb_ray_distance(_lib.ptr(s), _lib.ptr(d), _lib.ptr(t), n, _lib.ptr(k), _lib.ptr(pt), _lib.stream_of(s)))
    return k, pt
