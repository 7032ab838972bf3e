"""Rover-frame heightmap point pattern -- mirror of the reference's `Heightmap`
(tasks/utils/camera/heightmap_distribution.py:11-204): same constructor, same getters, same
1634-point pattern (634 sparse + 1112 dense, 112 shared).

The pattern is an init-time constant, generated on the host with vectorised numpy: the reference's
accumulating `x += delta` loops are reproduced with cumulative sums (sequential fp64 adds), its border
predicates with element-wise half-plane tests, its exact-value de-duplication with a set of fp64 pairs.
The per-step getters are plain column gathers on the device tensor handed in.
"""
import numpy as np
import torch

_COARSE = (((1.220, 0.118), (4.4455, 3.150), 'over'), ((-1.220, 0.118), (-4.4455, 3.150), 'over'),
           ((1.220, 0.118), (-1.220, 0.118), 'over'))                                   # :16
_FINE = (((1.0, 0.118), (1.0, 0.119), 'left'), ((-1.0, 0.118), (-1.0, 0.119), 'right'),
         ((1.0, 0.118), (-1.0, 0.118), 'over'), ((1.0, 1.400), (-1.0, 1.400), 'below'))  # :19


def _axis(delta, first_is_start):
    """Values visited by `v = -10; while v < 10: ...; v += delta` (rows) or `v += delta; ...` (columns)."""
    seq = np.cumsum(np.concatenate(([-10.0], np.full(400, delta))))
    seq = seq[: int(np.argmax(seq >= 10)) + 1]          # up to and including the first value >= 10
    return seq[:-1] if first_is_start else seq[1:]


def _inside(x, y, lines):
    """Vectorised `_inside_borders` (:153-193), including its treatment of 'left' on slanted lines."""
    ok = np.ones(x.shape, dtype=bool)
    for p0, p1, side in lines:
        d = np.subtract(p0, p1)
        if d[0] == 0:
            if side == 'right':
                ok &= ~(x < p0[0])
            if side == 'left':
                ok &= ~(x > p0[0])
            continue
        a = d[1] / d[0]
        b = p0[1] - a * p0[0]
        if a == 0:
            if side == 'below':
                ok &= ~(y > b)
            if side == 'over':
                ok &= ~(y < b)
            continue
        if side == 'over':
            ok &= ~(y < a * x + b)
        if side == 'below':
            ok &= ~(y > a * x + b)
        if side in ('right', 'left'):
            ok &= ~(x < (y - b) / a)
    return ok


def build_pattern(delta_coarse=0.15, delta_fine=0.05, coarse_radius=3.5, z_offset=-0.26878,
                  coarse_border=_COARSE, fine_border=_FINE):
    """-> (distribution f64 [P,3] (x fwd, y left), coarse_idx i64, fine_idx i64) as numpy arrays."""
    ys, xs = _axis(delta_coarse, True), _axis(delta_coarse, False)
    Y, X = np.meshgrid(ys, xs, indexing="ij")            # rows = y (outer loop), cols = x (inner loop)
    m = _inside(X, Y, coarse_border) & (np.sqrt(X * X + Y * Y) < coarse_radius)
    cx, cy = X[m], Y[m]
    n_coarse = cx.size
    ys, xs = _axis(delta_fine, True), _axis(delta_fine, False)
    Y, X = np.meshgrid(ys, xs, indexing="ij")
    m = _inside(X, Y, fine_border)
    fx, fy = X[m], Y[m]
    seen = set(zip(cx.tolist(), cy.tolist()))
    new = np.array([(a, b) not in seen for a, b in zip(fx.tolist(), fy.tolist())], dtype=bool)
    px = np.concatenate((cx, fx[new]))
    py = np.concatenate((cy, fy[new]))
    fine_idx = np.nonzero(_inside(px, py, fine_border))[0]
    pts = np.round(np.stack((px, py, np.full(px.shape, z_offset)), 1), 4)
    return pts[:, [1, 0, 2]].copy(), np.arange(n_coarse, dtype=np.int64), fine_idx.astype(np.int64)


class Heightmap():
    def __init__(self, device='cuda:0'):
        self.device = device
        self.delta_coarse, self.delta_fine = 0.15, 0.05
        self.coarse_radius, self.fine_radius = 3.5, 1.2
        self.z_offset = -0.26878
        pts, ci, fi = build_pattern()
        self.distribution = torch.from_numpy(pts).to(device)
        self.coarse_idx = torch.from_numpy(ci).to(device)
        self.fine_idx = torch.from_numpy(fi).to(device)
        self.beneath_idx = torch.tensor([], device=device)          # see_beneath is False (:27)

    def get_distribution(self):
        return self.distribution

    def get_sparse_vector(self, rays):
        return rays[:, self.coarse_idx]

    def get_dense_vector(self, rays):
        return rays[:, self.fine_idx]

    def get_beneath_vector(self, rays):
        return rays[:, self.beneath_idx]

    def get_num_sparse_vector(self):
        return self.coarse_idx.shape[0]

    def get_num_dense_vector(self):
        return self.fine_idx.shape[0]

    def get_num_beneath_vector(self):
        return self.beneath_idx.shape[0]

    def obs_columns(self, col0=4):
        """int32 [P] x2: for ray p the obs column fed by the sparse vector / by the dense vector (-1 = none).
        Used by the fused epilogue of the ray-cast (rover.py:324-325)."""
        P = self.distribution.shape[0]
        a = torch.full((P,), -1, dtype=torch.int32)
        b = torch.full((P,), -1, dtype=torch.int32)
        ci, fi = self.coarse_idx.cpu(), self.fine_idx.cpu()
        a[ci] = torch.arange(ci.numel(), dtype=torch.int32) + col0
        b[fi] = torch.arange(fi.numel(), dtype=torch.int32) + col0 + ci.numel()
        return a.to(self.device), b.to(self.device)
