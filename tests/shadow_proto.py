"""CPU proof-of-bound for the ray-cast "shadow" culling (csrc/raycast_shadow.cu).

For every (env, triangle) the kernel computes, in fp32, a conservative xy box of the ray sources that can pass the
packed-fp16 pre-filter of ray_casting.py:34-59 (which itself is a superset of the literal test).  This script
re-states that box computation in numpy fp32 (same formulas, same constants) and checks it against a brute-force
fp16 evaluation of the pre-filter on synthetic worlds: no passing (ray, triangle) pair may lie outside the box.
It also reports how tight the box is (rays per box vs. passing rays).

Run:  python tests/shadow_proto.py [--envs 12] [--length 200]
"""
import argparse
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import isaac_rover_b200 as R          # noqa: E402
import rover_oracle as O              # noqa: E402

F16 = np.float16
F32 = np.float32

# constants shared with the kernel
GAMMA = F32(2.0 ** -8)        # >= (1 + 2^-11)^6 - 1 with 33 % slack: six fp16 roundings per term
ALPHA = F32(2.0 ** -19)       # absolute (subnormal) rounding errors of one numerator
C1 = F16(np.frombuffer(np.uint16(0x2E68).tobytes(), dtype=F16)[0])
C2 = F16(np.frombuffer(np.uint16(0x3C6B).tobytes(), dtype=F16)[0])
TINY = F16(np.frombuffer(np.uint16(0x0002).tobytes(), dtype=F16)[0])
L_CAP = F32(64.0)


def h(x):
    return np.asarray(x, dtype=F16)


def records(vertices16, triangles):
    """a = v2, b = v1 - a, c = v0 - a, n = b x c with fp16 roundings (ray_casting.py:34-40)."""
    v = vertices16.numpy().astype(F16)
    t = triangles.numpy()
    v0, v1, a = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]
    b = v1 - a
    c = v0 - a
    n = np.stack((b[:, 1] * c[:, 2] - b[:, 2] * c[:, 1], b[:, 2] * c[:, 0] - b[:, 0] * c[:, 2],
                  b[:, 0] * c[:, 1] - b[:, 1] * c[:, 0]), 1)
    return a, b, c, n


def fma16(a, b, c):
    return (a.astype(np.float64) * np.float64(b) + np.float64(c)).astype(F16)


def prefilter_pass(s, d, a, b, c, n):
    """Brute force: s [P,3], d [3], triangle arrays [T,3] (fp16) -> bool [P,T] of the packed-fp16 pre-filter."""
    det = (n[:, 0] * d[0] + n[:, 1] * d[1]) + n[:, 2] * d[2]
    sgn = np.signbit(det)
    da = np.abs(det)
    tlo = -fma16(da, C1, TINY)
    thi = fma16(da, C2, TINY)
    g = [s[:, None, i] - a[None, :, i] for i in range(3)]
    cc = [c[None, :, i] for i in range(3)]
    bb = [b[None, :, i] for i in range(3)]
    ux = g[1] * cc[2] - g[2] * cc[1]
    uy = g[2] * cc[0] - g[0] * cc[2]
    uz = g[0] * cc[1] - g[1] * cc[0]
    Nn = (ux * d[0] + uy * d[1]) + uz * d[2]
    vx = bb[1] * g[2] - bb[2] * g[1]
    vy = bb[2] * g[0] - bb[0] * g[2]
    vz = bb[0] * g[1] - bb[1] * g[0]
    Mn = (vx * d[0] + vy * d[1]) + vz * d[2]
    Ns = np.where(sgn[None, :], -Nn, Nn)
    Ms = np.where(sgn[None, :], -Mn, Mn)
    with np.errstate(invalid="ignore"):
        return (Ns >= tlo[None, :]) & (Ms >= tlo[None, :]) & ((Ns + Ms) <= thi[None, :])


def half_ulp(maxabs):
    """largest |fp16(x) - x| for |x| <= maxabs (maxabs itself an fp16 value)."""
    m = np.float32(maxabs)
    if not np.isfinite(m):
        return F32(np.inf)
    e = math.floor(math.log2(m)) if m >= 2.0 ** -14 else -15
    return F32(max(2.0 ** (e - 11), 2.0 ** -25)) * F32(1.0 + 2.0 ** -10)


def plane_of(pos, trig, zlo, zhi):
    """Source plane of one env: nu . s = h for h in [hlo, hhi]; columns of the body transform (camera.py:197-199)."""
    sx, cx, sy, cy, sz, cz = [np.float64(v) for v in trig]
    c1 = np.array([cz * cy, -sz * cy, sy])
    c2 = np.array([sz * cx + cz * sy * sx, cz * cx - sz * sy * sx, -cy * sx])
    c3 = np.array([sz * sx - cz * sy * cx, cz * sx + sz * sy * cx, cy * cx])
    nu = np.cross(c1, c2)
    t = pos.astype(np.float64)
    h0 = nu @ t
    k = nu @ c3
    hs = sorted([h0 + zlo * k, h0 + zhi * k])
    return nu.astype(F32), F32(hs[0]), F32(hs[1])


class EnvConsts:
    """Per (env, ray tile) constants of the culling stages (computed once per CTA in the kernel)."""

    def __init__(self, pos, trig, zlo, zhi, d16, s16):
        self.d16 = d16
        self.d = d16.astype(F32)
        self.nu, hlo, hhi = plane_of(pos, trig, zlo, zhi)
        q = np.array([half_ulp(np.abs(s16[:, i]).max()) for i in range(3)], dtype=F32)
        # actual (fp16-rounded) sources: nu . s lies within tau of the ideal plane family [hlo, hhi]
        tau = (np.abs(self.nu) @ q) * F32(1.001) + F32(1e-6) * (np.abs(hlo) + np.abs(hhi))
        self.hlo, self.hhi = F32(hlo - tau), F32(hhi + tau)
        self.hmid = F32(0.5) * (self.hlo + self.hhi)
        self.nd = F32(self.nu @ self.d)
        self.inv_nd = F32(1.0) / self.nd
        dn = F32(np.sqrt((self.d * self.d).sum()))
        nn = F32(np.sqrt((self.nu * self.nu).sum()))
        self.dn = dn * F32(1.0 + 1e-6)
        self.kappa = dn * nn / np.abs(self.nd) * F32(1.0 + 1e-5)            # norm of the projection along d onto the plane
        self.lam = dn * (F32(0.5) * (self.hhi - self.hlo)) / np.abs(self.nd) * F32(1.0 + 1e-5)
        self.smax = F32(np.abs(s16.astype(F32)).max())


SQ3 = F32(1.7320509)
EPS0 = F32(0.1057)            # >= max(C_lo, C_hi - 1) of the pre-filter thresholds, outward rounded
OVF = F32(16000.0)


def stage1(a16, b16, c16, ec, rlo, rhi):
    """Cheap test per (superblock item, triangle).  rlo/rhi [2]: xy rectangle holding the sources of the item's rays.
    -> (reject, ill, gball): reject = no source of the rectangle can pass; ill = no bound (must be treated as
    'every ray of the item'); gball >= |s - a| (2-norm) for every passing source."""
    a, b, c = a16.astype(F32), b16.astype(F32), c16.astype(F32)
    d = ec.d
    ctr = a + (b + c) * F32(1.0 / 3.0)
    bc = b - c
    e2 = np.maximum(np.maximum((b * b).sum(1), (c * c).sum(1)), (bc * bc).sum(1))
    r = F32(2.0 / 3.0) * np.sqrt(e2) * F32(1.0 + 1e-5) + F32(1e-6) * np.abs(a).max(1)     # centroid -> farthest vertex (+ fp32 slack)
    nx = np.stack((b[:, 1] * c[:, 2] - b[:, 2] * c[:, 1], b[:, 2] * c[:, 0] - b[:, 0] * c[:, 2], b[:, 0] * c[:, 1] - b[:, 1] * c[:, 0]), 1)
    adet = np.abs(nx @ d)
    cb = np.maximum(np.abs(b).max(1), np.abs(c).max(1))
    amax = np.abs(a).max(1)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        # |det_h - det*| <= e_det: n16 and the dot product, <= 6 roundings on 6 terms each <= cb^2 * |d|max
        e_det = GAMMA * F32(6.1) * cb * cb + ALPHA
        rdet = F32(1.0) / adet
        eps0 = EPS0 + F32(1.3) * (e_det + F32(4) * ALPHA) * rdet
        rho = GAMMA * F32(2.01) * cb * rdet                    # d(eps)/d(Gsum) for l1, l2; l3 - 1 grows twice as fast
        tc = (ec.hmid - ctr @ ec.nu) * ec.inv_nd
        B = SQ3 * (ec.kappa * r * (F32(1) + F32(3) * eps0) + ec.lam + np.abs(tc) * ec.dn + r)
        den = F32(1) - F32(10.4) * ec.kappa * r * rho
        gsum = B / den
        eps = eps0 + F32(2) * rho * gsum
        R = ec.kappa * r * (F32(1) + F32(3) * eps) + ec.lam
        gball = R + np.abs(tc) * ec.dn + r
        ill = ~((den > F32(0.5)) & (adet > F32(4) * e_det) & (eps <= F32(16.0)) &
                ((ec.smax + amax) * np.maximum(cb, F32(1)) <= OVF))
        qx = ctr[:, 0] + tc * d[0]
        qy = ctr[:, 1] + tc * d[1]
        Rs = R * F32(1.0 + 1e-5) + F32(1e-5) * (np.abs(qx) + np.abs(qy))
        out = (qx + Rs < rlo[0]) | (qx - Rs > rhi[0]) | (qy + Rs < rlo[1]) | (qy - Rs > rhi[1])
    return out & ~ill, ill, gball


def chunk_records(a16, b16, c16, win=32):
    """Per window of `win` consecutive list entries: bounds of the stage-1 inputs (terrain.cu: sb_chunk_kernel)."""
    a, b, c = a16.astype(F32), b16.astype(F32), c16.astype(F32)
    ctr = a + (b + c) * F32(1.0 / 3.0)
    bc = b - c
    e2 = np.maximum(np.maximum((b * b).sum(1), (c * c).sum(1)), (bc * bc).sum(1))
    r = F32(2.0 / 3.0) * np.sqrt(e2) * F32(1.0 + 1e-5) + F32(1e-6) * np.abs(a).max(1)
    nx = np.stack((b[:, 1] * c[:, 2] - b[:, 2] * c[:, 1], b[:, 2] * c[:, 0] - b[:, 0] * c[:, 2], b[:, 0] * c[:, 1] - b[:, 1] * c[:, 0]), 1)
    cb = np.maximum(np.abs(b).max(1), np.abs(c).max(1))
    amax = np.abs(a).max(1)
    recs = []
    for w0 in range(0, a.shape[0], win):
        sl = slice(w0, w0 + win)
        recs.append(dict(lo=ctr[sl].min(0), hi=ctr[sl].max(0), rmax=r[sl].max(), cbmax=cb[sl].max(), amax=amax[sl].max(),
                         nlo=nx[sl].min(0), nhi=nx[sl].max(0), sl=sl))
    return recs


def chunk_cull(rec, ec, rlo, rhi):
    """True = NO triangle of the window can be kept by stage 1 for this rectangle (csrc/raycast_shadow.cu: chunk_cull)."""
    d, nu = ec.d, ec.nu
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        lo = sum(min(rec["nlo"][i] * d[i], rec["nhi"][i] * d[i]) for i in range(3))
        hi = sum(max(rec["nlo"][i] * d[i], rec["nhi"][i] * d[i]) for i in range(3))
        adet = F32(max(lo, -hi))
        adet = adet * F32(1 - 1e-5) - F32(1e-7) * (abs(lo) + abs(hi))
        if not (adet > 0):
            return False
        cbm, rm = rec["cbmax"], rec["rmax"]
        e_det = GAMMA * F32(6.1) * cbm * cbm + ALPHA
        rdet = F32(1.0) / adet * F32(1.0001)
        eps0 = EPS0 + F32(1.3) * (e_det + F32(4) * ALPHA) * rdet
        rho = GAMMA * F32(2.01) * cbm * rdet
        hl = sum(min(nu[i] * rec["lo"][i], nu[i] * rec["hi"][i]) for i in range(3))
        hh = sum(max(nu[i] * rec["lo"][i], nu[i] * rec["hi"][i]) for i in range(3))
        sl_h = F32(1e-6) * (abs(hl) + abs(hh)) + F32(1e-7)
        ta, tb = (ec.hmid - (hl - sl_h)) * ec.inv_nd, (ec.hmid - (hh + sl_h)) * ec.inv_nd
        tlo, thi = min(ta, tb), max(ta, tb)
        atc = max(abs(tlo), abs(thi)) * ec.dn
        kr = ec.kappa * rm
        B = SQ3 * (kr * (F32(1) + F32(3) * eps0) + ec.lam + atc + rm)
        den = F32(1) - F32(10.4) * kr * rho
        if not (den > F32(0.5)):
            return False
        gsum = B / den * F32(1.0001)
        eps = eps0 + F32(2) * rho * gsum
        fine = (adet > F32(4) * e_det) and (eps <= F32(16.0)) and ((ec.smax + rec["amax"]) * max(cbm, F32(1)) <= OVF)
        if not fine:
            return False
        R = (kr * (F32(1) + F32(3) * eps) + ec.lam) * F32(1.0001)
        qxl = rec["lo"][0] + min(tlo * d[0], thi * d[0]); qxh = rec["hi"][0] + max(tlo * d[0], thi * d[0])
        qyl = rec["lo"][1] + min(tlo * d[1], thi * d[1]); qyh = rec["hi"][1] + max(tlo * d[1], thi * d[1])
        Rs = R * F32(1.00001) + F32(1.1e-5) * (max(abs(qxl), abs(qxh)) + max(abs(qyl), abs(qyh))) + F32(1e-5)
        return bool((qxh + Rs < rlo[0]) or (qxl - Rs > rhi[0]) or (qyh + Rs < rlo[1]) or (qyl - Rs > rhi[1]))


def stage2(a16, b16, c16, n16, ec, gball):
    """Shadow box of the triangle's pre-filter prism on the source plane.  gball [T] >= |s - a| for passing sources.
    -> (x0, x1, y0, y1, full)."""
    a, b, c = a16.astype(F32), b16.astype(F32), c16.astype(F32)
    d, d16, nu = ec.d, ec.d16, ec.nu
    det_h = (n16[:, 0] * d16[0] + n16[:, 1] * d16[1]) + n16[:, 2] * d16[2]
    da = np.abs(det_h)
    tlo = fma16(da, C1, TINY).astype(F32)          # = -tlo of the kernel's make_cand
    thi = fma16(da, C2, TINY).astype(F32)
    sgn = np.where(np.signbit(det_h), F32(-1), F32(1))
    w = np.stack((c[:, 1] * d[2] - c[:, 2] * d[1], c[:, 2] * d[0] - c[:, 0] * d[2], c[:, 0] * d[1] - c[:, 1] * d[0]), 1)
    ad = np.abs(d)
    ac, ab = np.abs(c), np.abs(b)
    saw = (ac[:, 1] + ac[:, 2]) * ad[0] + (ac[:, 2] + ac[:, 0]) * ad[1] + (ac[:, 0] + ac[:, 1]) * ad[2]      # sum_i aw_i
    sawp = (ab[:, 1] + ab[:, 2]) * ad[0] + (ab[:, 2] + ab[:, 0]) * ad[1] + (ab[:, 0] + ab[:, 1]) * ad[2]
    dets = (b * w).sum(1)
    adet = np.abs(dets)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        rdet = F32(1.0) / adet * F32(1.0 + 2.0 ** -10)
        EN = GAMMA * gball * saw + ALPHA
        EM = GAMMA * gball * sawp + ALPHA
        l1 = (tlo + EN) * rdet
        l2 = (tlo + EM) * rdet
        l3 = (thi * F32(1.0 + 2.0 ** -10) + F32(2.0 ** -24) + EN + EM) * rdet
        good = (np.sign(dets) == sgn) & (np.maximum(np.maximum(l1, l2), l3) <= L_CAP)
        cn = np.stack((-l1, l3 + l2, -l1), 1)           # corners A, B, C in (n, m)
        cm = np.stack((-l2, -l2, l3 + l1), 1)
        xs, ys = [], []
        for k in range(3):
            P = a + cn[:, k:k + 1] * b + cm[:, k:k + 1] * c
            nP = P @ nu
            for hh in (ec.hlo, ec.hhi):
                t = (hh - nP) * ec.inv_nd
                xs.append(P[:, 0] + t * d[0])
                ys.append(P[:, 1] + t * d[1])
        xs, ys = np.stack(xs, 1), np.stack(ys, 1)
        sl = (np.abs(xs).max(1) + np.abs(ys).max(1)) * F32(2.0 ** -18) + F32(2.0 ** -20)
    # thresholds of the per-ray linear test (stage 3L of the kernel): the prism itself, un-normalised
    with np.errstate(invalid="ignore", over="ignore"):
        T1 = bf16_up((tlo + EN) * LIN_SLACK)
        T2 = bf16_up((tlo + EM) * LIN_SLACK)
        T3 = bf16_up((thi * F32(1.0 + 2.0 ** -10) + F32(2.0 ** -24) + EN + EM) * LIN_SLACK)
    return xs.min(1) - sl, xs.max(1) + sl, ys.min(1) - sl, ys.max(1) + sl, ~good, (T1, T2, T3, sgn)


LIN_SLACK = F32(1.0005)       # fp32 evaluation of N*, M* (error <= 2^-13 EN) + the threshold sums


def bf16_up(x):
    """positive fp32 -> the next value with 16 significant bits kept (how the kernel packs the thresholds into a task)."""
    u = np.asarray(x, dtype=F32).view(np.uint32).astype(np.uint64)
    u = ((u + 0xFFFF) & 0xFFFF0000).astype(np.uint32)
    return u.view(F32)


def lin_test(s16, a16, b16, c16, ec, T, full):
    """Stage 3L: per (ray, triangle) the exact linear forms N* = ((s-a) x c).d, M* = (b x (s-a)).d in fp32 against the
    thresholds of stage 2.  -> bool [P,T] (True = may pass the pre-filter; `full` triangles pass every ray)."""
    T1, T2, T3, sgn = T
    a, b, c = a16.astype(F32), b16.astype(F32), c16.astype(F32)
    ds = ec.d[None, :] * sgn[:, None]                                   # +-d
    w1 = np.stack((c[:, 1] * ds[:, 2] - c[:, 2] * ds[:, 1], c[:, 2] * ds[:, 0] - c[:, 0] * ds[:, 2], c[:, 0] * ds[:, 1] - c[:, 1] * ds[:, 0]), 1)
    w2 = np.stack((ds[:, 1] * b[:, 2] - ds[:, 2] * b[:, 1], ds[:, 2] * b[:, 0] - ds[:, 0] * b[:, 2], ds[:, 0] * b[:, 1] - ds[:, 1] * b[:, 0]), 1)
    s = s16.astype(F32)
    g = [s[:, None, i] - a[None, :, i] for i in range(3)]
    Nn = (g[0] * w1[None, :, 0] + g[1] * w1[None, :, 1]) + g[2] * w1[None, :, 2]
    Mn = (g[0] * w2[None, :, 0] + g[1] * w2[None, :, 1]) + g[2] * w2[None, :, 2]
    with np.errstate(invalid="ignore"):
        ok = (Nn >= -T1[None]) & (Mn >= -T2[None]) & (Nn + Mn <= T3[None])
    return ok | full[None, :]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=12)
    ap.add_argument("--length", type=float, default=200.0)
    ap.add_argument("--nv", type=int, default=708)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--wild", type=float, default=0.25)
    ap.add_argument("--norefine", action="store_true")
    args = ap.parse_args()
    hf, stones = R.synth.make_heightfield(args.length, args.nv, 200, args.seed)
    v, t = R.synth.heightfield_to_mesh(hf, args.length)
    v16 = torch.from_numpy(v).to(torch.float16)
    a, b, c, n = records(v16, torch.from_numpy(t))
    print("mesh: %d triangles, spacing %.3f m" % (a.shape[0], args.length / (args.nv - 1)))
    pat, _, _ = O.heightmap_pattern()
    zlo, zhi = float(pat[:, 2].min()), float(pat[:, 2].max())
    g = torch.Generator().manual_seed(args.seed)
    N = args.envs
    L = args.length
    xy = torch.rand(N, 2, generator=g) * (L - 10) + 5
    spacing = L / (args.nv - 1)
    ij = np.clip(np.rint(xy.numpy() / spacing).astype(int), 0, args.nv - 1)
    z = torch.from_numpy(hf[ij[:, 0], ij[:, 1]].copy()).float() + 0.5 + torch.rand(N, generator=g) * 0.05
    pos = torch.cat((xy, z[:, None]), 1).float()
    wild = torch.rand(N, generator=g) < args.wild
    rp = torch.where(wild[:, None], torch.rand(N, 2, generator=g) * 2.6 - 1.3, torch.rand(N, 2, generator=g) * 0.4 - 0.2)
    yaw = torch.rand(N, generator=g) * 2 * math.pi - math.pi
    euler = torch.cat((rp, yaw[:, None]), 1).float()
    src, dirs = O.depth_transform(pos, euler, pat)
    dneg = -torch.nn.functional.normalize(dirs)
    trig = torch.cat(O._neg_trig(euler), 1).numpy()
    tot_pass = tot_box = tot_tri = tot_full = tot_viol = 0
    stats = {}
    for e in range(N):
        s = src[e].numpy().astype(F16)
        d = dneg[e].numpy().astype(F16)
        lo, hi = s.astype(F32).min(0), s.astype(F32).max(0)
        sel = np.nonzero((a[:, 0].astype(F32) > lo[0] - 2.5) & (a[:, 0].astype(F32) < hi[0] + 2.5) &
                         (a[:, 1].astype(F32) > lo[1] - 2.5) & (a[:, 1].astype(F32) < hi[1] + 2.5))[0]
        if sel.size > 12000:
            sel = sel[np.random.RandomState(e).permutation(sel.size)[:12000]]
        ps = prefilter_pass(s, d, a[sel], b[sel], c[sel], n[sel])          # [P, T]
        ec = EnvConsts(pos[e].numpy(), trig[e], zlo, zhi, d, s)
        sx, sy = s[:, 0].astype(F32), s[:, 1].astype(F32)
        crecs = chunk_records(a[sel], b[sel], c[sel])          # `sel` is sorted by triangle id, like a superblock list
        # stage 1 against random 2.4 m windows of the ray set (what a superblock item is)
        rs = np.random.RandomState(e)
        for _ in range(6):
            cxy = np.array([rs.uniform(lo[0], hi[0]), rs.uniform(lo[1], hi[1])], dtype=F32)
            inw = (np.abs(sx - cxy[0]) <= 1.2) & (np.abs(sy - cxy[1]) <= 1.2)
            if not inw.any():
                continue
            rlo = np.array([sx[inw].min(), sy[inw].min()], dtype=F32)
            rhi = np.array([sx[inw].max(), sy[inw].max()], dtype=F32)
            rej, ill, gb = stage1(a[sel], b[sel], c[sel], ec, rlo, rhi)
            for rec in crecs:
                if chunk_cull(rec, ec, rlo, rhi):
                    stats["chunk_culled"] = stats.get("chunk_culled", 0) + 1
                    stats["chunk_viol"] = stats.get("chunk_viol", 0) + int((~rej[rec["sl"]]).sum())
                stats["chunks"] = stats.get("chunks", 0) + 1
            bad = ps[inw][:, rej].any()
            stats["s1_viol"] = stats.get("s1_viol", 0) + int(bad)
            stats["s1_keep"] = stats.get("s1_keep", 0) + float((~rej).mean()) / 6
            stats["s1_ill"] = stats.get("s1_ill", 0) + float(ill.mean()) / 6
        rej, ill, gb = stage1(a[sel], b[sel], c[sel], ec, lo[:2], hi[:2])
        gtrue = np.sqrt(((s.astype(F32)[:, None, :] - a[sel].astype(F32)[None]) ** 2).sum(2))
        gviol = (ps & (gtrue > gb[None, :]) & ~ill[None, :]).sum()
        stats["gball_viol"] = stats.get("gball_viol", 0) + int(gviol) + int(ps[:, rej].any())
        x0, x1, y0, y1, full, T = stage2(a[sel], b[sel], c[sel], n[sel], ec, gb)
        full = full | ill
        lin = lin_test(s, a[sel], b[sel], c[sel], ec, T, full)
        stats["lin_viol"] = stats.get("lin_viol", 0) + int((ps & ~lin).sum())
        stats["lin_pass"] = stats.get("lin_pass", 0) + int((lin & ~rej[None, :] & ~full[None, :]).sum())
        stats["pre_pass_bounded"] = stats.get("pre_pass_bounded", 0) + int((ps & ~full[None, :]).sum())
        stats["pre_pass"] = stats.get("pre_pass", 0) + int(ps.sum())
        stats["s2_in"] = stats.get("s2_in", 0) + int((~rej).sum())
        inbox = ((sx[:, None] >= x0[None, :]) & (sx[:, None] <= x1[None, :]) & (sy[:, None] >= y0[None, :]) &
                 (sy[:, None] <= y1[None, :])) | full[None, :]
        viol = ps & ~inbox
        for nm, cs in (("cell", 0.1), ("blk", 0.3)):
            cx = np.rint(sx / F32(0.1)).astype(np.int64) // int(round(cs / 0.1))
            cy = np.rint(sy / F32(0.1)).astype(np.int64) // int(round(cs / 0.1))
            bx0 = np.rint(x0 / F32(0.1)).astype(np.int64) // int(round(cs / 0.1)); bx1 = np.rint(x1 / F32(0.1)).astype(np.int64) // int(round(cs / 0.1))
            by0 = np.rint(y0 / F32(0.1)).astype(np.int64) // int(round(cs / 0.1)); by1 = np.rint(y1 / F32(0.1)).astype(np.int64) // int(round(cs / 0.1))
            inb = ((cx[:, None] >= bx0[None]) & (cx[:, None] <= bx1[None]) & (cy[:, None] >= by0[None]) & (cy[:, None] <= by1[None])) | full[None]
            stats[nm] = stats.get(nm, 0) + int(inb.sum())
            print("     %s-quantised box: %d pair tests" % (nm, int(inb.sum())))
            assert not (ps & ~inb).any()
            if nm == "blk":
                stats["tasks"] = stats.get("tasks", 0) + int((np.minimum(bx1, cx.max()) - np.maximum(bx0, cx.min()) + 1).clip(0)[inb.any(0)].sum())
                stats["tri_any"] = stats.get("tri_any", 0) + int(inb.any(0).sum())
        tot_viol += int(viol.sum())
        tot_pass += int(ps.sum())
        tot_box += int(inbox.sum())
        tot_tri += sel.size
        tot_full += int(full.sum())
        print("env %2d roll %+.2f pitch %+.2f |xy| %.0f  tris %5d  pass %5d  in-box %6d  full %3d  violations %d" % (
            e, euler[e, 0], euler[e, 1], float(pos[e, :2].abs().max()), sel.size, int(ps.sum()), int(inbox.sum()),
            int(full.sum()), int(viol.sum())))
    print("TOTAL pass %d, box tests %d (%.2fx), triangles %d, full-range %d, violations %d" % (
        tot_pass, tot_box, tot_box / max(tot_pass, 1), tot_tri, tot_full, tot_viol))
    print({k: v / N for k, v in stats.items()})
    return 0 if (tot_viol == 0 and stats.get("lin_viol", 0) == 0 and stats.get("chunk_viol", 0) == 0) else 1


if __name__ == "__main__":
    sys.exit(main())
