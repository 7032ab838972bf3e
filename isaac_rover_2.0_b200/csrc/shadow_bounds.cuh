// The conservative bounds of the "shadow" culling (derivation: header of raycast_shadow.cu; numpy restatement and brute-force proof:
// tests/shadow_proto.py).  They hold for ANY family of sources with nu . s in [hlo, hhi] and ONE direction d per family; the
// heightmap pattern of an env (sources on the camera plane) is the family the production kernel uses them for.  (The four points of
// a wheel are such a family too; DESIGN.md 4.1 "measured and rejected" says why the rock kernel does not cull with them.)
#pragma once
#include "raycast_common.cuh"

namespace rc {

constexpr float GAMMA = 0.00390625f;            // 2^-8
constexpr float ALPHA = 1.9073486328125e-06f;   // 2^-19
constexpr float EPS0 = 0.1057f;
constexpr float L_CAP = 64.0f;
constexpr float OVF = 16000.0f;
constexpr float SQ3 = 1.7320509f;
constexpr float LIN_SLACK = 1.0005f;            // stage 3L: fp32 evaluation of N*, M* (error <= 2^-13 of E_N, E_M) + threshold sums

struct EnvC {
    float dx, dy, dz;               // d = -normalize(dir) as the reference rounds it (fp16 values)
    float nux, nuy, nuz;            // normal of the source plane
    float inv_nd, hmid, hlo, hhi;   // nu . s in [hlo, hhi] for every source of the tile
    float kappa, lam, dn, smax;
};

__device__ __forceinline__ float rcp_up(float x) { return __fdividef(1.0f, x) * 1.000001f; }
__device__ __forceinline__ float hf(uint32_t bits16) { return __half2float(__ushort_as_half((unsigned short)bits16)); }
// Stage 1 (tests/shadow_proto.py: stage1).  Returns false if no source inside the rectangle can pass; gball >= |s - a|
// for every passing source (+inf: no bound, the caller must test every ray of the item).
__device__ __forceinline__ bool stage1(const uint4& s0, const uint4& s1, const EnvC& e, float rlox, float rhix, float rloy, float rhiy, float& gball) {
    // S1Rec (common.cuh): centroid, radius, b x c, magnitudes -- computed once per layer by build_records_kernel (terrain.cu)
    const float qx0 = __uint_as_float(s0.x), qy0 = __uint_as_float(s0.y), qz0 = __uint_as_float(s0.z), r = __uint_as_float(s0.w);
    const float nx = __uint_as_float(s1.x), ny = __uint_as_float(s1.y), nz = __uint_as_float(s1.z);
    const float cb = hf(s1.w & 0xffffu), amax = hf(s1.w >> 16);
    const float adet = fabsf(fmaf(nx, e.dx, fmaf(ny, e.dy, nz * e.dz)));
    const float e_det = GAMMA * 6.1f * cb * cb + ALPHA;
    const float rdet = rcp_up(adet);
    const float eps0 = EPS0 + 1.3f * (e_det + 4.0f * ALPHA) * rdet;
    const float rho = GAMMA * 2.01f * cb * rdet;
    const float tc = (e.hmid - fmaf(qx0, e.nux, fmaf(qy0, e.nuy, qz0 * e.nuz))) * e.inv_nd;
    const float atc = fabsf(tc) * e.dn;
    const float kr = e.kappa * r;
    const float Bn = SQ3 * (kr * (1.0f + 3.0f * eps0) + e.lam + atc + r);
    const float den = 1.0f - 10.4f * kr * rho;
    const float gsum = Bn * rcp_up(den);
    const float eps = eps0 + 2.0f * rho * gsum;
    const float R = kr * (1.0f + 3.0f * eps) + e.lam;
    const bool fine = (den > 0.5f) && (adet > 4.0f * e_det) && (eps <= 16.0f) && ((e.smax + amax) * fmaxf(cb, 1.0f) <= OVF);
    if (!fine) {                     // also every NaN case
        gball = __int_as_float(0x7f800000);
        return true;
    }
    gball = R + atc + r;
    const float qx = fmaf(tc, e.dx, qx0), qy = fmaf(tc, e.dy, qy0);
    const float Rs = R * 1.00001f + 1e-5f * (fabsf(qx) + fabsf(qy));
    const bool out = (qx + Rs < rlox) || (qx - Rs > rhix) || (qy + Rs < rloy) || (qy - Rs > rhiy);
    return !out;
}

// Window cull (tests/shadow_proto.py: chunk_cull): true = stage 1 would reject EVERY triangle whose record went into `c` for this
// rectangle, so the window's 32 list entries need not be enumerated.  Stage 1's formulas on the window's worst-case inputs
// (every quantity is monotone in them); any NaN makes a comparison false and the window is kept.
__device__ __forceinline__ bool chunk_cull(const ChunkRec& c, const EnvC& e, float rlox, float rhix, float rloy, float rhiy, float& overlap) {
    overlap = 1.0f;          // fraction of the window's (expanded) footprint that lies inside the rectangle: a cost proxy only
    const float lo = fminf(c.nxlo * e.dx, c.nxhi * e.dx) + fminf(c.nylo * e.dy, c.nyhi * e.dy) + fminf(c.nzlo * e.dz, c.nzhi * e.dz);
    const float hi = fmaxf(c.nxlo * e.dx, c.nxhi * e.dx) + fmaxf(c.nylo * e.dy, c.nyhi * e.dy) + fmaxf(c.nzlo * e.dz, c.nzhi * e.dz);
    float adet = fmaxf(lo, -hi);                                 // |n . d| >= adet for every triangle of the window
    adet = adet * 0.99999f - 1e-7f * (fabsf(lo) + fabsf(hi));
    if (!(adet > 0.0f) || !(c.x0 == c.x0)) return false;
    const float e_det = GAMMA * 6.1f * c.cbmax * c.cbmax + ALPHA;
    const float rdet = rcp_up(adet) * 1.0001f;
    const float eps0 = EPS0 + 1.3f * (e_det + 4.0f * ALPHA) * rdet;
    const float rho = GAMMA * 2.01f * c.cbmax * rdet;
    const float hl = fminf(e.nux * c.x0, e.nux * c.x1) + fminf(e.nuy * c.y0, e.nuy * c.y1) + fminf(e.nuz * c.z0, e.nuz * c.z1);
    const float hh = fmaxf(e.nux * c.x0, e.nux * c.x1) + fmaxf(e.nuy * c.y0, e.nuy * c.y1) + fmaxf(e.nuz * c.z0, e.nuz * c.z1);
    const float slh = 1e-6f * (fabsf(hl) + fabsf(hh)) + 1e-7f;
    const float ta = (e.hmid - (hl - slh)) * e.inv_nd, tb = (e.hmid - (hh + slh)) * e.inv_nd;
    const float tlo = fminf(ta, tb), thi = fmaxf(ta, tb);
    const float atc = fmaxf(fabsf(tlo), fabsf(thi)) * e.dn;
    const float kr = e.kappa * c.rmax;
    const float Bn = SQ3 * (kr * (1.0f + 3.0f * eps0) + e.lam + atc + c.rmax);
    const float den = 1.0f - 10.4f * kr * rho;
    if (!(den > 0.5f)) return false;
    const float gsum = Bn * rcp_up(den) * 1.0001f;
    const float eps = eps0 + 2.0f * rho * gsum;
    const bool fine = (adet > 4.0f * e_det) && (eps <= 16.0f) && ((e.smax + c.amax) * fmaxf(c.cbmax, 1.0f) <= OVF);
    if (!fine) return false;
    const float R = (kr * (1.0f + 3.0f * eps) + e.lam) * 1.0001f;
    const float qxl = c.x0 + fminf(tlo * e.dx, thi * e.dx), qxh = c.x1 + fmaxf(tlo * e.dx, thi * e.dx);
    const float qyl = c.y0 + fminf(tlo * e.dy, thi * e.dy), qyh = c.y1 + fmaxf(tlo * e.dy, thi * e.dy);
    const float Rs = R * 1.00001f + 1.1e-5f * (fmaxf(fabsf(qxl), fabsf(qxh)) + fmaxf(fabsf(qyl), fabsf(qyh))) + 1e-5f;
    const float ex0 = qxl - Rs, ex1 = qxh + Rs, ey0 = qyl - Rs, ey1 = qyh + Rs;
    const float ox = fminf(ex1, rhix) - fmaxf(ex0, rlox), oy = fminf(ey1, rhiy) - fmaxf(ey0, rloy);
    overlap = fminf(fmaxf(ox, 0.0f) * fmaxf(oy, 0.0f) * __fdividef(1.0f, (ex1 - ex0) * (ey1 - ey0)), 1.0f);
    return (qxh + Rs < rlox) || (qxl - Rs > rhix) || (qyh + Rs < rloy) || (qyl - Rs > rhiy);
}

}  // namespace rc
