// Pieces shared by the heightmap ray-cast kernels (raycast_tiled.cu, raycast_shadow.cu).
#pragma once
#include "common.cuh"

namespace rc {

constexpr uint32_t KEY_INIT = 0xC9800000u;     // (11.0, slot 0): the result when every candidate misses
constexpr uint32_t ORD_MISS = 0xC980u;        // order key of fp16 11.0 (0x4980 | 0x8000)

struct TiledParams {
    const int32_t* index;
    const TriRec* recs;
    const S1Rec* s1;          // stage-1 records of the shadow kernel
    const uint32_t* blk_off;
    const int32_t* blk_ids;
    const uint4* blk_slots;
    int nBy;
    int G0, G1, K, Ks;
    float res, inv_res, shift_x, shift_y;
    int sem;
    const float* pos;
    const float* euler;
    const float* trig;
    const double* pattern;
    int P, tiles, tile_size;
    __half* dist;
    int32_t* hit_slot;
    int32_t* hit_tri;
    __half* pt;
    __half* sources;
    float* obs;
    int64_t obs_ld;
    __half* obs16;            // optional packed observation: obs16[n * obs16_ld + (column - obs16_col0)] = fp16(dist / 2)
    int64_t obs16_ld;
    int obs16_col0;
    const int32_t* col_a;
    const int32_t* col_b;
    // shadow kernel (raycast_shadow.cu): superblock candidate lists + the list of (env, tile) work items it hands back
    const uint32_t* sb_off;
    const int32_t* sb_ids;
    const unsigned char* sb_slot9;
    const ChunkRec* sb_chunk;   // bounds per window of 32 list entries (chunk culling); NULL = none
    int nSBy;
    float cos_steep;          // envs whose ray direction is flatter than this go to the fall-back list
    int* fb_count;            // [1]
    int32_t* fb_list;         // [N * tiles]  env * tiles + tile
    // tiled kernel in work-list mode: CTAs loop over fb_list[0 .. *fb_count) instead of blockIdx
    const int* work_count;
    const int32_t* work_list;
    int work_slices;
    // shadow kernel: env order (envs with tilted rays first: they are the slow CTAs, so they must not start last) and the
    // pre-sorted steep envs (cast by the tiled kernel on a second stream, concurrently); both written by hm_classify_kernel
    const int32_t* order;     // [N] or NULL
    int presorted;            // steep envs are already on a work list: the shadow kernel just skips them
    int min_sh;               // bins are 2^sh x 2^sh cells, sh >= min_sh (0 = one cell; tuning hook)
    int spec_slot;            // stage 3b fetches the slot byte before the literal test (tuning hook)
    int task_rays;            // shadow kernel: rays per task of stage 3L (<= 16; tuning hook)
    int bulk_obs;             // shadow kernel: observation row leaves through one bulk asynchronous copy (A/B switch)
    int n_obs_cols;           // number of heightmap observation columns (sparse + dense) when obs is requested
    int64_t split_from;       // shadow kernel: order positions >= split_from are cast by two CTAs (half the rays each); -1 = none
    int split_at;             // first ray of the second half
    unsigned long long* dbg;  // optional [24] work counters / cycle counts of the shadow kernel (RVB_SHADOW_DBG=1)
};

// two candidates of one cell, component-wise packed: .x = candidate j, .y = candidate j+1
struct Tri2 {
    __half2 ax, ay, az, bx, by, bz, cx, cy, cz, nx, ny, nz;
};

__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 add2(__half2 a, __half2 b) { return __hadd2_rn(a, b); }
__device__ __forceinline__ __half2 sub2(__half2 a, __half2 b) { return __hsub2_rn(a, b); }
__device__ __forceinline__ __half2 mul2(__half2 a, __half2 b) { return __hmul2_rn(a, b); }

__device__ __forceinline__ Tri2 pack_tri2(const uint4& p0, const uint2& p1, const uint4& q0, const uint2& q1) {
    // record halves: q0 = [a.x a.y | a.z b.x | b.y b.z | c.x c.y], q1 = [c.z n.x | n.y n.z]
    Tri2 t;
    t.ax = u2h(__byte_perm(p0.x, q0.x, 0x5410)); t.ay = u2h(__byte_perm(p0.x, q0.x, 0x7632));
    t.az = u2h(__byte_perm(p0.y, q0.y, 0x5410)); t.bx = u2h(__byte_perm(p0.y, q0.y, 0x7632));
    t.by = u2h(__byte_perm(p0.z, q0.z, 0x5410)); t.bz = u2h(__byte_perm(p0.z, q0.z, 0x7632));
    t.cx = u2h(__byte_perm(p0.w, q0.w, 0x5410)); t.cy = u2h(__byte_perm(p0.w, q0.w, 0x7632));
    t.cz = u2h(__byte_perm(p1.x, q1.x, 0x5410)); t.nx = u2h(__byte_perm(p1.x, q1.x, 0x7632));
    t.ny = u2h(__byte_perm(p1.y, q1.y, 0x5410)); t.nz = u2h(__byte_perm(p1.y, q1.y, 0x7632));
    return t;
}

// ------------------------------------------------------------------------------------------------------------
// Conservative fp16 pre-filter.  With q_n = N/det, q_m = M/det (exact rationals of fp16 values), n = fp16(q_n),
// m = fp16(q_m), a candidate passes ray_casting.py:59 only if
//     n >= fp16(-0.1)        <=>  q_n >= -3277/32768   (rounding boundary below -0x1.998p-4, tie -> even = -0.1)
//     m >= fp16(-0.1)        <=>  q_m >= -3277/32768
//     fp16(n + m) <= fp16(1.1) =>  q_n + q_m < 2253/2048 + 2^-9   (n, m are within 2^-11 relative of q_n, q_m, both
//                                                                   in [-0.11, 1.31] there, so n + m moves by < 2^-9)
// Multiplying by |det| (N' = N*sgn(det), M' likewise):  N' >= -C_LO*|det|,  M' >= -C_LO*|det|,  N'+M' < C_FAIL*|det|.
// The filter evaluates these in packed fp16 against thresholds rounded OUTWARDS:
//     tlo = -RN(|det| * 0x2E68 + 2^-23)   <= -C_LO*|det|      (0x2E68 = 0.10009766 >= C_LO / (1 - 2^-11))
//     thi =  RN(|det| * 0x3C6B + 2^-23)   >= RN16(x) for every x < C_FAIL*|det|   (0x3C6B = 1.10449 >= C_FAIL*(1+2^-11)^2)
// (the 2^-23 term covers the absolute rounding error 2^-25 of subnormal results).  NaN thresholds or numerators
// compare false -- exactly the cases the reference rejects (det = 0/NaN gives n, m = +-inf/NaN; a NaN numerator
// gives a NaN quotient).  Every candidate the filter lets through is re-evaluated with the literal op sequence
// (pair_test, divisions included), so the filter only has to be conservative, never exact.
// ------------------------------------------------------------------------------------------------------------
#define H_C1 0x2E68u
#define H_C2 0x3C6Bu
#define H_TINY 0x0002u

struct Cand2 {
    uint32_t sgn;      // sign bits of det = (b x c) . d in both halves   (ray_casting.py:41; d is per env)
    __half2 tlo, thi;  // outward-rounded thresholds; NaN for a slot beyond K
};

__device__ __forceinline__ Cand2 make_cand2(const Tri2& t, __half2 dx, __half2 dy, __half2 dz, bool v0, bool v1) {
    Cand2 c;
    const __half2 det = add2(add2(mul2(t.nx, dx), mul2(t.ny, dy)), mul2(t.nz, dz));
    c.sgn = h2u(det) & 0x80008000u;
    const __half2 da = u2h(h2u(det) & 0x7fff7fffu);
    const __half2 tiny = u2h(H_TINY | (H_TINY << 16));
    c.tlo = __hneg2(__hfma2(da, u2h(H_C1 | (H_C1 << 16)), tiny));
    c.thi = __hfma2(da, u2h(H_C2 | (H_C2 << 16)), tiny);
    if (!v0) c.tlo = u2h((h2u(c.tlo) & 0xffff0000u) | 0x7fffu);
    if (!v1) c.tlo = u2h((h2u(c.tlo) & 0x0000ffffu) | 0x7fff0000u);
    return c;
}

// 0xffff in the half of every candidate that may pass
__device__ __forceinline__ uint32_t prefilter2(__half2 sx, __half2 sy, __half2 sz, __half2 dx, __half2 dy, __half2 dz,
                                               const Tri2& t, const Cand2& c) {
    const __half2 gx = sub2(sx, t.ax), gy = sub2(sy, t.ay), gz = sub2(sz, t.az);                 // ray_casting.py:37
    const __half2 ux = sub2(mul2(gy, t.cz), mul2(gz, t.cy));                                     // g x c  (:44)
    const __half2 uy = sub2(mul2(gz, t.cx), mul2(gx, t.cz));
    const __half2 uz = sub2(mul2(gx, t.cy), mul2(gy, t.cx));
    const __half2 Nn = add2(add2(mul2(ux, dx), mul2(uy, dy)), mul2(uz, dz));                     // :45 numerator
    const __half2 vx = sub2(mul2(t.by, gz), mul2(t.bz, gy));                                     // b x g  (:49)
    const __half2 vy = sub2(mul2(t.bz, gx), mul2(t.bx, gz));
    const __half2 vz = sub2(mul2(t.bx, gy), mul2(t.by, gx));
    const __half2 Mn = add2(add2(mul2(vx, dx), mul2(vy, dy)), mul2(vz, dz));                     // :50 numerator
    const __half2 Ns = u2h(h2u(Nn) ^ c.sgn), Ms = u2h(h2u(Mn) ^ c.sgn);
    return __hge2_mask(Ns, c.tlo) & __hge2_mask(Ms, c.tlo) & __hle2_mask(add2(Ns, Ms), c.thi);
}

// key of torch.min's total order: NaN < everything, -0 == +0, ties -> lower slot; bit 0 remembers a negative zero.
__device__ __forceinline__ uint32_t make_key(unsigned short b, uint32_t slot) {
    uint32_t ord, nz = 0;
    if ((b & 0x7fffu) > 0x7c00u) ord = 0u;
    else if ((b & 0x7fffu) == 0u) { ord = 0x8000u; nz = b >> 15; }
    else ord = (b & 0x8000u) ? (uint32_t)(unsigned short)~b : (uint32_t)(b | 0x8000u);
    return (ord << 16) | (slot << 1) | nz;
}
__device__ __forceinline__ unsigned short key_bits(uint32_t key) {
    const uint32_t ord = key >> 16;
    if (ord == 0u) return 0x7fffu;
    if (ord == 0x8000u) return (key & 1u) ? 0x8000u : 0u;
    return (ord & 0x8000u) ? (unsigned short)(ord & 0x7fffu) : (unsigned short)~ord;
}

__device__ __forceinline__ void unpack_rec(const TriRec* rec, H3& a, H3& b, H3& c, H3& n) {
    const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(rec));
    const uint2 q1 = __ldg(reinterpret_cast<const uint2*>(rec) + 2);
    a = {h_from_bits(q0.x & 0xffff), h_from_bits(q0.x >> 16), h_from_bits(q0.y & 0xffff)};
    b = {h_from_bits(q0.y >> 16), h_from_bits(q0.z & 0xffff), h_from_bits(q0.z >> 16)};
    c = {h_from_bits(q0.w & 0xffff), h_from_bits(q0.w >> 16), h_from_bits(q1.x & 0xffff)};
    n = {h_from_bits(q1.x >> 16), h_from_bits(q1.y & 0xffff), h_from_bits(q1.y >> 16)};
}

// A ray that holds a hit with k > 11.0 and no nearer one: the first slot whose value is <= 11.0 (a miss, or a hit at
// exactly 11.0) wins torch.min, and that slot is not necessarily 0.  Vanishingly rare (the rover would have to hover
// 11 m above the mesh), so one thread simply walks the K candidates with the literal op sequence.
static __device__ __noinline__ uint32_t literal_ray(const int32_t* row, int K, const TriRec* recs, H3 s, H3 d) {
    uint32_t best = 0xffffffffu;
    for (int j = 0; j < K; ++j) {
        H3 a, b, c, nn;
        unpack_rec(recs + __ldg(row + j), a, b, c, nn);
        best = min(best, make_key(h_bits(pair_test(s, d, a, b, c, nn)), (uint32_t)j));
    }
    return best;
}

// The same two look-ups for a layer whose K-contiguous index was released (rvb_terrain_release_index): the K-list of cell
// (cx, cy) is the set of entries of its 3x3 block's union list whose slot byte for the sub-cell is not 0xFF, and that byte IS the
// slot.  min() over keys that carry the slot does not depend on the order the candidates are visited in.
static __device__ __noinline__ uint32_t literal_ray_blk(const int32_t* ids, const uint4* slots, int n, int sub, const TriRec* recs,
                                                        H3 s, H3 d) {
    uint32_t best = 0xffffffffu;
    for (int e = 0; e < n; ++e) {
        const uint32_t slot = __ldg(reinterpret_cast<const unsigned char*>(slots + e) + sub);
        if (slot == 0xffu) continue;
        H3 a, b, c, nn;
        unpack_rec(recs + __ldg(ids + e), a, b, c, nn);
        best = min(best, make_key(h_bits(pair_test(s, d, a, b, c, nn)), slot));
    }
    return best;
}

static __device__ __noinline__ int32_t tri_of_slot_blk(const int32_t* ids, const uint4* slots, int n, int sub, int slot) {
    for (int e = 0; e < n; ++e)
        if ((int)__ldg(reinterpret_cast<const unsigned char*>(slots + e) + sub) == slot) return __ldg(ids + e);
    return 0;
}

// Epilogue in ray order: coalesced stores of dist / hit slot / hit triangle / pt / sources and the fused sparse+dense
// observation columns (heightmap_distribution.py:126-133, rover.py:324-325).  res[p] = best key of local ray p; a set bit
// in far[] marks a ray that saw a hit beyond the 11.0 miss sentinel (resolved literally here).
// `stage` (optional, shared memory, 16-byte aligned, >= (ncols + 8) floats): the observation row is assembled there and leaves
// the SM as ONE bulk asynchronous copy (cp.async.bulk shared -> global, the TMA unit's non-tensor mode) instead of one 4-byte
// store per column; only when the tile is the whole pattern.  ncols = number of heightmap observation columns.
__device__ __forceinline__ void epilogue(const TiledParams& q, int64_t n, int p0, int np, const Trig& tr, double tx, double ty,
                                         double tz, __half2 dx2, __half2 dy2, __half2 dz2, const uint32_t* res,
                                         const uint32_t* far, int tid, int nthreads, float* stage = nullptr, int ncols = 0) {
    const H3 dlit = {__low2half(dx2), __low2half(dy2), __low2half(dz2)};
    const bool want_geo = q.hit_tri || q.pt || q.sources;
    // destination of column 4 (the first heightmap column) and how the row splits into an unaligned head, a 16-byte aligned body
    // (the bulk copy) and a tail
    const bool bulk = stage != nullptr && q.obs != nullptr && np == q.P && ncols > 8;
    float* const grow = q.obs ? q.obs + n * q.obs_ld + 4 : nullptr;
    const int head = bulk ? (int)(((16u - (uint32_t)((uintptr_t)grow & 15u)) & 15u) >> 2) : 0;      // floats before the body
    float* const srow = bulk ? stage + ((4 - head) & 3) : nullptr;                                    // srow + head is 16-byte aligned
    // Fast path (the production call: distances + observation columns only, no ray of the tile saw a hit beyond the miss sentinel):
    // every load of the thread's <= 8 rays is issued before the first store -- in the general loop below a ray's column indices,
    // its stores and the next ray's loads form one dependent chain, and the CTA sits in this tail at memory latency (the obs
    // output cost 4.6 % of the kernel)
    {
        bool any_far = false;
        for (int i = tid; i < (np + 31) / 32; i += nthreads) any_far |= far[i] != 0u;
        any_far = __syncthreads_or(any_far ? 1 : 0) != 0;
        if (!any_far && !want_geo && !q.hit_slot && np <= 8 * nthreads && !bulk && (q.spec_slot & 16)) {
            uint32_t key[8];
            int ca[8], cb[8];
            const bool want_cols = q.obs != nullptr || q.obs16 != nullptr;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int p = tid + u * nthreads;
                key[u] = p < np ? res[p] : 0u;
                ca[u] = (want_cols && p < np) ? __ldg(q.col_a + p0 + p) : -1;
                cb[u] = (want_cols && p < np) ? __ldg(q.col_b + p0 + p) : -1;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int p = tid + u * nthreads;
                if (p < np) {
                    const __half d = h_from_bits(key_bits(key[u]));
                    q.dist[n * q.P + p0 + p] = d;
                    const __half hv = h_mul(d, __float2half_rn(0.5f));                       // fp16(dist / 2)
                    if (q.obs) {
                        const float v = __half2float(hv);
                        if (ca[u] >= 0) q.obs[n * q.obs_ld + ca[u]] = v;
                        if (cb[u] >= 0) q.obs[n * q.obs_ld + cb[u]] = v;
                    }
                    if (q.obs16) {
                        if (ca[u] >= 0) q.obs16[n * q.obs16_ld + (ca[u] - q.obs16_col0)] = hv;
                        if (cb[u] >= 0) q.obs16[n * q.obs16_ld + (cb[u] - q.obs16_col0)] = hv;
                    }
                }
            }
            return;
        }
    }
    for (int p = tid; p < np; p += nthreads) {
        uint32_t key = res[p];
        const bool far_hit = (far[p >> 5] >> (p & 31)) & 1u;
        if (far_hit && (key >> 16) == ORD_MISS) {
            const double* pp = q.pattern + (int64_t)(p0 + p) * 3;
            double xo, yo, zo;
            body_transform<double>(pp[0], pp[1], pp[2], tr, tx, ty, tz, xo, yo, zo);
            const H3 s = {h_from_double(xo), h_from_double(yo), h_from_double(zo)};
            const int cx = cell_coord(s.x, q.shift_x, q.res, q.inv_res, q.G0 - 1, q.sem);
            const int cy = min(cell_coord(s.y, q.shift_y, q.res, q.inv_res, q.G0 - 1, q.sem), q.G1 - 1);
            if (q.index) {
                key = literal_ray(q.index + ((int64_t)cx * q.G1 + cy) * q.Ks, q.K, q.recs, s, dlit);
            } else {
                const uint32_t bi = (uint32_t)(cx / RVB_BLK) * (uint32_t)q.nBy + (uint32_t)(cy / RVB_BLK);
                const uint32_t o0 = __ldg(q.blk_off + bi), o1 = __ldg(q.blk_off + bi + 1);
                key = literal_ray_blk(q.blk_ids + o0, q.blk_slots + o0, (int)(o1 - o0), (cx % RVB_BLK) * RVB_BLK + cy % RVB_BLK, q.recs, s, dlit);
            }
        }
        const unsigned short kb = key_bits(key);
        const int slot = (int)((key >> 1) & 0x7fffu);
        const int64_t o = n * q.P + p0 + p;
        q.dist[o] = h_from_bits(kb);
        if (q.hit_slot) q.hit_slot[o] = slot;
        if (want_geo) {
            const double* pp = q.pattern + (int64_t)(p0 + p) * 3;
            double xo, yo, zo;
            body_transform<double>(pp[0], pp[1], pp[2], tr, tx, ty, tz, xo, yo, zo);
            const __half hx = h_from_double(xo), hy = h_from_double(yo), hz = h_from_double(zo);
            if (q.sources) { q.sources[o * 3 + 0] = hx; q.sources[o * 3 + 1] = hy; q.sources[o * 3 + 2] = hz; }
            if (q.hit_tri) {
                const int cx = cell_coord(hx, q.shift_x, q.res, q.inv_res, q.G0 - 1, q.sem);
                const int cy = min(cell_coord(hy, q.shift_y, q.res, q.inv_res, q.G0 - 1, q.sem), q.G1 - 1);
                if (q.index) {
                    q.hit_tri[o] = __ldg(q.index + ((int64_t)cx * q.G1 + cy) * q.Ks + slot);
                } else {
                    const uint32_t bi = (uint32_t)(cx / RVB_BLK) * (uint32_t)q.nBy + (uint32_t)(cy / RVB_BLK);
                    const uint32_t o0 = __ldg(q.blk_off + bi), o1 = __ldg(q.blk_off + bi + 1);
                    q.hit_tri[o] = tri_of_slot_blk(q.blk_ids + o0, q.blk_slots + o0, (int)(o1 - o0), (cx % RVB_BLK) * RVB_BLK + cy % RVB_BLK, slot);
                }
            }
            if (q.pt) {
                const __half k = h_from_bits(kb);
                q.pt[o * 3 + 0] = h_sub(hx, h_mul(__low2half(dx2), k));     // ray_casting.py:63
                q.pt[o * 3 + 1] = h_sub(hy, h_mul(__low2half(dy2), k));
                q.pt[o * 3 + 2] = h_sub(hz, h_mul(__low2half(dz2), k));
            }
        }
        if (q.obs) {
            const float v = __half2float(h_mul(h_from_bits(kb), __float2half_rn(0.5f)));     // fp16(dist / 2) -> f32
            const int ca = q.col_a[p0 + p], cb = q.col_b[p0 + p];
            if (bulk) {
                if (ca >= 4) srow[ca - 4] = v;
                if (cb >= 4) srow[cb - 4] = v;
            } else {
                if (ca >= 0) q.obs[n * q.obs_ld + ca] = v;
                if (cb >= 0) q.obs[n * q.obs_ld + cb] = v;
            }
        }
        if (q.obs16) {                                                                        // the same value, kept in fp16
            const __half v = h_mul(h_from_bits(kb), __float2half_rn(0.5f));
            const int ca = q.col_a[p0 + p], cb = q.col_b[p0 + p];
            if (ca >= 0) q.obs16[n * q.obs16_ld + (ca - q.obs16_col0)] = v;
            if (cb >= 0) q.obs16[n * q.obs16_ld + (cb - q.obs16_col0)] = v;
        }
    }
    if (bulk) {
        __syncthreads();
        const int body = ((ncols - head) >> 2) << 2;                  // floats in the 16-byte aligned middle part
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(srow + head);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(grow + head), "r"(saddr), "r"(body * 4) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (tid < head) grow[tid] = srow[tid];
        if (tid >= 32 && tid - 32 < ncols - head - body) grow[head + body + tid - 32] = srow[head + body + tid - 32];
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the source must outlive the copy
    }
}

}  // namespace rc

