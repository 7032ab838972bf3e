/* TEST INFRASTRUCTURE ONLY -- never linked, loaded or called by the product (only tests/ and __graft_entry__.build() touch it).
 *
 * Plain-C restatement of the reference's ray / triangle test, independent of torch:
 *   ray_distance            utils/camera/ray_casting.py:3-66   (n rays against n triangles, fp16)
 *   min over the K candidates of a ray, value + first index    utils/camera/camera.py:116-117 (torch.min(dim=2))
 *   Camera._depth_transform + _height_lookup for one env         utils/camera/camera.py:165-264 (given sin/cos of the negated
 *                                                                euler angles, so that no libm enters the comparison)
 * Arithmetic model (what ATen does for Half tensors on either device): every elementwise op converts its operands to
 * fp32, computes, and rounds the result once to fp16; comparisons are exact; F.normalize = v / max(||v||, eps) with the
 * norm accumulated in fp32 and rounded to fp16, eps = 1e-12 -> 0 in fp16.  _Float16 <-> float conversions of gcc are IEEE
 * round-to-nearest-even; build with -ffp-contract=off so that no multiply-add is fused.
 * Pinned by tests/test_c_oracle_cpu.py against the golden vectors the reference itself produced (tests/golden/rover_golden.pt).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef _Float16 h16;

static inline h16 from_bits(uint16_t b) { h16 h; memcpy(&h, &b, 2); return h; }
static inline uint16_t to_bits(h16 h) { uint16_t b; memcpy(&b, &h, 2); return b; }
static inline h16 hadd(h16 a, h16 b) { return (h16)((float)a + (float)b); }
static inline h16 hsub(h16 a, h16 b) { return (h16)((float)a - (float)b); }
static inline h16 hmul(h16 a, h16 b) { return (h16)((float)a * (float)b); }
static inline h16 hdiv(h16 a, h16 b) { return (h16)((float)a / (float)b); }

typedef struct { h16 x, y, z; } v3;

static inline v3 load3(const uint16_t* p) { v3 v = {from_bits(p[0]), from_bits(p[1]), from_bits(p[2])}; return v; }
static inline v3 sub3(v3 u, v3 v) { v3 r = {hsub(u.x, v.x), hsub(u.y, v.y), hsub(u.z, v.z)}; return r; }
/* Tensor.cross (ray_casting.py:40,44,49,54): (u1 v2 - u2 v1, u2 v0 - u0 v2, u0 v1 - u1 v0), every op rounded */
static inline v3 cross3(v3 u, v3 v) {
    v3 r = {hsub(hmul(u.y, v.z), hmul(u.z, v.y)), hsub(hmul(u.z, v.x), hmul(u.x, v.z)), hsub(hmul(u.x, v.y), hmul(u.y, v.x))};
    return r;
}
/* x0*d0 + x1*d1 + x2*d2, left to right (ray_casting.py:41) */
static inline h16 dot3(v3 u, v3 v) { return hadd(hadd(hmul(u.x, v.x), hmul(u.y, v.y)), hmul(u.z, v.z)); }

/* One ray against one triangle.  tri = 3 vertices x 3 coordinates.  Returns k_after_check; pt (3 halves) optional. */
static h16 ray_triangle(v3 s, v3 dir, const uint16_t* tri, uint16_t* pt) {
    const h16 zeros = (h16)(0.0f - 0.1f);             /* ray_casting.py:26  torch.zeros(half) - epsilon */
    const h16 ones = (h16)(1.0f + 0.1f);              /* ray_casting.py:27 */
    const h16 err = (h16)((float)ones * 10.0f);       /* ray_casting.py:28 */
    /* ray_casting.py:31  d = -normalize(directions) */
    const float fx = (float)dir.x, fy = (float)dir.y, fz = (float)dir.z;
    h16 nrm = (h16)sqrtf(fx * fx + fy * fy + fz * fz);
    const h16 eps = (h16)1e-12f;
    h16 den = (nrm != nrm) ? nrm : ((float)nrm > (float)eps ? nrm : eps);     /* clamp_min, NaN propagates */
    v3 d = {(h16)(-(float)hdiv(dir.x, den)), (h16)(-(float)hdiv(dir.y, den)), (h16)(-(float)hdiv(dir.z, den))};
    v3 a = load3(tri + 6);                            /* ray_casting.py:34  a = triangles[:,2] */
    v3 b = sub3(load3(tri + 3), a);                   /* :35 */
    v3 c = sub3(load3(tri + 0), a);                   /* :36 */
    v3 g = sub3(s, a);                                /* :37 */
    v3 bc = cross3(b, c);
    h16 det = dot3(bc, d);                            /* :40-41 */
    h16 n = hdiv(dot3(cross3(g, c), d), det);         /* :44-45 */
    if (det == zeros) n = err;                        /* :46 */
    h16 m = hdiv(dot3(cross3(b, g), d), det);         /* :49-50 */
    if (det == ones) m = err;                         /* :51 */
    h16 k = hdiv(dot3(bc, g), det);                   /* :54-55 */
    if (det == ones) k = err;                         /* :56 */
    h16 kk = (n >= zeros && m >= zeros && hadd(n, m) <= ones) ? k : err;      /* :59 */
    if (pt) {                                         /* :63  pt = sources - d * k */
        pt[0] = to_bits(hsub(s.x, hmul(d.x, kk)));
        pt[1] = to_bits(hsub(s.y, hmul(d.y, kk)));
        pt[2] = to_bits(hsub(s.z, hmul(d.z, kk)));
    }
    return kk;
}

/* ray_distance(sources [n,3], directions [n,3], triangles [n,3,3]) -> k [n], pt [n,3]; all IEEE binary16 bit patterns */
void rvo_ray_distance(const uint16_t* sources, const uint16_t* directions, const uint16_t* triangles, int64_t n, uint16_t* k,
                      uint16_t* pt) {
    for (int64_t i = 0; i < n; ++i)
        k[i] = to_bits(ray_triangle(load3(sources + 3 * i), load3(directions + 3 * i), triangles + 9 * i, pt ? pt + 3 * i : 0));
}

/* One env's Camera.get_depths core (camera.py:84-120): ray r against the K candidates ids[r*K..] of its cell through the
 * reference's indirection vertices[triangles[id]]; torch.min(dim) = smallest value, first index on ties, NaN wins. */
void rvo_cast_min(const uint16_t* sources, const uint16_t* directions, int64_t n_rays, const int32_t* ids, int64_t K,
                  const int32_t* triangles, const uint16_t* vertices, uint16_t* dist, int32_t* slot) {
    for (int64_t r = 0; r < n_rays; ++r) {
        v3 s = load3(sources + 3 * r), dir = load3(directions + 3 * r);
        h16 best = 0;
        int32_t best_j = -1;
        for (int64_t j = 0; j < K; ++j) {
            const int32_t* t = triangles + 3 * (int64_t)ids[r * K + j];
            uint16_t tri[9];
            for (int v = 0; v < 3; ++v) memcpy(tri + 3 * v, vertices + 3 * (int64_t)t[v], 6);
            h16 kk = ray_triangle(s, dir, tri, 0);
            if (best_j < 0 || (best == best && (kk != kk || kk < best))) { best = kk; best_j = (int32_t)j; }
        }
        dist[r] = to_bits(best);
        slot[r] = best_j;
    }
}

/* Camera.get_depths for ONE env (camera.py:60-145), torch-CPU semantics, given trig = (sx, cx, sy, cy, sz, cz) =
 * sin/cos of the NEGATED roll, pitch, yaw in fp32 (camera.py:184-189).
 *   _depth_transform (camera.py:165-212): pattern f64 [P,3] (+ the extra point (0,0,-1), :179-181); fp32 inputs promote to
 *   fp64 (f64 [1,P] x f32 [N,1] -> f64); x' = tx + sz*(y*cx + z*sx) + cz*(x*cy - sy*(z*cx - y*sx)) etc. (:197-199);
 *   direction = transformed extra point - translation (:202-207); cast to fp16 through fp32 (:212).
 *   _height_lookup (camera.py:233-264): cell = round_half_even(clamp((xy16 - shift) / res, 0, G-1)) in fp32 (f16 - f32 -> f32),
 *   candidates = map_indices[:, cx, cy] of the [K,G,G] asset (the reference indexes its permuted [G,G,K] view, :157-158).
 *   then ray_distance per (ray, candidate) and torch.min over K (:110-117). */
void rvo_get_depths_env(const float* pos, const float* trig, const double* pattern, int64_t P, const int32_t* map_kgg, int64_t G,
                        int64_t K, const int32_t* triangles, const uint16_t* vertices, float shift_x, float shift_y, float res,
                        uint16_t* sources, uint16_t* dist, int32_t* slot) {
    const double sx = trig[0], cx = trig[1], sy = trig[2], cy = trig[3], sz = trig[4], cz = trig[5];
    const double tx = pos[0], ty = pos[1], tz = pos[2];
    double out[3];
#define RVO_XFORM(x, y, z)                                                  \
    do {                                                                    \
        const double A_ = (y) * cx + (z) * sx, C_ = (z) * cx - (y) * sx;    \
        const double B_ = (x) * cy - sy * C_;                               \
        out[0] = (tx + sz * A_) + cz * B_;                                  \
        out[1] = (ty + cz * A_) - sz * B_;                                  \
        out[2] = (tz + (x) * sy) + cy * C_;                                 \
    } while (0)
    RVO_XFORM(0.0, 0.0, -1.0);
    uint16_t dir[3];
    for (int i = 0; i < 3; ++i) dir[i] = to_bits((h16)(float)(out[i] - (double)pos[i]));
    const v3 d = load3(dir);
    for (int64_t r = 0; r < P; ++r) {
        RVO_XFORM(pattern[3 * r], pattern[3 * r + 1], pattern[3 * r + 2]);
        uint16_t* s16 = sources + 3 * r;
        for (int i = 0; i < 3; ++i) s16[i] = to_bits((h16)(float)out[i]);
        int64_t cell[2];
        const float shift[2] = {shift_x, shift_y};
        for (int i = 0; i < 2; ++i) {
            float v = ((float)from_bits(s16[i]) - shift[i]) / res;
            v = fminf(fmaxf(v, 0.0f), (float)(G - 1));
            cell[i] = (int64_t)rintf(v);                   /* round half to even (default rounding mode) */
        }
        const v3 s = load3(s16);
        h16 best = 0;
        int32_t best_j = -1;
        for (int64_t j = 0; j < K; ++j) {
            const int32_t* t = triangles + 3 * (int64_t)map_kgg[(j * G + cell[0]) * G + cell[1]];
            uint16_t tri[9];
            for (int v = 0; v < 3; ++v) memcpy(tri + 3 * v, vertices + 3 * (int64_t)t[v], 6);
            h16 kk = ray_triangle(s, d, tri, 0);
            if (best_j < 0 || (best == best && (kk != kk || kk < best))) { best = kk; best_j = (int32_t)j; }
        }
        dist[r] = to_bits(best);
        slot[r] = best_j;
    }
#undef RVO_XFORM
}

/* Rock_Detection.get_collisions' cast (rock_detect.py:65-115) for ONE env given its 26 rays (24 wheel + 2 body; sources and
 * directions fp16 [R,3]): per ray the cell lookup of rock_detect.py:373-401 (same arithmetic as camera.py:233-264), the K-list of
 * that cell, ray_distance per candidate and torch.min over K. */
void rvo_cast_rays_env(const uint16_t* sources, const uint16_t* directions, int64_t R, const int32_t* map_kgg, int64_t G, int64_t K,
                       const int32_t* triangles, const uint16_t* vertices, float shift_x, float shift_y, float res, uint16_t* dist,
                       int32_t* slot) {
    const float shift[2] = {shift_x, shift_y};
    for (int64_t r = 0; r < R; ++r) {
        const uint16_t* s16 = sources + 3 * r;
        int64_t cell[2];
        for (int i = 0; i < 2; ++i) {
            float v = ((float)from_bits(s16[i]) - shift[i]) / res;
            v = fminf(fmaxf(v, 0.0f), (float)(G - 1));
            cell[i] = (int64_t)rintf(v);
        }
        const v3 s = load3(s16), d = load3(directions + 3 * r);
        h16 best = 0;
        int32_t best_j = -1;
        for (int64_t j = 0; j < K; ++j) {
            const int32_t* t = triangles + 3 * (int64_t)map_kgg[(j * G + cell[0]) * G + cell[1]];
            uint16_t tri[9];
            for (int v = 0; v < 3; ++v) memcpy(tri + 3 * v, vertices + 3 * (int64_t)t[v], 6);
            h16 kk = ray_triangle(s, d, tri, 0);
            if (best_j < 0 || (best == best && (kk != kk || kk < best))) { best = kk; best_j = (int32_t)j; }
        }
        dist[r] = to_bits(best);
        slot[r] = best_j;
    }
}

/* RoverTask.check_collision (rover.py:663-668), torch-CPU semantics: min over the 24 wheel / 2 body distances (NaN wins), then
 * |min| < 0.8 resp. 0.45 with the Python scalar rounded to fp16 first (the comparison runs in the tensor's dtype on the CPU). */
int64_t rvo_check_collision(const uint16_t* wheel, int64_t n_wheel, const uint16_t* body, int64_t n_body) {
    h16 w = from_bits(wheel[0]), b = from_bits(body[0]);
    for (int64_t i = 1; i < n_wheel; ++i) { h16 v = from_bits(wheel[i]); if (w == w && (v != v || v < w)) w = v; }
    for (int64_t i = 1; i < n_body; ++i) { h16 v = from_bits(body[i]); if (b == b && (v != v || v < b)) b = v; }
    const h16 aw = (h16)fabsf((float)w), ab = (h16)fabsf((float)b);
    int64_t col = aw < (h16)0.8f ? 1 : 0;
    if (ab < (h16)0.45f) col = 1;
    return col;
}

/* Rock_Detection._get_wheel_rays + _get_body_rays (rock_detect.py:160-371) for ONE env, fp32 -> fp16.
 * trig = sin/cos of the negated body euler angles (given, as in rvo_get_depths_env); the sin/cos of the joint angles are
 * libm's here (Sleef's in torch): the rays equal the reference's to an fp32 ulp before the fp16 cast, i.e. bit for bit for
 * almost all of them.  joints = the 13 DOF positions in the order of rock_detect.py:174-188.
 * Output: 24 wheel rays (6 wheels x 4 points, :193-198,317) then 2 body rays (:326,338-340): sources, dirs fp16 [26,3]. */
static void body_xform_f(float x, float y, float z, const float* t6, float tx, float ty, float tz, float* o) {
    const float sx = t6[0], cx = t6[1], sy = t6[2], cy = t6[3], sz = t6[4], cz = t6[5];
    const float A = y * cx + z * sx, C = z * cx - y * sx, B = x * cy - sy * C;       /* :305-307 / :356-358 */
    o[0] = (tx + sz * A) + cz * B;
    o[1] = (ty + cz * A) - sz * B;
    o[2] = (tz + x * sy) + cy * C;
}

void rvo_rock_rays_env(const float* pos, const float* trig, const float* joints, uint16_t* sources, uint16_t* dirs) {
    static const float RAYS[5][3] = {{(float)(0.215 / 2), (float)(0.130 / 2), 0.1f}, {(float)(0.215 / 2), (float)(-0.130 / 2), 0.1f},
                                     {(float)(-0.215 / 2), (float)(0.130 / 2), 0.1f}, {(float)(-0.215 / 2), (float)(-0.130 / 2), 0.1f},
                                     {0.0f, 0.0f, -1.0f}};                                        /* :193-197 */
    static const float POS0[6][3] = {{0.286f, 0.385f, -0.197f}, {0.286f, -0.385f, -0.197f}, {-0.146f, 0.447f, -0.197f},
                                     {-0.146f, -0.447f, -0.197f}, {-0.440f, 0.385f, -0.197f}, {-0.440f, -0.385f, -0.197f}};   /* :201-206 */
    static const float POS1[6][3] = {{0.153f, 0.0f, 0.03f}, {0.153f, 0.0f, 0.03f}, {0.153f, 0.0f, 0.03f}, {0.153f, -0.0f, 0.03f},
                                     {0.0f, 0.0f, 0.03f}, {0.0f, 0.0f, 0.03f}};                  /* :210-215 */
    const float* j = joints;
    const float steer[6] = {j[4], j[6], 0.0f, 0.0f, -j[7], j[8]};                                 /* :248 */
    const float susy[6] = {-j[0], j[1], -j[0], j[1], 0.0f, 0.0f};                                 /* :263 */
    const float susx[6] = {0.0f, 0.0f, 0.0f, 0.0f, -j[2], -j[2]};                                 /* :264 */
    for (int w = 0; w < 6; ++w) {
        const float ss = sinf(-steer[w]), cs = cosf(-steer[w]);
        const float sux = sinf(susx[w]), cux = cosf(susx[w]), suy = sinf(susy[w]), cuy = cosf(susy[w]);
        float dir[3] = {0, 0, 0};
        float pts[4][3];
        for (int p = 4; p >= 0; --p) {                       /* entry 4 = the direction: zero translation at every stage */
            const int isdir = (p == 4);
            const float x = RAYS[p][0], y = RAYS[p][1], z = RAYS[p][2];
            const float t0x = isdir ? 0.0f : POS0[w][0], t0y = isdir ? 0.0f : POS0[w][1], t0z = isdir ? 0.0f : POS0[w][2];
            const float t1x = isdir ? 0.0f : POS1[w][0], t1y = isdir ? 0.0f : POS1[w][1], t1z = isdir ? 0.0f : POS1[w][2];
            const float x1 = (t0x + x * cs) + y * ss;                                             /* :256-258 */
            const float y1 = (t0y + y * cs) - x * ss;
            const float z1 = t0z + z;
            const float Cq = z1 * cux - y1 * sux;
            const float x2 = (t1x + x1 * cuy) - suy * Cq;                                         /* :275-277 */
            const float y2 = (t1y + y1 * cux) + z1 * sux;
            const float z2 = (t1z + x1 * suy) + cuy * Cq;
            float o[3];
            body_xform_f(x2, y2, z2, trig, isdir ? 0.0f : pos[0], isdir ? 0.0f : pos[1], isdir ? 0.0f : pos[2], o);
            if (isdir) { dir[0] = o[0]; dir[1] = o[1]; dir[2] = o[2]; }
            else { pts[p][0] = o[0]; pts[p][1] = o[1]; pts[p][2] = o[2]; }
        }
        for (int p = 0; p < 4; ++p)
            for (int i = 0; i < 3; ++i) {
                sources[3 * (4 * w + p) + i] = to_bits((h16)pts[p][i]);                           /* :317 */
                dirs[3 * (4 * w + p) + i] = to_bits((h16)dir[i]);                                 /* :314 */
            }
    }
    static const float BODY[3][3] = {{0.340f, 0.0f, -0.01f}, {-0.485f, 0.0f, -0.01f}, {0.0f, 1.0f, 0.0f}};   /* :326,338 */
    float o[3][3];
    for (int p = 0; p < 3; ++p) body_xform_f(BODY[p][0], BODY[p][1], BODY[p][2], trig, pos[0], pos[1], pos[2], o[p]);
    for (int p = 0; p < 2; ++p)
        for (int i = 0; i < 3; ++i) {
            sources[3 * (24 + p) + i] = to_bits((h16)o[p][i]);
            dirs[3 * (24 + p) + i] = to_bits((h16)(o[2][i] - pos[i]));                            /* :340 */
        }
}
