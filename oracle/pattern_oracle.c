/* TEST INFRASTRUCTURE ONLY -- never linked, loaded or called by the product.
 *
 * Plain-C restatement of the reference's heightmap point pattern, tasks/utils/camera/heightmap_distribution.py:11-204:
 * constants :16-30, coarse fan :45-59, fine box with exact-value de-duplication :62-78, np.round(..., 4) :100, x <-> y swap :105,
 * index vectors :107-109, the half-plane tests :153-193 (with the reference's quirk: for slanted lines the 'left' test is the
 * same '<' as the 'right' one).  All arithmetic is fp64 and accumulated exactly like the reference's Python loops.
 * Pinned by tests/test_c_oracle_cpu.py against the pattern the reference's Heightmap class produced (golden ref_pattern,
 * ref_coarse_idx, ref_fine_idx; 634 / 1112 points as teacher_loader.py:47-48 states).
 */
#include <math.h>
#include <stdint.h>

enum { OVER, BELOW, LEFT, RIGHT };
typedef struct { double p0[2], p1[2]; int side; } line_t;

static int line_test(double x, double y, const line_t* lines, int n) {
    int ok = 1;
    for (int i = 0; i < n; ++i) {
        const line_t* L = &lines[i];
        const double dx = L->p0[0] - L->p1[0], dy = L->p0[1] - L->p1[1];
        if (dx == 0) {                                            /* vertical line (:160-167) */
            if (x < L->p0[0] && L->side == RIGHT) ok = 0;
            if (x > L->p0[0] && L->side == LEFT) ok = 0;
            continue;
        }
        const double a = dy / dx, b = L->p0[1] - a * L->p0[0];
        if (a == 0) {                                             /* horizontal line (:172-179) */
            if (y > b && L->side == BELOW) ok = 0;
            if (y < b && L->side == OVER) ok = 0;
            continue;
        }
        if (y < a * x + b && L->side == OVER) ok = 0;             /* :181-192 */
        if (y > a * x + b && L->side == BELOW) ok = 0;
        if (x < (y - b) / a && (L->side == RIGHT || L->side == LEFT)) ok = 0;
    }
    return ok;
}

/* points f64 [cap,3] (rover frame after the swap), coarse_idx / fine_idx i64 [cap]; returns the number of points, the two
 * index counts through n_coarse / n_fine; -1 if cap is too small. */
int64_t rvo_heightmap_pattern(double* points, int64_t cap, int64_t* coarse_idx, int64_t* n_coarse, int64_t* fine_idx, int64_t* n_fine) {
    const line_t coarse[3] = {{{1.220, 0.118}, {4.4455, 3.150}, OVER}, {{-1.220, 0.118}, {-4.4455, 3.150}, OVER},
                              {{1.220, 0.118}, {-1.220, 0.118}, OVER}};                              /* :16-17 */
    const line_t fine[4] = {{{1.0, 0.118}, {1.0, 0.119}, LEFT}, {{-1.0, 0.118}, {-1.0, 0.119}, RIGHT},
                            {{1.0, 0.118}, {-1.0, 0.118}, OVER}, {{1.0, 1.400}, {-1.0, 1.400}, BELOW}};   /* :19 */
    const double z = -0.26878;                                                                         /* :30 */
    int64_t n = 0;
    for (double y = -10; y < 10; y += 0.15)                                                            /* :45-55 */
        for (double x = -10; x < 10;) {
            x += 0.15;
            if (line_test(x, y, coarse, 3) && sqrt(x * x + y * y) < 3.5) {
                if (n >= cap) return -1;
                points[3 * n] = x; points[3 * n + 1] = y; points[3 * n + 2] = z; ++n;
            }
        }
    *n_coarse = n;
    for (int64_t i = 0; i < n; ++i) coarse_idx[i] = i;                                                 /* :57-59 */
    for (double y = -10; y < 10; y += 0.05)                                                            /* :62-74 */
        for (double x = -10; x < 10;) {
            x += 0.05;
            if (!line_test(x, y, fine, 4)) continue;
            int seen = 0;                                                                              /* :71 exact list membership */
            for (int64_t i = 0; i < n && !seen; ++i) seen = (points[3 * i] == x && points[3 * i + 1] == y);
            if (seen) continue;
            if (n >= cap) return -1;
            points[3 * n] = x; points[3 * n + 1] = y; points[3 * n + 2] = z; ++n;
        }
    int64_t nf = 0;
    for (int64_t i = 0; i < n; ++i)                                                                    /* :107-109 */
        if (line_test(points[3 * i], points[3 * i + 1], fine, 4)) fine_idx[nf++] = i;
    *n_fine = nf;
    for (int64_t i = 0; i < n; ++i) {                                                                  /* :100 np.round(.., 4); :105 swap */
        const double x = rint(points[3 * i] * 1e4) / 1e4, y = rint(points[3 * i + 1] * 1e4) / 1e4;
        points[3 * i] = y; points[3 * i + 1] = x; points[3 * i + 2] = rint(points[3 * i + 2] * 1e4) / 1e4;
    }
    return n;
}
