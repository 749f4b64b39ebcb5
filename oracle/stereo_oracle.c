/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (stable-diffusion-webui-depthmap-script_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it.
 *
 * Plain-C, single-threaded-per-row restatement of the reference's stereo warp + gap fill:
 *   /root/reference/src/stereoimage_generation.py
 *     apply_stereo_divergence            :77-92   (depth normalisation, px conversion)
 *     apply_stereo_divergence_naive      :95-159  (forward scatter + none / naive / naive_interpolating fill)
 *     apply_stereo_divergence_polylines  :162-283 (polyline morph, stable insertion sort, sweep rasteriser)
 *     overlap_red_cyan                   :286-307
 * The reference runs these bodies through numba (LLVM, fp64, no fast-math, no FMA contraction); compile this file
 * with -O2 -ffp-contract=off so gcc does the same.  Pinned against the reference itself by tests/test_oracle_pin.py
 * (run in the build container, where /root/reference is importable) and by tests/golden/ fixtures everywhere else.
 *
 * Cast semantics pinned to numba 0.65 / x86-64:
 *   int(float64)          -> truncation toward zero to int64 (NaN / out-of-range -> INT64_MIN, as cvttsd2si does)
 *   float64 -> uint8      -> truncate toward zero to a signed integer, then wrap modulo 256
 *   uint8 + uint8         -> wraps modulo 256
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline int64_t py_int(double v) {
    if (!(v > -9.2e18 && v < 9.2e18)) return INT64_MIN; /* NaN or out of range: x86 cvttsd2si "indefinite" */
    return (int64_t)v;
}

static inline uint8_t f64_to_u8(double v) {
    if (!(v > -2.0e9 && v < 2.0e9)) return 0;
    return (uint8_t)(int32_t)v;
}

/* stereoimage_generation.py:79-81 — (depth - min) / (max - min) on uint16 input promotes to float64 true-divide. */
void oracle_normalize_depth_u16(const uint16_t *depth, int64_t n, double *nd) {
    uint16_t mn = 65535, mx = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (depth[i] < mn) mn = depth[i];
        if (depth[i] > mx) mx = depth[i];
    }
    double den = (double)(uint16_t)(mx - mn);
    for (int64_t i = 0; i < n; ++i) nd[i] = (double)(uint16_t)(depth[i] - mn) / den; /* 0/0 -> NaN like numpy */
}

/* stereoimage_generation.py:95-159.  fill: 0 = none, 1 = naive, 2 = naive_interpolating.  c must be 3. */
int oracle_stereo_naive(const uint8_t *img, const double *nd, int h, int w, double div_px, double sep_px,
                        double expo, int fill, uint8_t *out) {
    const int c = 3;
    uint8_t *derived = out;
    memset(derived, 0, (size_t)h * w * c);
    uint8_t *filled = (uint8_t *)calloc((size_t)h * w, 1);
    if (!filled) return -1;

    for (int row = 0; row < h; ++row) {
        /* :107 sweep order so that nearer pixels overwrite farther ones at their destination */
        int start = div_px < 0 ? 0 : w - 1, stop = div_px < 0 ? w : -1, step = div_px < 0 ? 1 : -1;
        for (int col = start; col != stop; col += step) {
            double v = pow(nd[(size_t)row * w + col], expo) * div_px + sep_px;
            int64_t col_d = (int64_t)col + py_int(v);
            if (py_int(v) == INT64_MIN) continue;
            if (0 <= col_d && col_d < w) {
                memcpy(derived + ((size_t)row * w + col_d) * c, img + ((size_t)row * w + col) * c, c);
                filled[(size_t)row * w + col_d] = 1;
            }
        }
    }

    if (fill == 2) { /* :114-141 */
        for (int row = 0; row < h; ++row) {
            uint8_t *r = derived + (size_t)row * w * c;
            const uint8_t *f = filled + (size_t)row * w;
            for (int l = 0; l < w; ++l) {
                if ((int)r[l * c] + r[l * c + 1] + r[l * c + 2] != 0 || f[l]) continue;
                uint8_t lb[3] = {0, 0, 0}, rb[3] = {0, 0, 0};
                if (l > 0) memcpy(lb, r + (l - 1) * c, c);
                int rp = l + 1;
                while (rp < w) {
                    if ((int)r[rp * c] + r[rp * c + 1] + r[rp * c + 2] != 0 && f[rp]) {
                        memcpy(rb, r + rp * c, c);
                        break;
                    }
                    ++rp;
                }
                if ((int)lb[0] + lb[1] + lb[2] == 0) memcpy(lb, rb, c);
                else if ((int)rb[0] + rb[1] + rb[2] == 0) memcpy(rb, lb, c);
                int total_steps = 1 + rp - l;
                double stepv[3];
                for (int k = 0; k < c; ++k) stepv[k] = ((double)rb[k] - (double)lb[k]) / (double)total_steps;
                for (int col = l; col < rp; ++col)
                    for (int k = 0; k < c; ++k)
                        r[col * c + k] = (uint8_t)(lb[k] + f64_to_u8(stepv[k] * (double)(col - l + 1)));
            }
        }
    } else if (fill == 1) { /* :142-157 — reads the pre-fill image, writes a copy */
        uint8_t *fix = (uint8_t *)malloc((size_t)h * w * c);
        if (!fix) { free(filled); return -1; }
        memcpy(fix, derived, (size_t)h * w * c);
        int64_t lim = py_int(div_px);
        if (lim < 0) lim = -lim;
        for (int row = 0; row < h; ++row) {
            const uint8_t *f = filled + (size_t)row * w;
            for (int col = 0; col < w; ++col) {
                if (f[col]) continue;
                for (int64_t off = 1; off < lim + 2; ++off) {
                    int64_t ro = col + off, lo = col - off;
                    if (ro < w && f[ro]) { memcpy(fix + ((size_t)row * w + col) * c, derived + ((size_t)row * w + ro) * c, c); break; }
                    if (0 <= lo && f[lo]) { memcpy(fix + ((size_t)row * w + col) * c, derived + ((size_t)row * w + lo) * c, c); break; }
                }
            }
        }
        memcpy(derived, fix, (size_t)h * w * c);
        free(fix);
    }
    free(filled);
    return 0;
}

/* stereoimage_generation.py:162-283.  One call processes rows [row0, row1) so callers can thread over rows. */
static int polylines_rows(const uint8_t *img, const double *nd, int w, int row0, int row1, double div_px,
                          double sep_px, double expo, int sharp, uint8_t *out) {
    const int c = 3;
    const double EPSILON = 1e-7;
    const double PIXEL_HALF_WIDTH = sharp ? 0.45 : 0.0;
    const int npt_cap = 5 + 2 * w;
    int64_t csg_cap = 5 * llabs(py_int(fabs(div_px))) + 25;
    double(*pt)[3] = malloc(sizeof(double[3]) * (size_t)(npt_cap + 1));
    double(*sg)[6] = malloc(sizeof(double[6]) * (size_t)npt_cap);
    /* the reference never bounds-checks csg (:223); give the oracle slack so pathological inputs do not crash */
    int64_t csg_alloc = csg_cap + 4 * (int64_t)w + 64;
    double(*csg)[6] = malloc(sizeof(double[6]) * (size_t)csg_alloc);
    if (!pt || !sg || !csg) { free(pt); free(sg); free(csg); return -1; }

    for (int row = row0; row < row1; ++row) {
        const uint8_t *irow = img + (size_t)row * w * c;
        const double *drow = nd + (size_t)row * w;
        memset(pt, 0, sizeof(double[3]) * (size_t)(npt_cap + 1));
        /* :177-192 vertices: (morphed x, closeness, source column) */
        int pt_end = 0;
        pt[pt_end][0] = -1.0 * w; pt[pt_end][1] = 0.0; pt[pt_end][2] = 0.0; ++pt_end;
        for (int col = 0; col < w; ++col) {
            double coord_d = pow(drow[col], expo) * div_px;
            double coord_x = (double)col + 0.5 + coord_d + sep_px;
            if (PIXEL_HALF_WIDTH < EPSILON) {
                pt[pt_end][0] = coord_x; pt[pt_end][1] = fabs(coord_d); pt[pt_end][2] = col; ++pt_end;
            } else {
                pt[pt_end][0] = coord_x - PIXEL_HALF_WIDTH; pt[pt_end][1] = fabs(coord_d); pt[pt_end][2] = col;
                pt[pt_end + 1][0] = coord_x + PIXEL_HALF_WIDTH; pt[pt_end + 1][1] = fabs(coord_d); pt[pt_end + 1][2] = col;
                pt_end += 2;
            }
        }
        pt[pt_end][0] = 2.0 * w; pt[pt_end][1] = 0.0; pt[pt_end][2] = w - 1; ++pt_end;

        /* :196-199 segments between consecutive original-order vertices */
        int sg_end = pt_end - 1;
        for (int i = 0; i < sg_end; ++i) {
            memcpy(sg[i], pt[i], sizeof(double[3]));
            memcpy(sg[i] + 3, pt[i + 1], sizeof(double[3]));
        }

        /* :214-219 stable insertion sort of pt[0:sg_end] by x, same swaps applied to sg.
         * (The reference evaluates pt[u][0] > pt[u+1][0] before 0 <= u; with u == -1 numba wraps to the last,
         *  all-zero row of pt, whose outcome is then discarded by the `and`.) */
        for (int i = 1; i < sg_end; ++i) {
            int u = i - 1;
            while (u >= 0 && pt[u][0] > pt[u + 1][0]) {
                double t3[3], t6[6];
                memcpy(t3, pt[u], sizeof t3); memcpy(pt[u], pt[u + 1], sizeof t3); memcpy(pt[u + 1], t3, sizeof t3);
                memcpy(t6, sg[u], sizeof t6); memcpy(sg[u], sg[u + 1], sizeof t6); memcpy(sg[u + 1], t6, sizeof t6);
                --u;
            }
        }

        /* :223-281 sweep rasteriser */
        memset(csg, 0, sizeof(double[6]) * (size_t)csg_alloc);
        int64_t csg_end = 0;
        int sg_pointer = 0, pt_i = 0;
        for (int col = 0; col < w; ++col) {
            double color[3] = {0.5, 0.5, 0.5};
            while (pt[pt_i][0] < col) ++pt_i;
            --pt_i;
            while (pt[pt_i][0] < col + 1) {
                double a = pt[pt_i][0], b = pt[pt_i + 1][0];
                /* Python max(col, x) == (x > col ? x : col), min(col+1, x) == (x < col+1 ? x : col+1): NaN x -> the integer */
                double coord_from = (a > (double)col ? a : (double)col) + EPSILON;
                double coord_to = (b < (double)(col + 1) ? b : (double)(col + 1)) - EPSILON;
                double significance = coord_to - coord_from;
                double coord_center = coord_from + 0.5 * significance;

                while (sg_pointer < sg_end && sg[sg_pointer][0] < coord_center) {
                    if (csg_end >= csg_alloc) { free(pt); free(sg); free(csg); return -2; }
                    memcpy(csg[csg_end], sg[sg_pointer], sizeof(double[6]));
                    ++sg_pointer; ++csg_end;
                }
                int64_t ci = 0;
                while (ci < csg_end) {
                    if (csg[ci][3] < coord_center) { memcpy(csg[ci], csg[csg_end - 1], sizeof(double[6])); --csg_end; }
                    else ++ci;
                }
                int64_t best = 0;
                if (csg_end != 1) {
                    double best_closeness = -EPSILON;
                    for (ci = 0; ci < csg_end; ++ci) {
                        double ip_k = (coord_center - csg[ci][0]) / (csg[ci][3] - csg[ci][0]);
                        double closeness = (1.0 - ip_k) * csg[ci][1] + ip_k * csg[ci][4];
                        if (best_closeness < closeness && 0.0 < ip_k && ip_k < 1.0) { best_closeness = closeness; best = ci; }
                    }
                }
                int64_t col_l = py_int(csg[best][2] + EPSILON), col_r = py_int(csg[best][5] + EPSILON);
                if (col_l < 0) col_l = 0; if (col_l >= w) col_l = w - 1;   /* oracle-only memory safety */
                if (col_r < 0) col_r = 0; if (col_r >= w) col_r = w - 1;
                if (col_l == col_r) {
                    for (int k = 0; k < c; ++k) color[k] += (double)irow[col_l * c + k] * significance;
                } else {
                    double ip_k = (coord_center - csg[best][0]) / (csg[best][3] - csg[best][0]);
                    for (int k = 0; k < c; ++k)
                        color[k] += ((double)irow[col_l * c + k] * (1.0 - ip_k) + (double)irow[col_r * c + k] * ip_k) * significance;
                }
                ++pt_i;
            }
            for (int k = 0; k < c; ++k) out[((size_t)row * w + col) * c + k] = f64_to_u8(color[k]);
        }
    }
    free(pt); free(sg); free(csg);
    return 0;
}

int oracle_stereo_polylines(const uint8_t *img, const double *nd, int h, int w, double div_px, double sep_px,
                            double expo, int sharp, int nthreads, uint8_t *out) {
    int rc = 0;
    if (nthreads < 1) nthreads = 1;
    const int chunk = 16;
    const int nchunks = (h + chunk - 1) / chunk;
    /* the reference is @njit(parallel=True) with prange over rows (:174); rows are independent */
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int ch = 0; ch < nchunks; ++ch) {
        int r0 = ch * chunk, r1 = r0 + chunk < h ? r0 + chunk : h;
        int r = polylines_rows(img, nd, w, r0, r1, div_px, sep_px, expo, sharp, out);
        if (r) {
#pragma omp critical
            rc = r;
        }
    }
    return rc;
}

/* stereoimage_generation.py:286-307 */
void oracle_overlap_red_cyan(const uint8_t *im1, const uint8_t *im2, int h, int w, uint8_t *out) {
    for (size_t i = 0; i < (size_t)h * w; ++i) {
        out[i * 3 + 0] = im1[i * 3 + 0];
        out[i * 3 + 1] = im2[i * 3 + 1];
        out[i * 3 + 2] = im2[i * 3 + 2];
    }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
