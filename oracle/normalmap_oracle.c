/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/stereo_oracle.c header for the rules).
 *
 * Plain-C fp64 restatement of
 *   /root/reference/src/normalmap_generation.py:5-56   create_normalmap
 *   /root/reference/src/core.py:44-50                   convert_to_i16
 *   /root/reference/src/core.py:189-206                 per-image min/max normalisation (+ "Range" clip)
 * The reference delegates the gradients to OpenCV (cv2.Sobel / cv2.GaussianBlur, CV_64F, BORDER_REFLECT_101; third
 * party, opencv-python-headless 4.13.0.92 in this image, unpinned by the reference).  Restated here from OpenCV's
 * published algorithm: getSobelKernels (repeated [1 1] smoothing then [-1 1] differencing, unnormalised),
 * getGaussianKernel (exp(-x^2/(2 sigma^2)) normalised to sum 1, computed in double), separable row-then-column
 * filtering.  For u16 input without pre-blur every Sobel partial sum is an exact multiple of 2^-8 below 2^53, so the
 * summation order cannot matter for small kernels; large Sobel kernels (binomial taps up to C(30,15)) and Gaussian
 * blurs round, so the tap order follows OpenCV's CV_64F filter engine (see sep_filter).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * n - 2 - p;
    }
    return p;
}

/* OpenCV getSobelKernels for one axis: order 0 (smoothing) or 1 (derivative); ksize odd >= 1. Returns tap count. */
static int sobel_kernel(int ksize, int order, double *k) {
    if (ksize == 1) {
        if (order == 0) { k[0] = 1.0; return 1; }
        k[0] = -1.0; k[1] = 0.0; k[2] = 1.0; return 3;  /* OpenCV widens a ksize-1 derivative to 3 taps */
    }
    if (ksize == 3) {
        if (order == 0) { k[0] = 1; k[1] = 2; k[2] = 1; }
        else { k[0] = -1; k[1] = 0; k[2] = 1; }
        return 3;
    }
    int64_t *ker = (int64_t *)calloc((size_t)ksize + 1, sizeof(int64_t));
    ker[0] = 1;
    for (int i = 0; i < ksize - order - 1; ++i) {
        int64_t oldval = ker[0];
        for (int j = 1; j <= ksize; ++j) { int64_t newval = ker[j] + ker[j - 1]; ker[j - 1] = oldval; oldval = newval; }
    }
    for (int i = 0; i < order; ++i) {
        int64_t oldval = -ker[0];
        for (int j = 1; j <= ksize; ++j) { int64_t newval = ker[j - 1] - ker[j]; ker[j - 1] = oldval; oldval = newval; }
    }
    for (int i = 0; i < ksize; ++i) k[i] = (double)ker[i];
    free(ker);
    return ksize;
}

/* OpenCV getGaussianKernel(n, sigma) with sigma > 0, double precision. */
static void gaussian_kernel(int n, double sigma, double *k) {
    double scale2x = -0.5 / (sigma * sigma), sum = 0.0;
    for (int i = 0; i < n; ++i) {
        double x = i - (n - 1) * 0.5;
        k[i] = exp(scale2x * x * x);
        sum += k[i];
    }
    sum = 1.0 / sum;
    for (int i = 0; i < n; ++i) k[i] *= sum;
}

/* separable correlation, channels interleaved (cn), REFLECT_101, following OpenCV's CV_64F filter engine:
 *   row pass    = RowFilter<double,double>: plain left-to-right accumulation, first tap initialises the sum;
 *   column pass = SymmColumnFilter: symmetric kernels   s = k[c]*x[c];  s += k[c+j]*(x[c+j] + x[c-j]), j = 1..r
 *                                   asymmetric kernels  s = 0;          s += k[c+j]*(x[c+j] - x[c-j]), j = 1..r
 * ysym: +1 symmetric column kernel, -1 anti-symmetric column kernel. */
static void sep_filter(const double *src, double *dst, int h, int w, int cn, const double *kx, int nx,
                       const double *ky, int ny, int ysym) {
    double *tmp = (double *)malloc(sizeof(double) * (size_t)h * w * cn);
    int rx = nx / 2, ry = ny / 2;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < cn; ++c) {
                double s = kx[0] * src[((size_t)y * w + reflect101(x - rx, w)) * cn + c];
                for (int k = 1; k < nx; ++k) s += kx[k] * src[((size_t)y * w + reflect101(x + k - rx, w)) * cn + c];
                tmp[((size_t)y * w + x) * cn + c] = s;
            }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < cn; ++c) {
                double s;
                if (ysym > 0) {
                    s = ky[ry] * tmp[((size_t)y * w + x) * cn + c];
                    for (int k = 1; k <= ry; ++k)
                        s += ky[ry + k] * (tmp[((size_t)reflect101(y + k, h) * w + x) * cn + c] +
                                           tmp[((size_t)reflect101(y - k, h) * w + x) * cn + c]);
                } else {
                    s = 0.0;
                    for (int k = 1; k <= ry; ++k)
                        s += ky[ry + k] * (tmp[((size_t)reflect101(y + k, h) * w + x) * cn + c] -
                                           tmp[((size_t)reflect101(y - k, h) * w + x) * cn + c]);
                }
                dst[((size_t)y * w + x) * cn + c] = s;
            }
    free(tmp);
}

/* normalmap_generation.py:5-56.  pre_blur / sobel / post_blur: <= 0 means "None". */
int oracle_normalmap(const uint16_t *depth, int h, int w, int pre_blur, int sobel, int post_blur, int invert,
                     uint8_t *out_rgb) {
    size_t n = (size_t)h * w;
    double *z = (double *)malloc(sizeof(double) * n);
    double *zx = (double *)malloc(sizeof(double) * n);
    double *zy = (double *)malloc(sizeof(double) * n);
    double *nrm = (double *)malloc(sizeof(double) * n * 3);
    double kbuf[64], kbuf2[64];
    if (!z || !zx || !zy || !nrm) return -1;
    if (pre_blur > 63 || sobel > 63 || post_blur > 63) return -2;

    for (size_t i = 0; i < n; ++i) {                      /* :20-21 */
        double v = invert ? (double)depth[i] : (double)depth[i] * (-1.0);
        z[i] = v / 256.0;
    }
    if (pre_blur > 0) {                                   /* :23-24 GaussianBlur(z, (k,k), sigma=k) */
        gaussian_kernel(pre_blur, (double)pre_blur, kbuf);
        sep_filter(z, z, h, w, 1, kbuf, pre_blur, kbuf, pre_blur, +1);
    }
    if (sobel > 0) {                                      /* :27-29 */
        int nd_ = sobel_kernel(sobel, 1, kbuf), ns = sobel_kernel(sobel, 0, kbuf2);
        sep_filter(z, zx, h, w, 1, kbuf, nd_, kbuf2, ns, +1);
        sep_filter(z, zy, h, w, 1, kbuf2, ns, kbuf, nd_, -1);
    } else {                                              /* :31 np.gradient, edge_order 1 */
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                size_t i = (size_t)y * w + x;
                if (w == 1) zx[i] = 0.0;
                else if (x == 0) zx[i] = z[i + 1] - z[i];
                else if (x == w - 1) zx[i] = z[i] - z[i - 1];
                else zx[i] = (z[i + 1] - z[i - 1]) / 2.0;
                if (h == 1) zy[i] = 0.0;
                else if (y == 0) zy[i] = z[i + w] - z[i];
                else if (y == h - 1) zy[i] = z[i] - z[i - w];
                else zy[i] = (z[i + w] - z[i - w]) / 2.0;
            }
    }
    for (size_t i = 0; i < n; ++i) {                      /* :34-39 */
        double a = zx[i], b = -zy[i], c = 1.0;
        double nn = sqrt(a * a + b * b + c * c);
        nrm[i * 3 + 0] = a / nn; nrm[i * 3 + 1] = b / nn; nrm[i * 3 + 2] = c / nn;
    }
    if (post_blur > 0) {                                  /* :42-48 */
        gaussian_kernel(post_blur, (double)post_blur, kbuf);
        sep_filter(nrm, nrm, h, w, 3, kbuf, post_blur, kbuf, post_blur, +1);
        for (size_t i = 0; i < n; ++i) {
            double a = nrm[i * 3], b = nrm[i * 3 + 1], c = nrm[i * 3 + 2];
            double nn = sqrt(a * a + b * b + c * c);
            nrm[i * 3] = a / nn; nrm[i * 3 + 1] = b / nn; nrm[i * 3 + 2] = c / nn;
        }
    }
    for (size_t i = 0; i < n * 3; ++i) {                  /* :51-54 */
        double v = nrm[i];
        v += 1; v /= 2; v = v * 256;
        if (v < 0) v = 0;
        if (v > 256 - 0.1) v = 256 - 0.1;
        out_rgb[i] = (uint8_t)(int32_t)v;
    }
    free(z); free(zx); free(zy); free(nrm);
    return 0;
}

/* core.py:189-211 + :44-50 for a model prediction (float32 arithmetic throughout, as numpy keeps float32).
 * clip_mode: 0 = off, 1 = "Range" (normalise, clip to [far, near], renormalise).  Returns 1 when the prediction is
 * degenerate (all-zero output), 0 otherwise. */
int oracle_normalize_u16(const float *pred, int64_t n, int invert, int clip_mode, float clip_far, float clip_near,
                         uint16_t *out) {
    float mn = pred[0], mx = pred[0];
    for (int64_t i = 1; i < n; ++i) { if (pred[i] < mn) mn = pred[i]; if (pred[i] > mx) mx = pred[i]; }
    volatile float rng0 = mx - mn;                        /* float32 subtraction, compared with float64 eps */
    if (!(fabs((double)rng0) > 2.220446049250313e-16)) { memset(out, 0, sizeof(uint16_t) * (size_t)n); return 1; }
    float *tmp = (float *)malloc(sizeof(float) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) tmp[i] = invert ? pred[i] * -1.0f : pred[i];
    for (int pass = 0; pass < (clip_mode == 1 ? 2 : 1); ++pass) {
        float lo = tmp[0], hi = tmp[0];
        for (int64_t i = 1; i < n; ++i) { if (tmp[i] < lo) lo = tmp[i]; if (tmp[i] > hi) hi = tmp[i]; }
        volatile float den = hi - lo;
        for (int64_t i = 0; i < n; ++i) { volatile float num = tmp[i] - lo; tmp[i] = num / den; }
        if (clip_mode == 1 && pass == 0)
            for (int64_t i = 0; i < n; ++i) { if (tmp[i] < clip_far) tmp[i] = clip_far; if (tmp[i] > clip_near) tmp[i] = clip_near; }
    }
    const float hi_clip = (float)(65536 - 0.1);           /* np.clip bound cast to float32 = 65535.8984375 */
    for (int64_t i = 0; i < n; ++i) {                     /* convert_to_i16 */
        volatile float a = tmp[i] * 65536.0f;
        volatile float b = a + 0.0001f;
        float v = b;
        if (v < 0.0f) v = 0.0f;
        if (v > hi_clip) v = hi_clip;
        out[i] = (uint16_t)(int32_t)v;
    }
    free(tmp);
    return 0;
}
