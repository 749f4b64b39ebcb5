// M1 — tangent-space normal map.  Replaces src/normalmap_generation.py:5-56.
//
// Fast path (the reference's default: Sobel k=3, no blur) = ONE fused HBM-bound kernel, 2 B/px in, 3 B/px out:
//   * Sobel sums are done in exact integer arithmetic (u16 inputs, |sum| <= 4*65535), zx = -g/256 is exact in fp64;
//   * the reference then evaluates n = sqrt(zx^2 + zy^2 + 1), c/n, ((c/n)+1)/2*256, clip, TRUNCATE in fp64.  Only
//     the integer part of the last value is observable.  Each component is first evaluated in fp32 with a proven error
//     bound (< 2^-12 absolute on the [0,256] scale); when the fp32 value is farther than 2^-8 from an integer its
//     floor equals the fp64 floor and is used directly, otherwise the pixel re-evaluates the reference's exact fp64
//     sequence (IEEE sqrt.rn / div.rn).  Exact zeros (flat gradients) are resolved without either.
// General path (other Sobel sizes, np.gradient branch, Gaussian pre/post blur) = fp64 separable filters through a
// workspace, tap order following OpenCV's filters (see oracle/normalmap_oracle.c).
#include <math.h>

#include "common.cuh"

namespace dm {

__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

// reference's exact fp64 tail for one component: c = num / n; v = ((c + 1) / 2) * 256; clip; truncate
__device__ __forceinline__ uint8_t quant_exact(double num, double n) {
    double c = __ddiv_rn(num, n);
    double v = __dmul_rn(__ddiv_rn(__dadd_rn(c, 1.0), 2.0), 256.0);
    v = fmin(fmax(v, 0.0), 255.9);
    return (uint8_t)(int)v;
}

__device__ __forceinline__ uint8_t floor_if_safe(float v, bool &safe) {
    float r = rintf(v);
    safe = fabsf(v - r) > 0.00390625f && v > 0.0f;
    float f = floorf(v);
    return (uint8_t)(int)fminf(f, 255.0f);
}

// gx, gy: integer Sobel sums of the raw u16 depth (dx and dy), sgn = invert ? +1 : -1  (z = sgn * d / 256)
__device__ __forceinline__ void normal_from_grad(int gx, int gy, int sgn, uint8_t &r, uint8_t &g, uint8_t &b) {
    // zx = sgn*gx/256 ; normal = (zx, -zy, 1)/n
    const int ax = sgn * gx, ay = -sgn * gy;
    const float fx = (float)ax * 0.00390625f, fy = (float)ay * 0.00390625f;  // exact: |a| < 2^19
    const float s = fmaf(fx, fx, fmaf(fy, fy, 1.0f));
    const float rs = rsqrtf(s);
    bool sx, sy, sz;
    uint8_t qx = floor_if_safe(fmaf(fx * rs, 128.0f, 128.0f), sx);
    uint8_t qy = floor_if_safe(fmaf(fy * rs, 128.0f, 128.0f), sy);
    uint8_t qz = floor_if_safe(fmaf(rs, 128.0f, 128.0f), sz);
    if (ax == 0) { qx = 128; sx = true; }          // 0/n = 0 -> exactly 128
    if (ay == 0) { qy = 128; sy = true; }
    if (ax == 0 && ay == 0) { qz = 255; sz = true; }  // 1/1 -> 256 -> clipped to 255.9
    if (!(sx && sy && sz)) {
        const double dx = (double)ax * 0.00390625, dy = (double)ay * 0.00390625;
        // np.linalg.norm: sqrt((zx*zx + zy*zy) + 1*1); every term is exact here
        const double n = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), 1.0));
        if (!sx) qx = quant_exact(dx, n);
        if (!sy) qy = quant_exact(dy, n);
        if (!sz) qz = quant_exact(1.0, n);
    }
    r = qx; g = qy; b = qz;
}

// ---- lean variant of normal_from_grad for the W % 8 == 0 kernel: no XU-pipe conversions (int->float and floor through
// the 2^23 magic number), a tighter "safe" band (the fp32 value carries < 2^-13 absolute error on the [0,256] scale —
// rsqrt.approx 2 ulp, two multiplies, one fma at magnitude <= 256 — so 2^-10 still leaves a 8x margin) so that only
// ~0.6% of the pixels, i.e. ~1 warp in 6 instead of 1 in 2, takes the fp64 detour.
__device__ __forceinline__ float int_to_float_exact(int a) {        // |a| < 2^22
    return __int_as_float(0x4B400000 + a) - 12582912.0f;
}
__device__ __forceinline__ uint32_t floor_u8_if_safe(float v, bool &safe) {     // 0 <= v <= 256.01
    const float t = v + 8388608.0f;                 // round to nearest integer, result in the low mantissa bits
    const float d = v - (t - 8388608.0f);
    safe = fabsf(d) > 0.0009765625f;
    int q = (__float_as_int(t) & 0x1ff) - (d < 0.f ? 1 : 0);
    return (uint32_t)min(q, 255);
}
__device__ __forceinline__ uint32_t normal_rgb_packed(int gx, int gy, int sgn) {   // r | g << 8 | b << 16
    const int ax = sgn * gx, ay = -sgn * gy;
    const float fx = int_to_float_exact(ax) * 0.00390625f, fy = int_to_float_exact(ay) * 0.00390625f;
    const float s = fmaf(fx, fx, fmaf(fy, fy, 1.0f));
    const float rs = rsqrtf(s);
    const float r128 = rs * 128.0f;
    bool sx, sy, sz;
    uint32_t qx = floor_u8_if_safe(fmaf(fx, r128, 128.0f), sx);
    uint32_t qy = floor_u8_if_safe(fmaf(fy, r128, 128.0f), sy);
    uint32_t qz = floor_u8_if_safe(r128 + 128.0f, sz);
    if (ax == 0) { qx = 128; sx = true; }
    if (ay == 0) { qy = 128; sy = true; }
    if ((ax | ay) == 0) { qz = 255; sz = true; }
    if (!(sx && sy && sz)) {
        const double dx = (double)ax * 0.00390625, dy = (double)ay * 0.00390625;
        const double n = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), 1.0));
        if (!sx) qx = quant_exact(dx, n);
        if (!sy) qy = quant_exact(dy, n);
        if (!sz) qz = quant_exact(1.0, n);
    }
    return qx | (qy << 8) | (qz << 16);
}

constexpr int NM_TX = 128;  // threads per block, 4 px per thread -> 512 px of one row per block
constexpr int NM_PX = 4;

__global__ void __launch_bounds__(NM_TX) normalmap_sobel3_kernel(const uint16_t *__restrict__ depth, int H, int W, int sgn,
                                                                 uint8_t *__restrict__ out) {
    __shared__ __align__(16) uint8_t stage[NM_TX * NM_PX * 3];
    const int b = blockIdx.z, y = blockIdx.y;
    const int x_blk = blockIdx.x * (NM_TX * NM_PX);
    const uint16_t *img = depth + (int64_t)b * H * W;
    const int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H);
    const uint16_t *r0 = img + (int64_t)ym * W, *r1 = img + (int64_t)y * W, *r2 = img + (int64_t)yp * W;
    const int x0 = x_blk + threadIdx.x * NM_PX;
    if (x0 < W) {
        int c0[NM_PX + 2], c1[NM_PX + 2], c2[NM_PX + 2];
#pragma unroll
        for (int i = 0; i < NM_PX + 2; ++i) {
            int xx = x0 - 1 + i;
            xx = xx < W ? reflect101(xx, W) : reflect101(xx > W ? W - 1 : xx, W);
            c0[i] = __ldg(r0 + xx); c1[i] = __ldg(r1 + xx); c2[i] = __ldg(r2 + xx);
        }
#pragma unroll
        for (int i = 0; i < NM_PX; ++i) {
            const int gx = (c0[i + 2] - c0[i]) + 2 * (c1[i + 2] - c1[i]) + (c2[i + 2] - c2[i]);
            const int gy = (c2[i] - c0[i]) + 2 * (c2[i + 1] - c0[i + 1]) + (c2[i + 2] - c0[i + 2]);
            uint8_t r, g, bl;
            normal_from_grad(gx, gy, sgn, r, g, bl);
            uint8_t *s = stage + (threadIdx.x * NM_PX + i) * 3;
            s[0] = r; s[1] = g; s[2] = bl;
        }
    }
    __syncthreads();
    // coalesced write of this block's byte range of the output row
    const int npx = min(NM_TX * NM_PX, W - x_blk);
    const int nbytes = npx * 3;
    uint8_t *dst = out + ((int64_t)b * H + y) * (int64_t)W * 3 + (int64_t)x_blk * 3;
    const int head = (int)((4 - ((uintptr_t)dst & 3)) & 3);
    const int h = head < nbytes ? head : nbytes;
    if ((int)threadIdx.x < h) dst[threadIdx.x] = stage[threadIdx.x];
    const int nwords = (nbytes - h) >> 2;
    for (int i = threadIdx.x; i < nwords; i += NM_TX) {
        const uint8_t *s = stage + h + i * 4;
        uint32_t wv = (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24);
        *reinterpret_cast<uint32_t *>(dst + h + i * 4) = wv;
    }
    const int tail0 = h + nwords * 4;
    if ((int)threadIdx.x < nbytes - tail0) dst[tail0 + threadIdx.x] = stage[tail0 + threadIdx.x];
}

// W % 8 == 0: the three depth rows of the block (512 + 2 halo pixels each) are staged in shared memory with 16-byte
// loads, every thread turns 4 pixels into three packed 32-bit words, and the block's 1536 output bytes leave as 8-byte
// stores.  ~50 instructions per pixel instead of ~190 (per-load border reflection, byte-wise staging and XU-pipe
// conversions were what made the first version issue-bound at 12% of the HBM rate).
__global__ void __launch_bounds__(NM_TX) normalmap_sobel3_w8_kernel(const uint16_t *__restrict__ depth, int H, int W, int sgn,
                                                                    uint8_t *__restrict__ out) {
    constexpr int SPAN = NM_TX * NM_PX;                       // 512 pixels
    __shared__ __align__(16) uint16_t rows[3][SPAN + 8];      // element e of a row = pixel x_blk - 1 + (e - 3)... see ROFF
    __shared__ __align__(16) uint32_t words[SPAN * 3 / 4];
    constexpr int ROFF = 4;                                   // pixel x_blk sits at element ROFF (8-byte aligned), halo at ROFF-1
    const int b = blockIdx.z, y = blockIdx.y;
    const int x_blk = blockIdx.x * SPAN;
    const uint16_t *img = depth + (int64_t)b * H * W;
    const int ys[3] = {reflect101(y - 1, H), y, reflect101(y + 1, H)};
    const int npx = min(SPAN, W - x_blk);                     // multiple of 8
    for (int i = threadIdx.x; i < 3 * (SPAN / 8); i += NM_TX) {
        const int r = i / (SPAN / 8), v = i - r * (SPAN / 8);
        if (v * 8 < npx) {
            const uint4 u = __ldg(reinterpret_cast<const uint4 *>(img + (int64_t)ys[r] * W + x_blk + v * 8));
            *reinterpret_cast<uint2 *>(&rows[r][ROFF + v * 8]) = make_uint2(u.x, u.y);
            *reinterpret_cast<uint2 *>(&rows[r][ROFF + v * 8 + 4]) = make_uint2(u.z, u.w);
        }
    }
    if (threadIdx.x < 6) {                                    // the two halo pixels of each row (border: reflect 101)
        const int r = threadIdx.x >> 1, right = threadIdx.x & 1;
        const int xx = right ? x_blk + npx : x_blk - 1;
        rows[r][right ? ROFF + npx : ROFF - 1] = __ldg(img + (int64_t)ys[r] * W + reflect101(xx, W));
    }
    __syncthreads();
    const int px0 = threadIdx.x * NM_PX;
    if (px0 < npx) {
        int c[3][6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            // elements ROFF-1+px0 .. ROFF+4+px0: one 2-byte, one 8-byte and one 2-byte access around the aligned quad
            const uint16_t *rp = &rows[r][ROFF + px0];
            const uint2 m = *reinterpret_cast<const uint2 *>(rp);
            c[r][0] = rp[-1];
            c[r][1] = m.x & 0xffff; c[r][2] = m.x >> 16; c[r][3] = m.y & 0xffff; c[r][4] = m.y >> 16;
            c[r][5] = rp[4];
        }
        uint32_t px[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gx = (c[0][i + 2] - c[0][i]) + 2 * (c[1][i + 2] - c[1][i]) + (c[2][i + 2] - c[2][i]);
            const int gy = (c[2][i] - c[0][i]) + 2 * (c[2][i + 1] - c[0][i + 1]) + (c[2][i + 2] - c[0][i + 2]);
            px[i] = normal_rgb_packed(gx, gy, sgn);
        }
        words[threadIdx.x * 3 + 0] = px[0] | (px[1] << 24);
        words[threadIdx.x * 3 + 1] = (px[1] >> 8) | (px[2] << 16);
        words[threadIdx.x * 3 + 2] = (px[2] >> 16) | (px[3] << 8);
    }
    __syncthreads();
    // row pitch W * 3 and the block offset 1536 k are multiples of 8 (W % 8 == 0), not of 16: 8-byte stores
    uint8_t *dst = out + ((int64_t)b * H + y) * (int64_t)W * 3 + (int64_t)x_blk * 3;
    const int nvec = npx * 3 / 8;                             // npx % 8 == 0 -> exact
    for (int i = threadIdx.x; i < nvec; i += NM_TX)
        *reinterpret_cast<uint2 *>(dst + i * 8) = *reinterpret_cast<const uint2 *>(&words[i * 2]);
}

// ---------------------------------------------------------------------------------------------------------------
// general fp64 path
// ---------------------------------------------------------------------------------------------------------------
constexpr int MAX_TAPS = 64;
struct Taps { double k[MAX_TAPS]; int n; int symmetric; };

__global__ void depth_to_z_kernel(const uint16_t *__restrict__ depth, int64_t total, int invert, double *__restrict__ z) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        double d = (double)depth[i];
        d = invert ? d : __dmul_rn(d, -1.0);
        z[i] = __ddiv_rn(d, 256.0);
    }
}

// 1-D correlation along x (axis=1) or y (axis=0); cn interleaved channels; REFLECT_101
__global__ void filter1d_kernel(const double *__restrict__ src, double *__restrict__ dst, int B, int H, int W, int cn, int axis, Taps t) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)B * H * W * cn;
    if (idx >= total) return;
    const int c = (int)(idx % cn);
    int64_t p = idx / cn;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    const double *img = src + (int64_t)b * H * W * cn;
    const int r = t.n / 2;
    const int len = axis ? W : H, pos = axis ? x : y;
    auto at = [&](int q) -> double {
        q = reflect101(q, len);
        return axis ? img[((int64_t)y * W + q) * cn + c] : img[((int64_t)q * W + x) * cn + c];
    };
    // Tap order follows OpenCV's CV_64F engine (see oracle/normalmap_oracle.c): rows accumulate left to right with the
    // first tap initialising the sum; columns use the symmetric / anti-symmetric pair forms, centre outward.
    double s;
    if (axis == 1 || t.symmetric == 0) {
        s = __dmul_rn(t.k[0], at(pos - r));
        for (int k = 1; k < t.n; ++k) s = __dadd_rn(s, __dmul_rn(t.k[k], at(pos + k - r)));
    } else if (t.symmetric > 0) {
        s = __dmul_rn(t.k[r], at(pos));
        for (int k = 1; k <= r; ++k) s = __dadd_rn(s, __dmul_rn(t.k[r + k], __dadd_rn(at(pos + k), at(pos - k))));
    } else {
        s = 0.0;
        for (int k = 1; k <= r; ++k) s = __dadd_rn(s, __dmul_rn(t.k[r + k], __dsub_rn(at(pos + k), at(pos - k))));
    }
    dst[idx] = s;
}

__global__ void gradient_kernel(const double *__restrict__ z, int B, int H, int W, double *__restrict__ zx, double *__restrict__ zy) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)B * H * W;
    if (idx >= total) return;
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    double gx, gy;
    if (W == 1) gx = 0.0;
    else if (x == 0) gx = __dsub_rn(z[idx + 1], z[idx]);
    else if (x == W - 1) gx = __dsub_rn(z[idx], z[idx - 1]);
    else gx = __ddiv_rn(__dsub_rn(z[idx + 1], z[idx - 1]), 2.0);
    if (H == 1) gy = 0.0;
    else if (y == 0) gy = __dsub_rn(z[idx + W], z[idx]);
    else if (y == H - 1) gy = __dsub_rn(z[idx], z[idx - W]);
    else gy = __ddiv_rn(__dsub_rn(z[idx + W], z[idx - W]), 2.0);
    zx[idx] = gx;
    zy[idx] = gy;
}

__global__ void make_normal_kernel(const double *__restrict__ zx, const double *__restrict__ zy, int64_t total, double *__restrict__ nrm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double a = zx[i], b = __dmul_rn(zy[i], -1.0);
    const double n = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), 1.0));
    nrm[i * 3 + 0] = __ddiv_rn(a, n);
    nrm[i * 3 + 1] = __ddiv_rn(b, n);
    nrm[i * 3 + 2] = __ddiv_rn(1.0, n);
}

__global__ void renorm_quant_kernel(const double *__restrict__ nrm, int64_t total, int renorm, uint8_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    double a = nrm[i * 3], b = nrm[i * 3 + 1], c = nrm[i * 3 + 2];
    if (renorm) {
        const double n = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
        a = __ddiv_rn(a, n); b = __ddiv_rn(b, n); c = __ddiv_rn(c, n);
    }
    double v[3] = {a, b, c};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double q = __dmul_rn(__ddiv_rn(__dadd_rn(v[k], 1.0), 2.0), 256.0);
        q = fmin(fmax(q, 0.0), 255.9);
        if (!(q == q)) q = 0.0;
        out[i * 3 + k] = (uint8_t)(int)q;
    }
}

static int sobel_taps(int ksize, int order, Taps &t) {
    t.symmetric = order ? -1 : 1;
    if (ksize == 1) {
        if (order == 0) { t.k[0] = 1.0; t.n = 1; }
        else { t.k[0] = -1.0; t.k[1] = 0.0; t.k[2] = 1.0; t.n = 3; }
        return 0;
    }
    if (ksize > MAX_TAPS - 1) return -1;
    long long ker[MAX_TAPS + 1];
    for (int i = 0; i <= ksize; ++i) ker[i] = 0;
    ker[0] = 1;
    for (int i = 0; i < ksize - order - 1; ++i) {
        long long oldval = ker[0];
        for (int j = 1; j <= ksize; ++j) { long long nv = ker[j] + ker[j - 1]; ker[j - 1] = oldval; oldval = nv; }
    }
    for (int i = 0; i < order; ++i) {
        long long oldval = -ker[0];
        for (int j = 1; j <= ksize; ++j) { long long nv = ker[j - 1] - ker[j]; ker[j - 1] = oldval; oldval = nv; }
    }
    for (int i = 0; i < ksize; ++i) t.k[i] = (double)ker[i];
    t.n = ksize;
    return 0;
}

static int gauss_taps(int n, Taps &t) {
    if (n > MAX_TAPS - 1) return -1;
    const double sigma = (double)n;
    const double scale2x = -0.5 / (sigma * sigma);
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        double x = i - (n - 1) * 0.5;
        t.k[i] = exp(scale2x * x * x);
        sum += t.k[i];
    }
    sum = 1.0 / sum;
    for (int i = 0; i < n; ++i) t.k[i] *= sum;
    t.n = n;
    t.symmetric = 1;
    return 0;
}

static bool fast_path(int pre_blur, int sobel, int post_blur) { return pre_blur <= 0 && post_blur <= 0 && sobel == 3; }

}  // namespace dm

extern "C" __attribute__((visibility("default"))) size_t dm_normalmap_workspace_bytes(int B, int H, int W, int pre_blur, int sobel, int post_blur) {
    if (dm::fast_path(pre_blur, sobel, post_blur)) return 256;
    // z, tmp, zx, zy (1 plane each) + nrm, nrm_tmp (3 planes each), fp64
    return dm::align_up((size_t)B * H * W * sizeof(double) * 10, 256) + 256;
}

extern "C" __attribute__((visibility("default"))) int dm_normalmap(const uint16_t *depth, int B, int H, int W, int pre_blur, int sobel, int post_blur, int invert,
                            uint8_t *rgb_out, void *workspace, size_t workspace_bytes, void *stream_) {
    using namespace dm;
    if (!depth || !rgb_out || B <= 0 || H <= 0 || W <= 0) { set_error("dm_normalmap: bad arguments"); return DM_E_INVALID; }
    cudaStream_t stream = (cudaStream_t)stream_;
    if (fast_path(pre_blur, sobel, post_blur)) {
        dim3 grid((W + NM_TX * NM_PX - 1) / (NM_TX * NM_PX), H, B);
        if (H > 65535 || B > 65535) { set_error("dm_normalmap: H or B too large"); return DM_E_UNSUPPORTED; }
        if (W % 8 == 0) normalmap_sobel3_w8_kernel<<<grid, NM_TX, 0, stream>>>(depth, H, W, invert ? 1 : -1, rgb_out);
        else normalmap_sobel3_kernel<<<grid, NM_TX, 0, stream>>>(depth, H, W, invert ? 1 : -1, rgb_out);
        DM_LAUNCH_CHECK("normalmap_sobel3_kernel");
        return DM_OK;
    }
    if ((pre_blur > 0 && pre_blur % 2 == 0) || (post_blur > 0 && post_blur % 2 == 0) || (sobel > 0 && sobel % 2 == 0)) {
        set_error("dm_normalmap: kernel sizes must be odd");  // cv2 raises for even sizes too
        return DM_E_INVALID;
    }
    if (!workspace || workspace_bytes < dm_normalmap_workspace_bytes(B, H, W, pre_blur, sobel, post_blur)) {
        set_error("dm_normalmap: workspace too small");
        return DM_E_WORKSPACE;
    }
    const int64_t n = (int64_t)B * H * W;
    double *z = (double *)workspace, *tmp = z + n, *zx = tmp + n, *zy = zx + n, *nrm = zy + n, *nrm_tmp = nrm + 3 * n;
    const int T = 256;
    const unsigned g1 = (unsigned)((n + T - 1) / T), g3 = (unsigned)((3 * n + T - 1) / T);
    depth_to_z_kernel<<<g1, T, 0, stream>>>(depth, n, invert ? 1 : 0, z);
    DM_LAUNCH_CHECK("depth_to_z_kernel");
    Taps tg, td, ts;
    if (pre_blur > 0) {
        if (gauss_taps(pre_blur, tg)) { set_error("dm_normalmap: pre_blur kernel too large"); return DM_E_UNSUPPORTED; }
        filter1d_kernel<<<g1, T, 0, stream>>>(z, tmp, B, H, W, 1, 1, tg);
        filter1d_kernel<<<g1, T, 0, stream>>>(tmp, z, B, H, W, 1, 0, tg);
        DM_LAUNCH_CHECK("filter1d_kernel(pre_blur)");
    }
    if (sobel > 0) {
        if (sobel_taps(sobel, 1, td) || sobel_taps(sobel, 0, ts)) { set_error("dm_normalmap: sobel kernel too large"); return DM_E_UNSUPPORTED; }
        filter1d_kernel<<<g1, T, 0, stream>>>(z, tmp, B, H, W, 1, 1, td);
        filter1d_kernel<<<g1, T, 0, stream>>>(tmp, zx, B, H, W, 1, 0, ts);
        filter1d_kernel<<<g1, T, 0, stream>>>(z, tmp, B, H, W, 1, 1, ts);
        filter1d_kernel<<<g1, T, 0, stream>>>(tmp, zy, B, H, W, 1, 0, td);
        DM_LAUNCH_CHECK("filter1d_kernel(sobel)");
    } else {
        gradient_kernel<<<g1, T, 0, stream>>>(z, B, H, W, zx, zy);
        DM_LAUNCH_CHECK("gradient_kernel");
    }
    make_normal_kernel<<<g1, T, 0, stream>>>(zx, zy, n, nrm);
    DM_LAUNCH_CHECK("make_normal_kernel");
    if (post_blur > 0) {
        if (gauss_taps(post_blur, tg)) { set_error("dm_normalmap: post_blur kernel too large"); return DM_E_UNSUPPORTED; }
        filter1d_kernel<<<g3, T, 0, stream>>>(nrm, nrm_tmp, B, H, W, 3, 1, tg);
        filter1d_kernel<<<g3, T, 0, stream>>>(nrm_tmp, nrm, B, H, W, 3, 0, tg);
        DM_LAUNCH_CHECK("filter1d_kernel(post_blur)");
    }
    renorm_quant_kernel<<<g1, T, 0, stream>>>(nrm, n, post_blur > 0 ? 1 : 0, rgb_out);
    DM_LAUNCH_CHECK("renorm_quant_kernel");
    return DM_OK;
}
