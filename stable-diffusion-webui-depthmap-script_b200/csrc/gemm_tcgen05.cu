// Tensor-core GEMM / implicit-GEMM convolution for sm_100a: C[M,N] = A[M,K] * W[N,K]^T with fused epilogues.
//
// Used for every Linear layer of the ViT (QKV, proj, fc1, fc2 — reference: timm Block / dinov2_layers/{attention,mlp}.py),
// the patch embedding, the DPT 1x1 convs, ConvTranspose k=s layers (GEMM + pixel-shuffle store) and, with CONV=true,
// the 3x3 stride-1 pad-1 convolutions of the DPT decoder (reference: dmidas/blocks.py, depth_anything_v2/util/blocks.py)
// as implicit GEMM over NHWC activations: the 9 taps are 9 shifted TMA boxes, zero padding = TMA out-of-bounds fill.
//
// Structure (one 128 x BN output tile per CTA, 2 CTAs resident per SM so one tile's epilogue overlaps another's MMA):
//   warp 0  : TMA producer   — cp.async.bulk.tensor into a STAGES-deep ring of 128B-swizzled smem tiles (mbarrier tx)
//   warp 1  : MMA issuer     — one elected thread issues tcgen05.mma (kind::f16, fp32 accumulate in TMEM, 128 x BN x 16),
//                              tcgen05.commit releases smem slots / signals the epilogue
//   warps 2-5: epilogue      — tcgen05.ld (thread = accumulator row), bias / GELU / ReLU / LayerScale+residual /
//                              pixel-shuffle / fused 1x1 head, vectorised global stores
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dm {
using namespace tc;

enum GemmEpi : int {
    EPI_STORE_F16 = 0,   // C = act(acc + bias) [+ R]           -> fp16 [M, ldc]   (+ optional relu copy C2)
    EPI_RESID_F32 = 1,   // X += gamma[n] * (acc + bias[n])     -> fp32 in place (LayerScale + residual)
    EPI_PIXSHUF = 2,     // ConvTranspose k=s: n = (i, j, co) scattered to out[b, s*y+i, s*x+j, co]
    EPI_HEAD = 3,        // relu(acc + bias) . w2 + b2 -> relu -> fp32 [M]   (conv3x3 -> ReLU -> conv1x1 -> ReLU fused; BN = N)
    EPI_STORE_F32 = 4,   // C = acc + bias -> fp32 [M, ldc]
};
enum GemmAct : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

struct GemmParams {
    int M, N, K;
    int epi, act;
    const float *bias;      // [N] or null
    __half *C; int ldc;     // fp16 output
    __half *C2;             // optional relu(C) copy (same layout) or null
    const __half *R; int ldr;  // optional fp16 residual added before the store (EPI_STORE_F16)
    const __half *R2; int ldr2;  // optional second fp16 residual
    float *X; int ldx;      // fp32 residual stream (EPI_RESID_F32) / fp32 output (EPI_STORE_F32, EPI_HEAD)
    const float *gamma;     // [N] LayerScale (EPI_RESID_F32) / w2 (EPI_HEAD)
    float head_b2;
    // pixel shuffle
    int ps_s, ps_cout, ps_h, ps_w;
    // implicit conv geometry (CONV): activations [B, H, W, Cin]; tile = hbox x wbox pixels
    int cB, cH, cW, cCin, hbox, wbox, tiles_x, tiles_y;
    int tile_group;         // m-tiles per L2 group of the persistent tile order (set by the launcher)
};

template <int BN>
struct GemmCfg {
    static constexpr int BM = 128, BK = 64;
    static constexpr int STAGES = BN >= 256 ? 3 : (BN == 128 ? 3 : 4);
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGING_BYTES = 4 * 4096;   // one 32x32 fp32 transpose buffer per epilogue warp
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 256 /*barriers*/;   // base is __align__(1024)
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

// packed fp32x2 arithmetic (FADD2 / FFMA2 on sm_100): one issue slot for two accumulator columns
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// exact-erf GELU for two columns, one MUFU each.  With u = min(|x|, 5.75):
//     gelu(x) = max(x, 0) - u * 2^(-u Q(u) - 1),      Q(u) = -log2(erfc(u / sqrt 2)) / u
// Q is smooth and nearly linear on [0, 5.75]; the degree-5 fit below (Lawson-weighted on the GELU error) keeps
// |gelu - x Phi(x)| < 3.2e-7 in fp32 over all x (tests/test_gelu_formula.py checks the same numbers on
// the host), i.e. three orders below the fp16 rounding of the stored activation.  Past the clamp u*2^(..) < 3e-8.
__device__ __forceinline__ void gelu_erf2(float &x0, float &x1) {
    const float u0 = fminf(fabsf(x0), 5.75f), u1 = fminf(fabsf(x1), 5.75f);
    const uint64_t u = pack2(u0, u1);
    uint64_t q = pack2(-2.992472582263872e-05f, -2.992472582263872e-05f);
    q = fma2(q, u, pack2(0.0007398786256089807f, 0.0007398786256089807f));
    q = fma2(q, u, pack2(-0.007977476343512535f, -0.007977476343512535f));
    q = fma2(q, u, pack2(0.05323820561170578f, 0.05323820561170578f));
    q = fma2(q, u, pack2(0.45891568064689636f, 0.45891568064689636f));
    q = fma2(q, u, pack2(1.1511471271514893f, 1.1511471271514893f));
    const uint64_t nu = pack2(-u0, -u1);
    float t0, t1, e0, e1;
    unpack2(fma2(nu, q, pack2(-1.f, -1.f)), t0, t1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
    unpack2(fma2(nu, pack2(e0, e1), pack2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), x0, x1);
}
__device__ __forceinline__ float gelu_erf(float x) {
    float y = x;
    gelu_erf2(x, y);
    return x;
}

// Epilogue for one accumulator tile: thread owns accumulator row (TMEM lane) `q*32 + lane`, walks BN columns in chunks of 32.
template <int BN>
__device__ __forceinline__ void epilogue_rows(const GemmParams &p, uint32_t tmem_acc, int q, long long m, bool row_ok, int n_base) {
        float head_acc = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            const int n0 = n_base + c0;
            if (!row_ok || n0 >= p.N) continue;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (p.bias) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4 *>(p.bias + n0 + j));
                    v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                }
            }
            if (p.epi == EPI_RESID_F32) {
                float *xr = p.X + m * p.ldx + n0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 x4 = *reinterpret_cast<float4 *>(xr + j);
                    const float4 g4 = __ldg(reinterpret_cast<const float4 *>(p.gamma + n0 + j));
                    x4.x = fmaf(g4.x, v[j], x4.x); x4.y = fmaf(g4.y, v[j + 1], x4.y);
                    x4.z = fmaf(g4.z, v[j + 2], x4.z); x4.w = fmaf(g4.w, v[j + 3], x4.w);
                    *reinterpret_cast<float4 *>(xr + j) = x4;
                }
                continue;
            }
            if (p.epi == EPI_STORE_F32) {
                float *xr = p.X + m * p.ldx + n0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(xr + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                continue;
            }
            if (p.act == ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) gelu_erf2(v[j], v[j + 1]);
            } else if (p.act == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.epi == EPI_HEAD) {
#pragma unroll
                for (int j = 0; j < 32; ++j) head_acc = fmaf(v[j], __ldg(p.gamma + n0 + j), head_acc);
                continue;
            }
            if (p.R) {
                const __half *rr = p.R + m * p.ldr + n0;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    const uint4 u = __ldg(reinterpret_cast<const uint4 *>(rr + j));
                    const __half2 *h2 = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(h2[k]); v[j + 2 * k] += f.x; v[j + 2 * k + 1] += f.y; }
                }
            }
            if (p.R2) {
                const __half *rr = p.R2 + m * p.ldr2 + n0;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    const uint4 u = __ldg(reinterpret_cast<const uint4 *>(rr + j));
                    const __half2 *h2 = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(h2[k]); v[j + 2 * k] += f.x; v[j + 2 * k + 1] += f.y; }
                }
            }
            __half *dst;
            __half *dst2 = nullptr;
            if (p.epi == EPI_PIXSHUF) {
                const int s = p.ps_s, ij = n0 / p.ps_cout, co = n0 % p.ps_cout;
                const int i = ij / s, jx = ij % s;
                const long long bb = m / ((long long)p.ps_h * p.ps_w);
                const int rem = (int)(m % ((long long)p.ps_h * p.ps_w));
                const int y = rem / p.ps_w, x = rem % p.ps_w;
                dst = p.C + (((bb * (p.ps_h * s) + (y * s + i)) * (long long)(p.ps_w * s)) + (x * s + jx)) * p.ps_cout + co;
            } else {
                dst = p.C + m * p.ldc + n0;
                if (p.C2) dst2 = p.C2 + m * p.ldc + n0;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint4 u;
                __half2 *h2 = reinterpret_cast<__half2 *>(&u);
#pragma unroll
                for (int k = 0; k < 4; ++k) h2[k] = __floats2half2_rn(v[j + 2 * k], v[j + 2 * k + 1]);
                *reinterpret_cast<uint4 *>(dst + j) = u;
                if (dst2) {
                    const __half2 z = __float2half2_rn(0.f);
#pragma unroll
                    for (int k = 0; k < 4; ++k) h2[k] = __hmax2(h2[k], z);
                    *reinterpret_cast<uint4 *>(dst2 + j) = u;
                }
            }
        }
        if (p.epi == EPI_HEAD && row_ok) p.X[m] = fmaxf(head_acc + p.head_b2, 0.f);
}

// Coalesced epilogue.  tcgen05.ld hands each thread one accumulator ROW, so storing straight from registers makes
// every warp-level store touch 32 different cache lines (16 B each).  Instead each warp transposes its 32x32 fp32
// chunk through a private 4 KB smem buffer (16-byte chunks XOR-swizzled by row, conflict-free both ways) and then
// 8 lanes cover one row's 128 B, so a warp-level access touches 4 full lines: 8x fewer LSU line transactions for
// the fp32 residual-stream read-modify-write, 4x fewer for fp16 stores, and fp16 residual reads ride the same pattern.
template <int BN, int SPEC = 0>
__device__ __forceinline__ void epilogue_rows_staged(const GemmParams &p, uint32_t tmem_acc, int q, int lane, long long m, bool row_ok,
                                                     int n_base, float *stage, int col_begin = 0, int col_end = BN) {
    const int sub = lane >> 3, cc = (lane & 7) * 4;
    // SPEC > 0: compile-time epilogue (1 = fp16 store + bias, 2 = the same + GELU; no residual inputs, no ReLU copy): the
    // mode tests below fold away instead of being re-evaluated (constant-bank loads, branches) for every chunk and row
    const int epi = SPEC ? (int)EPI_STORE_F16 : p.epi;
    const int act = SPEC == 2 ? (int)ACT_GELU : (SPEC == 1 ? (int)ACT_NONE : p.act);
    const __half *R = SPEC ? nullptr : p.R, *R2 = SPEC ? nullptr : p.R2;
    __half *C2 = SPEC ? nullptr : p.C2;
    const bool has_bias = SPEC ? true : p.bias != nullptr;
    // bias (thread = row domain: all 32 columns of the chunk) is prefetched one chunk ahead so its L2 latency hides
    // behind the previous chunk's work instead of sitting between tcgen05.ld and the first FADD
    float4 bcur[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bcur[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_bias && n_base + col_begin < p.N) {
#pragma unroll
        for (int j = 0; j < 8; ++j) bcur[j] = __ldg(reinterpret_cast<const float4 *>(p.bias + n_base + col_begin + 4 * j));
    }
    // the accumulator chunk is software-pipelined as well: chunk c+1 is in flight from TMEM while chunk c is processed
    uint32_t rn[32];
    tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)col_begin, rn);
#pragma unroll 1
    for (int c0 = col_begin; c0 < col_end; c0 += 32) {
        const int n0 = n_base + c0;
        const bool col_ok = n0 < p.N;
        float4 bnext[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bnext[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_bias && c0 + 32 < col_end && n0 + 32 < p.N) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bnext[j] = __ldg(reinterpret_cast<const float4 *>(p.bias + n0 + 32 + 4 * j));
        }
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (epi == EPI_RESID_F32 && col_ok) g4 = __ldg(reinterpret_cast<const float4 *>(p.gamma + n0 + cc));
        uint32_t r[32];
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = rn[j];
        if (c0 + 32 < col_end) tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c0 + 32), rn);
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            unpack2(add2(pack2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])), pack2(bcur[j].x, bcur[j].y)), v[4 * j], v[4 * j + 1]);
            unpack2(add2(pack2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), pack2(bcur[j].z, bcur[j].w)), v[4 * j + 2], v[4 * j + 3]);
        }
        if (epi == EPI_STORE_F16 || epi == EPI_PIXSHUF) {
            if (act == ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) gelu_erf2(v[j], v[j + 1]);
            } else if (act == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
        }
        const bool f16_out = epi == EPI_STORE_F16 || epi == EPI_PIXSHUF;
        const int hsel = ((c0 - col_begin) >> 5) & 1;          // fp16 outputs: two 32-column chunks share one 64-column round
        if (f16_out) {
            // row `lane` of the staging tile holds 64 halves (128 B); this chunk fills 16-byte slots hsel*4 .. hsel*4+3
            if (hsel == 0) __syncwarp();
            uint8_t *srow = reinterpret_cast<uint8_t *>(stage) + lane * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 u;
                __half2 *h2 = reinterpret_cast<__half2 *>(&u);
#pragma unroll
                for (int k = 0; k < 4; ++k) h2[k] = __floats2half2_rn(v[8 * c + 2 * k], v[8 * c + 2 * k + 1]);
                *reinterpret_cast<uint4 *>(srow + (((hsel * 4 + c) ^ (lane & 7)) << 4)) = u;
            }
            const bool round_done = hsel == 1 || c0 + 32 >= col_end;
            if (!round_done) {
#pragma unroll
                for (int j = 0; j < 8; ++j) bcur[j] = bnext[j];
                continue;
            }
            __syncwarp();
        } else {
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4 *>(stage + lane * 32 + ((c ^ (lane & 7)) << 2)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            __syncwarp();
        }
        // pixel-shuffle column mapping is uniform for the chunk
        const int nr0 = f16_out ? n0 - hsel * 32 : n0;          // first column of this round
        const int ncols_round = f16_out ? (hsel + 1) * 32 : 32;
        int ps_i = 0, ps_j = 0, ps_co = 0;
        if (epi == EPI_PIXSHUF) { const int ij = nr0 / p.ps_cout; ps_co = nr0 % p.ps_cout; ps_i = ij / p.ps_s; ps_j = ij % p.ps_s; }
        // ---- transposed domain: 8 lanes cover one row's 128 B; all global loads of the chunk are issued first ----
        long long m_r[8];
        bool ok[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int rr = 4 * it + sub;
            m_r[it] = __shfl_sync(0xffffffffu, m, rr);
            ok[it] = __shfl_sync(0xffffffffu, (int)row_ok, rr) != 0 && (f16_out || col_ok);
        }
        if (epi == EPI_RESID_F32) {
            float4 x4[8];
#pragma unroll
            for (int it = 0; it < 8; ++it)
                if (ok[it]) x4[it] = *reinterpret_cast<const float4 *>(p.X + m_r[it] * p.ldx + n0 + cc);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = 4 * it + sub;
                const float4 a = *reinterpret_cast<const float4 *>(stage + rr * 32 + (((lane & 7) ^ (rr & 7)) << 2));
                if (!ok[it]) continue;
                float4 x = x4[it];
                x.x = fmaf(g4.x, a.x, x.x); x.y = fmaf(g4.y, a.y, x.y); x.z = fmaf(g4.z, a.z, x.z); x.w = fmaf(g4.w, a.w, x.w);
                *reinterpret_cast<float4 *>(p.X + m_r[it] * p.ldx + n0 + cc) = x;
            }
        } else if (epi == EPI_STORE_F32) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = 4 * it + sub;
                const float4 a = *reinterpret_cast<const float4 *>(stage + rr * 32 + (((lane & 7) ^ (rr & 7)) << 2));
                if (ok[it]) *reinterpret_cast<float4 *>(p.X + m_r[it] * p.ldx + n0 + cc) = a;
            }
        } else {
            // fp16 outputs: 8 lanes cover one row's 64 columns (128 B = one full line), 16 B per lane
            const int ch = lane & 7;
            const bool lane_ok = ch * 8 < ncols_round && nr0 + ch * 8 < p.N;
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {      // two batches of four rows keep the residual prefetch at 32 registers
                uint4 r1[4], r2[4];
                if (R) {
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) { const int it = hb * 4 + i4; if (ok[it] && lane_ok) r1[i4] = __ldg(reinterpret_cast<const uint4 *>(R + m_r[it] * p.ldr + nr0 + ch * 8)); }
                }
                if (R2) {
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) { const int it = hb * 4 + i4; if (ok[it] && lane_ok) r2[i4] = __ldg(reinterpret_cast<const uint4 *>(R2 + m_r[it] * p.ldr2 + nr0 + ch * 8)); }
                }
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const int it = hb * 4 + i4;
                    const int rr = 4 * it + sub;
                    uint4 u = *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(stage) + rr * 128 + ((ch ^ (rr & 7)) << 4));
                    if (!ok[it] || !lane_ok) continue;
                    if (R || R2) {
                        __half2 *h2 = reinterpret_cast<__half2 *>(&u);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float2 f = __half22float2(h2[k]);
                            if (R) { const float2 a = __half22float2(reinterpret_cast<const __half2 *>(&r1[i4])[k]); f.x += a.x; f.y += a.y; }
                            if (R2) { const float2 a = __half22float2(reinterpret_cast<const __half2 *>(&r2[i4])[k]); f.x += a.x; f.y += a.y; }
                            h2[k] = __floats2half2_rn(f.x, f.y);
                        }
                    }
                    __half *dst;
                    if (epi == EPI_PIXSHUF) {
                        const int s_ = p.ps_s;
                        const long long bb = m_r[it] / ((long long)p.ps_h * p.ps_w);
                        const int rem = (int)(m_r[it] % ((long long)p.ps_h * p.ps_w));
                        const int y = rem / p.ps_w, x = rem % p.ps_w;
                        dst = p.C + (((bb * (p.ps_h * s_) + (y * s_ + ps_i)) * (long long)(p.ps_w * s_)) + (x * s_ + ps_j)) * p.ps_cout + ps_co + ch * 8;
                    } else {
                        dst = p.C + m_r[it] * p.ldc + nr0 + ch * 8;
                    }
                    *reinterpret_cast<uint4 *>(dst) = u;
                    if (C2 && epi != EPI_PIXSHUF) {
                        const __half2 z = __float2half2_rn(0.f);
                        __half2 *h2 = reinterpret_cast<__half2 *>(&u);
#pragma unroll
                        for (int k = 0; k < 4; ++k) h2[k] = __hmax2(h2[k], z);
                        *reinterpret_cast<uint4 *>(C2 + m_r[it] * p.ldc + nr0 + ch * 8) = u;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) bcur[j] = bnext[j];
    }
}

// Residual-stream epilogue (EPI_RESID_F32: X += gamma * (acc + bias)) of the persistent kernels.  The fp32 residual row
// segments are the only long-latency operands of the whole epilogue, so they are fetched one 32-column chunk AHEAD
// (chunk 0 even before the accumulator-ready barrier is waited on) and the proj / fc2 tiles no longer pay one exposed
// HBM/L2 round trip per chunk.  bias and gamma are applied in the transposed (store) domain, where a lane owns four
// fixed columns, so they cost one float4 each per chunk instead of 32 registers per thread.
template <int BN>
__device__ __forceinline__ void epilogue_resid_staged(const GemmParams &p, uint32_t tmem_acc, int q, int lane, long long m, bool row_ok,
                                                      int n_base, float *stage, int col_begin, int col_end, uint64_t *bar, uint32_t parity) {
    const int sub = lane >> 3, cc = (lane & 7) * 4;
    // plain GEMM rows are consecutive: row rr of this warp is m - lane + rr, the store-domain rows of a lane are 4 apart.
    // Rows past M (last m-tile only) are clamped for the prefetch so every load is unconditional; only the store is guarded.
    // The persistent kernels require N % 256 == 0, so every column of the tile exists.
    const long long m_first = m - lane + sub;
    const float *xbase = p.X + n_base + cc;
    uint32_t xoff[8];                                  // element offsets: the caller guarantees M * ldx < 2^31
    unsigned okmask = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const long long mm = m_first + 4 * it;
        if (mm < p.M) okmask |= 1u << it;
        xoff[it] = (uint32_t)(mm < p.M ? mm : (long long)p.M - 1) * (uint32_t)p.ldx;
    }
    auto load_x = [&](float4 (&x)[8], int c0) {
#pragma unroll
        for (int it = 0; it < 8; ++it) x[it] = *reinterpret_cast<const float4 *>(xbase + xoff[it] + c0);
    };
    float4 xa[8], xb[8];
    load_x(xa, col_begin);
    mbar_wait(bar, parity);
    tc_fence_after();
    uint32_t rn[32];
    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
    tmem_ld_32x32(taddr + (uint32_t)col_begin, rn);
    auto step = [&](int c0, float4 (&xcur)[8], float4 (&xnext)[8]) {
        const int n0 = n_base + c0;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 g4 = __ldg(reinterpret_cast<const float4 *>(p.gamma + n0 + cc));
        if (p.bias) b4 = __ldg(reinterpret_cast<const float4 *>(p.bias + n0 + cc));
        if (c0 + 32 < col_end) load_x(xnext, c0 + 32);
        tmem_ld_wait();
        __syncwarp();                                  // the previous chunk's reads of the staging tile are done
#pragma unroll
        for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4 *>(stage + lane * 32 + ((c ^ (lane & 7)) << 2)) = make_uint4(rn[4 * c], rn[4 * c + 1], rn[4 * c + 2], rn[4 * c + 3]);
        if (c0 + 32 < col_end) tmem_ld_32x32(taddr + (uint32_t)(c0 + 32), rn);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int rr = 4 * it + sub;
            const float4 a = *reinterpret_cast<const float4 *>(stage + rr * 32 + (((lane & 7) ^ (rr & 7)) << 2));
            if (!((okmask >> it) & 1u)) continue;
            float4 x = xcur[it];
            x.x = fmaf(g4.x, a.x + b4.x, x.x); x.y = fmaf(g4.y, a.y + b4.y, x.y);
            x.z = fmaf(g4.z, a.z + b4.z, x.z); x.w = fmaf(g4.w, a.w + b4.w, x.w);
            *reinterpret_cast<float4 *>(p.X + n_base + cc + xoff[it] + c0) = x;
        }
    };
#pragma unroll 1
    for (int c0 = col_begin; c0 < col_end; c0 += 64) {
        step(c0, xa, xb);
        if (c0 + 32 < col_end) step(c0 + 32, xb, xa);
    }
}

template <int BN, bool CONV>
__global__ void __launch_bounds__(192, 2) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, GemmParams p) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw;
    float *staging = reinterpret_cast<float *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
    uint64_t *empty = full + Cfg::STAGES;
    uint64_t *tmem_full = empty + Cfg::STAGES;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blk = blockIdx.x, n_blk = blockIdx.y;
    const int num_kb = CONV ? 9 * (p.cCin / Cfg::BK) : p.K / Cfg::BK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // conv tile coordinates
    int cb = 0, cy0 = 0, cx0 = 0;
    if (CONV) {
        const int tiles_per_img = p.tiles_x * p.tiles_y;
        cb = m_blk / tiles_per_img;
        const int t = m_blk % tiles_per_img;
        cy0 = (t / p.tiles_x) * p.hbox;
        cx0 = (t % p.tiles_x) * p.wbox;
    }

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t *sa = smem + stage * Cfg::STAGE_BYTES, *sb = sa + Cfg::A_BYTES;
                mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                if (CONV) {
                    const int cblks = p.cCin / Cfg::BK;
                    const int tap = kb / cblks, cblk = kb % cblks;
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    tma_load_4d(sa, &tmA, &full[stage], cblk * Cfg::BK, cx0 + dx, cy0 + dy, cb);
                } else {
                    tma_load_2d(sa, &tmA, &full[stage], kb * Cfg::BK, m_blk * Cfg::BM);
                }
                tma_load_2d(sb, &tmB, &full[stage], kb * Cfg::BK, n_blk * BN);
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(Cfg::BM, BN, 0, 0, 0);
            int stage = 0, phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES), sb = sa + Cfg::A_BYTES;
                const uint64_t adesc = make_desc_kmajor_sw128(sa), bdesc = make_desc_kmajor_sw128(sb);
#pragma unroll
                for (int k = 0; k < Cfg::BK / 16; ++k)
                    umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                umma_commit(&empty[stage]);
                if (kb == num_kb - 1) umma_commit(tmem_full);
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue: warps 2..5; TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int row = q * 32 + lane;  // row inside the tile == TMEM lane
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        long long m;      // logical output row
        bool row_ok;
        if (CONV) {
            const int ly = row / p.wbox, lx = row % p.wbox;
            const int y = cy0 + ly, x = cx0 + lx;
            row_ok = (y < p.cH) && (x < p.cW);
            m = ((long long)cb * p.cH + y) * p.cW + x;
        } else {
            m = (long long)m_blk * Cfg::BM + row;
            row_ok = m < p.M;
        }
        if (p.epi == EPI_HEAD) epilogue_rows<BN>(p, tmem_base, q, m, row_ok, n_blk * BN);
        else epilogue_rows_staged<BN>(p, tmem_base, q, lane, m, row_ok, n_blk * BN, staging + q * 1024);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// tile id -> (m_blk, n_blk).  Tiles are ordered in groups of `group` m-tiles; inside a group all n panels are swept with
// m fastest.  The CTAs running concurrently therefore share ONE weight panel, and the group's A rows (group x 128 x K
// fp16, e.g. 148 x 256 KB = 37 MB for K = 1024) stay L2-resident across the n sweep, so A is streamed from HBM once
// instead of once per n panel (A of the fc1 GEMM at B=64 is 180 MB, larger than the 126 MB L2).  `group` is chosen by
// the launcher (pick_tile_group) so that one group's A rows stay under ~40 MB: for K = 4096 (fc2) a group of one m-tile
// per CTA would be 155 MB and every n panel would re-stream A from HBM; with a smaller group the concurrent CTAs span
// several n panels of the same few m-tiles instead.
__device__ __forceinline__ void tile_coords(int tile, int num_m_tiles, int n_tiles, int group, int &m_blk, int &n_blk) {
    const int per_group = group * n_tiles;
    const int g = tile / per_group;
    const int rem = tile - g * per_group;
    const int m0 = g * group;
    const int gl = (num_m_tiles - m0) < group ? (num_m_tiles - m0) : group;
    n_blk = rem / gl;
    m_blk = m0 + (rem - n_blk * gl);
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant for N % 256 == 0: one CTA per SM walks 128 x 256 output tiles (m fastest inside an n panel so the
// CTAs running concurrently share one weight panel in L2).  UMMA 128x256x16 halves the smem operand traffic per FLOP
// relative to 128x128 (A 4 KB + B 8 KB per 128 tensor cycles = 96 B/clk, under the 128 B/clk smem port), and two TMEM
// accumulators (2 x 256 columns = all 512) let the epilogue of tile i overlap the MMAs of tile i+1.
// ---------------------------------------------------------------------------------------------------------------------
struct PersistCfg {
    static constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGING_BYTES = 8 * 4096;   // 8 epilogue warps
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 256;
    static constexpr int TMEM_COLS = 512;
};

constexpr int PERSIST_THREADS = 64 + 256;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

template <bool CONV, int SPEC>
__global__ void __launch_bounds__(PERSIST_THREADS, 1) gemm_tcgen05_persist_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                     const __grid_constant__ CUtensorMap tmB, GemmParams p,
                                                                     int num_m_tiles, int num_tiles) {
    using Cfg = PersistCfg;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw;
    float *staging = reinterpret_cast<float *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
    uint64_t *empty = full + Cfg::STAGES;
    uint64_t *tmem_full = empty + Cfg::STAGES;   // [2]
    uint64_t *tmem_empty = tmem_full + 2;        // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = CONV ? 9 * (p.cCin / Cfg::BK) : p.K / Cfg::BK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 256); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int m_blk, n_blk;
                tile_coords(tile, num_m_tiles, num_tiles / num_m_tiles, p.tile_group, m_blk, n_blk);
                int cb = 0, cy0 = 0, cx0 = 0;
                if (CONV) {
                    const int tiles_per_img = p.tiles_x * p.tiles_y;
                    cb = m_blk / tiles_per_img;
                    const int t = m_blk % tiles_per_img;
                    cy0 = (t / p.tiles_x) * p.hbox;
                    cx0 = (t % p.tiles_x) * p.wbox;
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t *sa = smem + stage * Cfg::STAGE_BYTES, *sb = sa + Cfg::A_BYTES;
                    mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                    if (CONV) {
                        const int cblks = p.cCin / Cfg::BK;
                        const int tap = kb / cblks, cblk = kb % cblks;
                        tma_load_4d(sa, &tmA, &full[stage], cblk * Cfg::BK, cx0 + tap % 3 - 1, cy0 + tap / 3 - 1, cb);
                    } else {
                        tma_load_2d(sa, &tmA, &full[stage], kb * Cfg::BK, m_blk * Cfg::BM);
                    }
                    tma_load_2d(sb, &tmB, &full[stage], kb * Cfg::BK, n_blk * Cfg::BN);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(Cfg::BM, Cfg::BN, 0, 0, 0);
            int stage = 0, phase = 0, it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * Cfg::BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES), sb = sa + Cfg::A_BYTES;
                    const uint64_t adesc = make_desc_kmajor_sw128(sa), bdesc = make_desc_kmajor_sw128(sb);
#pragma unroll
                    for (int k = 0; k < Cfg::BK / 16; ++k)
                        umma_f16(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    umma_commit(&empty[stage]);
                    if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // 8 epilogue warps: TMEM lane quarter = warp % 4 (hardware rule); warps 2-5 take columns [0,128), 6-9 [128,256)
        const int q = warp & 3;
        const int ehalf = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            int m_blk, n_blk;
            tile_coords(tile, num_m_tiles, num_tiles / num_m_tiles, p.tile_group, m_blk, n_blk);
            long long m;
            bool row_ok;
            if (CONV) {
                const int tiles_per_img = p.tiles_x * p.tiles_y;
                const int cb = m_blk / tiles_per_img;
                const int t = m_blk % tiles_per_img;
                const int y = (t / p.tiles_x) * p.hbox + row / p.wbox, x = (t % p.tiles_x) * p.wbox + row % p.wbox;
                row_ok = (y < p.cH) && (x < p.cW);
                m = ((long long)cb * p.cH + y) * p.cW + x;
            } else {
                m = (long long)m_blk * Cfg::BM + row;
                row_ok = m < p.M;
            }
            if (SPEC == 3) {     // instantiated separately: the residual-stream epilogue does not share registers with the fp16 one
                epilogue_resid_staged<Cfg::BN>(p, tmem_base + (uint32_t)(acc * Cfg::BN), q, lane, m, row_ok, n_blk * Cfg::BN,
                                               staging + (warp - 2) * 1024, ehalf * 128, ehalf * 128 + 128, &tmem_full[acc], (it >> 1) & 1);
            } else {
                mbar_wait(&tmem_full[acc], (it >> 1) & 1);
                tc_fence_after();
                epilogue_rows_staged<Cfg::BN, SPEC == 3 ? 0 : SPEC>(p, tmem_base + (uint32_t)(acc * Cfg::BN), q, lane, m, row_ok, n_blk * Cfg::BN,
                                                                    staging + (warp - 2) * 1024, ehalf * 128, ehalf * 128 + 128);
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// 2-SM variant (cta_group::2): a cluster of two CTAs (one TPC) computes one 256 x 256 tile.  Each CTA stages only ITS
// 128 rows of A and ITS 128 of the 256 weight rows (16 KB + 16 KB per k-block instead of 16 + 32), the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256) and the hardware feeds both tensor cores from both shared memories.  Per SM that
// is 64 KB of smem traffic (32 written by TMA, 32 read by the MMA) per 512 tensor cycles = the 128 B/clk port limit,
// versus 96 KB (1.5x over the port) for the single-CTA 128 x 256 kernel, which measured ~1.05 PFLOP/s for that reason.
//   barriers: TMA of BOTH CTAs completes on the leader's full[s]; tcgen05.commit multicasts to both CTAs' empty[s] and
//   tmem_full[a]; epilogue warps of both CTAs arrive (remote mbarrier.arrive) on the leader's tmem_empty[a].
// ---------------------------------------------------------------------------------------------------------------------
struct PairCfg {
    static constexpr int BM = 128, BN = 256, BNH = 128, BK = 64, STAGES = 6;
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BNH * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGING_BYTES = 8 * 4096;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 256;
    static constexpr int TMEM_COLS = 512;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> CTA 0 of the pair
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar) {   // arrives on `bar` of BOTH CTAs of the pair
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {   // arrive on CTA 0's copy of `bar`
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

template <bool CONV, int SPEC>
__global__ void __launch_bounds__(PERSIST_THREADS, 1) gemm_tcgen05_2sm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                             const __grid_constant__ CUtensorMap tmB, GemmParams p,
                                                                             int num_m_pairs, int num_tiles) {
    using Cfg = PairCfg;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw;
    float *staging = reinterpret_cast<float *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
    uint64_t *empty = full + Cfg::STAGES;
    uint64_t *tmem_full = empty + Cfg::STAGES;   // [2]
    uint64_t *tmem_empty = tmem_full + 2;        // [2]  (the leader's copy is the one that is waited on)
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int num_kb = CONV ? 9 * (p.cCin / Cfg::BK) : p.K / Cfg::BK;
    const int n_tiles = num_tiles / num_m_pairs;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 16); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_ptr, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                int mp, n_blk;
                tile_coords(tile, num_m_pairs, n_tiles, p.tile_group, mp, n_blk);
                const int m_blk = mp * 2 + (int)rank;
                int cb = 0, cy0 = 0, cx0 = 0;
                if (CONV) {
                    const int tiles_per_img = p.tiles_x * p.tiles_y;
                    cb = m_blk / tiles_per_img;
                    const int t = m_blk % tiles_per_img;
                    cy0 = (t / p.tiles_x) * p.hbox;
                    cx0 = (t % p.tiles_x) * p.wbox;
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t *sa = smem + stage * Cfg::STAGE_BYTES, *sb = sa + Cfg::A_BYTES;
                    if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);   // bytes of both CTAs
                    if (CONV) {
                        const int cblks = p.cCin / Cfg::BK;
                        const int tap = kb / cblks, cblk = kb % cblks;
                        tma_load_4d_2sm(sa, &tmA, &full[stage], cblk * Cfg::BK, cx0 + tap % 3 - 1, cy0 + tap / 3 - 1, cb);
                    } else {
                        tma_load_2d_2sm(sa, &tmA, &full[stage], kb * Cfg::BK, m_blk * Cfg::BM);
                    }
                    tma_load_2d_2sm(sb, &tmB, &full[stage], kb * Cfg::BK, n_blk * Cfg::BN + (int)rank * Cfg::BNH);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc_f16(2 * Cfg::BM, Cfg::BN, 0, 0, 0);
            int stage = 0, phase = 0, it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * Cfg::BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES), sb = sa + Cfg::A_BYTES;
                    const uint64_t adesc = make_desc_kmajor_sw128(sa), bdesc = make_desc_kmajor_sw128(sb);
#pragma unroll
                    for (int k = 0; k < Cfg::BK / 16; ++k)
                        umma_f16_2sm(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    umma_commit_2sm(&empty[stage]);
                    if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[acc]);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        const int q = warp & 3;
        const int ehalf = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        int it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            const int acc = it & 1;
            int mp, n_blk;
            tile_coords(tile, num_m_pairs, n_tiles, p.tile_group, mp, n_blk);
            const int m_blk = mp * 2 + (int)rank;
            long long m;
            bool row_ok;
            if (CONV) {
                const int tiles_per_img = p.tiles_x * p.tiles_y;
                const int cb = m_blk / tiles_per_img;
                const int t = m_blk % tiles_per_img;
                const int y = (t / p.tiles_x) * p.hbox + row / p.wbox, x = (t % p.tiles_x) * p.wbox + row % p.wbox;
                row_ok = (cb < p.cB) && (y < p.cH) && (x < p.cW);
                m = ((long long)cb * p.cH + y) * p.cW + x;
            } else {
                m = (long long)m_blk * Cfg::BM + row;
                row_ok = m < p.M;
            }
            if (SPEC == 3) {     // instantiated separately: the residual-stream epilogue does not share registers with the fp16 one
                epilogue_resid_staged<Cfg::BN>(p, tmem_base + (uint32_t)(acc * Cfg::BN), q, lane, m, row_ok, n_blk * Cfg::BN,
                                               staging + (warp - 2) * 1024, ehalf * 128, ehalf * 128 + 128, &tmem_full[acc], (it >> 1) & 1);
            } else {
                mbar_wait(&tmem_full[acc], (it >> 1) & 1);
                tc_fence_after();
                epilogue_rows_staged<Cfg::BN, SPEC == 3 ? 0 : SPEC>(p, tmem_base + (uint32_t)(acc * Cfg::BN), q, lane, m, row_ok, n_blk * Cfg::BN,
                                                                    staging + (warp - 2) * 1024, ehalf * 128, ehalf * 128 + 128);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
        fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 2-D fp16 row-major [rows, cols] with row pitch ld (elements); box = box_rows x 64 cols, 128B swizzle
int make_tmap_2d(CUtensorMap *tm, const void *ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return DM_E_CUDA; }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(2d) failed: %d (rows=%llu cols=%llu ld=%llu)", (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld); return DM_E_CUDA; }
    return DM_OK;
}

// 4-D fp16 NHWC [B, H, W, C]; box = [1, hbox, wbox, 64]
int make_tmap_nhwc(CUtensorMap *tm, const void *ptr, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t hbox, uint32_t wbox) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return DM_E_CUDA; }
    cuuint64_t dims[4] = {C, W, H, B};
    cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    cuuint32_t box[4] = {64, wbox, hbox, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(4d) failed: %d", (int)r); return DM_E_CUDA; }
    return DM_OK;
}

template <int BN, bool CONV>
static int launch_gemm(const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmParams &p, int m_tiles, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    static PerDeviceFlag configured;
    if (!configured.test_and_set())
        DM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    dim3 grid(m_tiles, (p.N + BN - 1) / BN);
    gemm_tcgen05_kernel<BN, CONV><<<grid, 192, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
    DM_LAUNCH_CHECK("gemm_tcgen05_kernel");
    return DM_OK;
}

static int pick_bn(int N) { return N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32); }

static int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// m-tiles (of `rows` rows each) per L2 group: A rows of one group (rows x K fp16 each) within ~40 MB, at least 8, at most
// one per concurrently running CTA / cluster.  Implicit-conv tiles re-read the same pixels for all 9 taps, so their
// footprint is rows x Cin, not rows x K.
static int pick_tile_group(const GemmParams &p, int rows, int concurrent, bool conv) {
    const long long bytes = (long long)rows * (conv ? p.cCin : p.K) * 2;
    long long g = (40ll << 20) / (bytes > 0 ? bytes : 1);
    if (g < 8) g = 8;
    if (g > concurrent) g = concurrent;
    return (int)g;
}

// Epilogue specialisation of the persistent kernels (plain GEMMs only): 3 = residual stream (EPI_RESID_F32, X within
// 32-bit offsets), 1 = fp16 store + bias, 2 = fp16 store + bias + GELU, 0 = everything else (runtime modes)
static int epilogue_spec(const GemmParams &p, bool conv) {
    if (conv) return 0;
    if (p.epi == EPI_RESID_F32 && (long long)p.M * p.ldx < (1ll << 31)) return 3;
    if (p.epi == EPI_STORE_F16 && p.bias && !p.R && !p.R2 && !p.C2) {
        if (p.act == ACT_NONE) return 1;
        if (p.act == ACT_GELU) return 2;
    }
    return 0;
}

template <bool CONV, int SPEC>
static int launch_persist_t(const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmParams &p, int m_tiles, cudaStream_t stream) {
    static PerDeviceFlag configured;
    if (!configured.test_and_set())
        DM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_persist_kernel<CONV, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, PersistCfg::SMEM_BYTES));
    const int n_tiles = p.N / PersistCfg::BN;
    const int tiles = m_tiles * n_tiles;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    GemmParams q = p;
    q.tile_group = pick_tile_group(p, PersistCfg::BM, grid, CONV);
    gemm_tcgen05_persist_kernel<CONV, SPEC><<<grid, PERSIST_THREADS, PersistCfg::SMEM_BYTES, stream>>>(tmA, tmB, q, m_tiles, tiles);
    DM_LAUNCH_CHECK("gemm_tcgen05_persist_kernel");
    return DM_OK;
}
template <bool CONV>
static int launch_persist(const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmParams &p, int m_tiles, cudaStream_t stream) {
    switch (epilogue_spec(p, CONV)) {
        case 3: return launch_persist_t<false, 3>(tmA, tmB, p, m_tiles, stream);
        case 2: return launch_persist_t<false, 2>(tmA, tmB, p, m_tiles, stream);
        case 1: return launch_persist_t<false, 1>(tmA, tmB, p, m_tiles, stream);
        default: return launch_persist_t<CONV, 0>(tmA, tmB, p, m_tiles, stream);
    }
}

template <bool CONV, int SPEC>
static int launch_2sm_t(const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmParams &p, int m_tiles, cudaStream_t stream) {
    static PerDeviceFlag configured;
    if (!configured.test_and_set())
        DM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_2sm_kernel<CONV, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg::SMEM_BYTES));
    const int m_pairs = (m_tiles + 1) / 2;
    const int n_tiles = p.N / PairCfg::BN;
    const int tiles = m_pairs * n_tiles;
    int clusters = num_sms() / 2;
    if (tiles < clusters) clusters = tiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(PERSIST_THREADS);
    cfg.dynamicSmemBytes = PairCfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GemmParams q = p;
    q.tile_group = pick_tile_group(p, 2 * PairCfg::BM, clusters, CONV);
    DM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tcgen05_2sm_kernel<CONV, SPEC>, tmA, tmB, q, m_pairs, tiles));
    return DM_OK;
}
template <bool CONV>
static int launch_2sm(const CUtensorMap &tmA, const CUtensorMap &tmB, const GemmParams &p, int m_tiles, cudaStream_t stream) {
    switch (epilogue_spec(p, CONV)) {
        case 3: return launch_2sm_t<false, 3>(tmA, tmB, p, m_tiles, stream);
        case 2: return launch_2sm_t<false, 2>(tmA, tmB, p, m_tiles, stream);
        case 1: return launch_2sm_t<false, 1>(tmA, tmB, p, m_tiles, stream);
        default: return launch_2sm_t<CONV, 0>(tmA, tmB, p, m_tiles, stream);
    }
}

static bool use_2sm(const GemmParams &p, int m_tiles) {
    static int env = -1;
    if (env < 0) { const char *e = getenv("DEPTHMAP_B200_2SM"); env = (e && e[0] == '0') ? 0 : 1; }
    return env && p.epi != EPI_HEAD && p.N % 256 == 0 && (long long)((m_tiles + 1) / 2) * (p.N / 256) >= 37;
}

static bool use_persist(const GemmParams &p, int m_tiles) {
    static int env = -1;
    if (env < 0) { const char *e = getenv("DEPTHMAP_B200_NO_PERSIST"); env = (e && e[0] == '1') ? 0 : 1; }
    return env && p.epi != EPI_HEAD && p.N % 256 == 0 && (long long)m_tiles * (p.N / 256) >= 64;
}

// Plain GEMM: A fp16 [M, K] (pitch lda), W fp16 [N, K] (pitch ldw)
int gemm_f16(const __half *A, int lda, const __half *W, int ldw, GemmParams p, cudaStream_t stream) {
    if (p.K % 64 != 0 || p.N % 32 != 0) { set_error("gemm_f16: K must be a multiple of 64 and N of 32 (K=%d N=%d)", p.K, p.N); return DM_E_INVALID; }
    if ((lda % 8) || (ldw % 8)) { set_error("gemm_f16: row pitches must be multiples of 8 elements"); return DM_E_INVALID; }
    const int m_tiles = (p.M + 127) / 128;
    const bool pair = use_2sm(p, m_tiles);
    const bool persist = !pair && use_persist(p, m_tiles);
    const int bn = pair ? 128 : (persist ? 256 : ((p.epi == EPI_HEAD) ? p.N : pick_bn(p.N)));
    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d(&tmA, A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)lda, 128, 64);
    if (rc) return rc;
    rc = make_tmap_2d(&tmB, W, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)ldw, (uint32_t)bn, 64);
    if (rc) return rc;
    if (pair) return launch_2sm<false>(tmA, tmB, p, m_tiles, stream);
    if (persist) return launch_persist<false>(tmA, tmB, p, m_tiles, stream);
    switch (bn) {
        case 128: return launch_gemm<128, false>(tmA, tmB, p, m_tiles, stream);
        case 64: return launch_gemm<64, false>(tmA, tmB, p, m_tiles, stream);
        case 32: return launch_gemm<32, false>(tmA, tmB, p, m_tiles, stream);
    }
    set_error("gemm_f16: unsupported N tile %d", bn);
    return DM_E_UNSUPPORTED;
}

// 3x3 stride-1 pad-1 convolution, NHWC fp16 activations [B,H,W,Cin], weights fp16 [Cout, 9*Cin] ordered (ky, kx, cin)
int conv3x3_f16(const __half *act, int B, int H, int W, int Cin, const __half *Wt, GemmParams p, cudaStream_t stream) {
    if (Cin % 64 != 0 || p.N % 32 != 0) { set_error("conv3x3_f16: Cin must be a multiple of 64 and Cout of 32"); return DM_E_INVALID; }
    // tile = hbox x wbox pixels = 128 rows; choose the wbox in {128,64,32,16,8} with the least padding waste
    int best_w = 128; double best_eff = -1;
    for (int wb = 128; wb >= 8; wb >>= 1) {
        const int hb = 128 / wb;
        const double eff = (double)(W * H) / ((double)((W + wb - 1) / wb * wb) * ((H + hb - 1) / hb * hb));
        if (eff > best_eff + 1e-9) { best_eff = eff; best_w = wb; }
    }
    p.wbox = best_w; p.hbox = 128 / best_w;
    p.tiles_x = (W + p.wbox - 1) / p.wbox; p.tiles_y = (H + p.hbox - 1) / p.hbox;
    p.cB = B; p.cH = H; p.cW = W; p.cCin = Cin;
    p.M = B * H * W; p.K = 9 * Cin;
    const int m_tiles = B * p.tiles_x * p.tiles_y;
    const bool pair = use_2sm(p, m_tiles);
    const bool persist = !pair && use_persist(p, m_tiles);
    const int bn = pair ? 128 : (persist ? 256 : ((p.epi == EPI_HEAD) ? p.N : pick_bn(p.N)));
    CUtensorMap tmA, tmB;
    int rc = make_tmap_nhwc(&tmA, act, B, H, W, Cin, p.hbox, p.wbox);
    if (rc) return rc;
    rc = make_tmap_2d(&tmB, Wt, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K, (uint32_t)bn, 64);
    if (rc) return rc;
    if (pair) return launch_2sm<true>(tmA, tmB, p, m_tiles, stream);
    if (persist) return launch_persist<true>(tmA, tmB, p, m_tiles, stream);
    switch (bn) {
        case 128: return launch_gemm<128, true>(tmA, tmB, p, m_tiles, stream);
        case 64: return launch_gemm<64, true>(tmA, tmB, p, m_tiles, stream);
        case 32: return launch_gemm<32, true>(tmA, tmB, p, m_tiles, stream);
    }
    set_error("conv3x3_f16: unsupported N tile %d", bn);
    return DM_E_UNSUPPORTED;
}

}  // namespace dm

// ---- C-ABI test / building-block entry points ---------------------------------------------------------------------
extern "C" __attribute__((visibility("default"))) int dm_gemm_f16(const void *A, int lda, const void *W, int ldw, const float *bias, void *C,
                                                               int ldc, int M, int N, int K, int act, int out_f32, void *stream) {
    using namespace dm;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K; p.act = act; p.bias = bias;
    if (out_f32) { p.epi = EPI_STORE_F32; p.X = (float *)C; p.ldx = ldc; }
    else { p.epi = EPI_STORE_F16; p.C = (__half *)C; p.ldc = ldc; }
    return gemm_f16((const __half *)A, lda, (const __half *)W, ldw, p, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int dm_conv3x3_f16(const void *act, int B, int H, int W, int Cin, const void *Wt, const float *bias,
                                                                  void *out, int Cout, int relu, void *stream) {
    using namespace dm;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.N = Cout; p.act = relu ? ACT_RELU : ACT_NONE; p.bias = bias; p.epi = EPI_STORE_F16; p.C = (__half *)out; p.ldc = Cout;
    return conv3x3_f16((const __half *)act, B, H, W, Cin, (const __half *)Wt, p, (cudaStream_t)stream);
}

static void desc_to_params(const dm_gemm_desc *d, dm::GemmParams &p) {
    memset(&p, 0, sizeof(p));
    p.M = d->M; p.N = d->N; p.K = d->K; p.epi = d->epi; p.act = d->act; p.bias = d->bias;
    p.C = (__half *)d->C; p.ldc = d->ldc; p.C2 = (__half *)d->C2;
    p.R = (const __half *)d->R; p.ldr = d->ldr; p.R2 = (const __half *)d->R2; p.ldr2 = d->ldr2;
    p.X = d->X; p.ldx = d->ldx; p.gamma = d->gamma; p.head_b2 = d->head_b2;
    p.ps_s = d->ps_s; p.ps_cout = d->ps_cout; p.ps_h = d->ps_h; p.ps_w = d->ps_w;
}

extern "C" __attribute__((visibility("default"))) int dm_gemm_ex(const void *A, int lda, const void *W, int ldw, const dm_gemm_desc *d, void *stream) {
    if (!A || !W || !d) { dm::set_error("dm_gemm_ex: null argument"); return DM_E_INVALID; }
    dm::GemmParams p;
    desc_to_params(d, p);
    return dm::gemm_f16((const __half *)A, lda, (const __half *)W, ldw, p, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int dm_conv3x3_ex(const void *act, int B, int H, int W, int Cin, const void *Wt, const dm_gemm_desc *d, void *stream) {
    if (!act || !Wt || !d) { dm::set_error("dm_conv3x3_ex: null argument"); return DM_E_INVALID; }
    dm::GemmParams p;
    desc_to_params(d, p);
    return dm::conv3x3_f16((const __half *)act, B, H, W, Cin, (const __half *)Wt, p, (cudaStream_t)stream);
}
