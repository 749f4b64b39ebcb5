// Fused multi-head attention forward for the ViT backbones (head_dim 64, N <= a few thousand tokens, no mask):
//   O = softmax(scale * Q K^T [+ rel_pos_bias]) V
// Replaces the materialised attention of the reference (dinov2_layers/attention.py:49-62; dmidas/backbones/beit.py:65-91,
// which writes a [B,16,N,N] tensor per block and rebuilds the relative-position bias every forward).
//
// One CTA per (128-query tile, head, image); two CTAs are resident per SM so one CTA's softmax overlaps the other's MMAs.
//   warp 0    TMA producer: Q tile once, then K_j / V_j tiles (128 keys x 64) into a 2-stage ring, straight out of the
//             packed qkv activation [B*N, 3C] (no head split / transpose pass)
//   warp 1    MMA issuer:  S = Q K_j^T   (tcgen05.mma 128x128x16, K-major A and B, fp32 S in TMEM)
//                          PV = P_j V_j  (128x64x16; A = P from smem, B = V_j used MN-major so V needs no transpose)
//   warps 2-5 softmax: thread = query row.  Pass 1 reads S from TMEM for the row max, pass 2 re-reads it, exponentiates
//             (exp2 with folded log2e), writes P as fp16 into the 128B-swizzled K-major smem tile the PV MMA consumes,
//             keeps the running (max, sum) and the fp32 output row O in registers (O = O*alpha + PV_j from TMEM).
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dm {
using namespace tc;

// single-thread role warps (TMA / MMA issuer) wait with a short sleep so their polling does not steal issue slots
// from the softmax warps that share their SM sub-partitions
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(20);
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

int make_tmap_2d(CUtensorMap *tm, const void *ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols);

struct AttnParams {
    int B, N, H;            // images, tokens per image, heads
    int C;                  // H * 64
    float scale_log2e;      // softmax scale * log2(e)
    const __half *bias;     // optional [H, N, bias_ld] additive bias (already multiplied by nothing; natural-log domain)
    int bias_ld;
    __half *out;            // [B*N, C]
};

constexpr int AT_BQ = 128, AT_BKV = 128, AT_D = 64;
constexpr int AT_Q_BYTES = AT_BQ * AT_D * 2;        // 16 KB
constexpr int AT_KV_BYTES = AT_BKV * AT_D * 2;      // 16 KB each for K and V
constexpr int AT_P_BYTES = AT_BQ * AT_BKV * 2;      // 32 KB (two 64-key swizzle atoms)
constexpr int AT_STAGES = 2;
constexpr int AT_SMEM = AT_Q_BYTES + AT_STAGES * 2 * AT_KV_BYTES + AT_P_BYTES + 256;  // base is __align__(1024); 2 CTAs fit one SM
constexpr int AT_TMEM_COLS = 256;                   // S: cols [0,128), PV staging: cols [128,192)

template <bool HAS_BIAS>
__global__ void __launch_bounds__(192, 2) attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sQ = smem;
    uint8_t *sKV = sQ + AT_Q_BYTES;                       // stage s: K at sKV + s*32K, V at +16K
    uint8_t *sP = sKV + AT_STAGES * 2 * AT_KV_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sP + AT_P_BYTES);
    uint64_t *q_full = bars, *kv_full = bars + 1, *kv_empty = bars + 3, *s_full = bars + 5, *p_full = bars + 6, *pv_full = bars + 7;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * AT_BQ;
    const int num_kv = (p.N + AT_BKV - 1) / AT_BKV;
    const int row_base = b * p.N;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < AT_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        mbar_init(s_full, 1);
        mbar_init(p_full, 128);
        mbar_init(pv_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, AT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S = tmem_base, tmem_PV = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, AT_Q_BYTES);
            tma_load_2d(sQ, &tmQKV, q_full, h * AT_D, row_base + q0);
            for (int j = 0; j < num_kv; ++j) {
                const int s = j & 1;
                mbar_wait_backoff(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                uint8_t *sk = sKV + s * 2 * AT_KV_BYTES, *sv = sk + AT_KV_BYTES;
                mbar_arrive_expect_tx(&kv_full[s], 2 * AT_KV_BYTES);
                tma_load_2d(sk, &tmQKV, &kv_full[s], p.C + h * AT_D, row_base + j * AT_BKV);
                tma_load_2d(sv, &tmQKV, &kv_full[s], 2 * p.C + h * AT_D, row_base + j * AT_BKV);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_f16(AT_BQ, AT_BKV, 0, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_f16(AT_BQ, AT_D, 0, 0, 1);  // B (= V) is MN-major
            mbar_wait_backoff(q_full, 0);
            const uint64_t qdesc = make_desc_kmajor_sw128(smem_u32(sQ));
            const uint32_t sp = smem_u32(sP);
            for (int j = 0; j < num_kv; ++j) {
                const int s = j & 1;
                mbar_wait_backoff(&kv_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t sk = smem_u32(sKV + s * 2 * AT_KV_BYTES), sv = sk + AT_KV_BYTES;
                const uint64_t kdesc = make_desc_kmajor_sw128(sk);
#pragma unroll
                for (int k = 0; k < AT_D / 16; ++k) umma_f16(tmem_S, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_qk, k != 0);
                umma_commit(s_full);
                mbar_wait_backoff(p_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < AT_BKV / 16; ++k) {
                    const uint64_t pdesc = make_desc_kmajor_sw128(sp + (k >> 2) * (AT_BQ * 128) + (k & 3) * 32);
                    const uint64_t vdesc = make_desc_mnmajor_sw128(sv + k * 16 * 128, 16 * 128);
                    umma_f16(tmem_PV, pdesc, vdesc, idesc_pv, k != 0);
                }
                umma_commit(pv_full);
                umma_commit(&kv_empty[s]);
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int qi = q0 + row;                       // query index inside the image
        float m_run = -INFINITY, l_run = 0.f;
        float o[AT_D];
#pragma unroll
        for (int d = 0; d < AT_D; ++d) o[d] = 0.f;
        const __half *brow = HAS_BIAS ? p.bias + ((size_t)h * p.N + (qi < p.N ? qi : 0)) * p.bias_ld : nullptr;
        constexpr float LOG2E = 1.4426950408889634f;

        for (int j = 0; j < num_kv; ++j) {
            const int kbase = j * AT_BKV;
            const int nvalid = min(AT_BKV, p.N - kbase);
            const bool full_tile = nvalid == AT_BKV;
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            // ---- pass 1: row max (log2 domain).  Without bias max(scale*s) = scale*max(s) since scale > 0 ----
            float mx = -INFINITY;
#pragma unroll 1
            for (int c0 = 0; c0 < AT_BKV; c0 += 32) {
                if (c0 >= nvalid) break;
                uint32_t r[32];
                tmem_ld_32x32(tmem_S + lane_off + c0, r);
                tmem_ld_wait();
                if (HAS_BIAS) {
                    const uint4 *bp = reinterpret_cast<const uint4 *>(brow + kbase + c0);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint4 u = __ldg(bp + g);
                        const __half2 *h2 = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 bf = __half22float2(h2[k]);
                            const int i = g * 8 + 2 * k;
                            float s0 = fmaf(bf.x, LOG2E, __uint_as_float(r[i]) * p.scale_log2e);
                            float s1 = fmaf(bf.y, LOG2E, __uint_as_float(r[i + 1]) * p.scale_log2e);
                            if (!full_tile) { s0 = (c0 + i < nvalid) ? s0 : -INFINITY; s1 = (c0 + i + 1 < nvalid) ? s1 : -INFINITY; }
                            mx = fmaxf(mx, fmaxf(s0, s1));
                        }
                    }
                } else if (full_tile) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c0 + i < nvalid) ? __uint_as_float(r[i]) : -INFINITY);
                }
            }
            if (!HAS_BIAS) mx *= p.scale_log2e;
            // ---- fold in PV of the previous tile, then rescale ----
            if (j > 0) {
                mbar_wait(pv_full, (j - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < AT_D; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_PV + lane_off + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[c0 + i] += __uint_as_float(r[i]);
                }
            }
            const float m_new = fmaxf(m_run, mx);
            const float alpha = ex2_approx(m_run - m_new);
            l_run *= alpha;
#pragma unroll
            for (int d = 0; d < AT_D; ++d) o[d] *= alpha;
            m_run = m_new;
            // ---- pass 2: P = exp2(scale*s [+ bias] - m), fp16, into the swizzled K-major smem tile ----
            float lsum = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < AT_BKV; c0 += 32) {
                uint32_t packed[16];
                if (c0 < nvalid) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_S + lane_off + c0, r);
                    tmem_ld_wait();
                    float pv[32];
                    if (HAS_BIAS) {
                        const uint4 *bp = reinterpret_cast<const uint4 *>(brow + kbase + c0);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const uint4 u = __ldg(bp + g);
                            const __half2 *h2 = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 bf = __half22float2(h2[k]);
                                const int i = g * 8 + 2 * k;
                                pv[i] = ex2_approx(fmaf(bf.x, LOG2E, fmaf(__uint_as_float(r[i]), p.scale_log2e, -m_new)));
                                pv[i + 1] = ex2_approx(fmaf(bf.y, LOG2E, fmaf(__uint_as_float(r[i + 1]), p.scale_log2e, -m_new)));
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) pv[i] = ex2_approx(fmaf(__uint_as_float(r[i]), p.scale_log2e, -m_new));
                    }
                    if (!full_tile) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) pv[i] = (c0 + i < nvalid) ? pv[i] : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const __half2 h2 = __floats2half2_rn(pv[i], pv[i + 1]);
                        // the sum must match what the MMA sees: accumulate the fp16-rounded probabilities
                        const float2 f2 = __half22float2(h2);
                        lsum += f2.x + f2.y;
                        packed[i >> 1] = *reinterpret_cast<const uint32_t *>(&h2);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) packed[i] = 0u;
                }
                // 32 keys = 64 B = four 16-byte chunks of atom (c0 / 64), chunk index ((c0 % 64) / 8 + t) ^ (row % 8)
                uint8_t *atom = sP + (c0 >> 6) * (AT_BQ * 128) + row * 128;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int chunk = (((c0 & 63) >> 3) + t) ^ (row & 7);
                    *reinterpret_cast<uint4 *>(atom + chunk * 16) = make_uint4(packed[4 * t], packed[4 * t + 1], packed[4 * t + 2], packed[4 * t + 3]);
                }
            }
            l_run += lsum;
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(p_full);
        }
        // ---- last PV, normalise, store ----
        mbar_wait(pv_full, (num_kv - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < AT_D; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_PV + lane_off + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[c0 + i] += __uint_as_float(r[i]);
        }
        if (qi < p.N) {
            const float inv = 1.0f / l_run;
            __half *dst = p.out + (size_t)(row_base + qi) * p.C + h * AT_D;
#pragma unroll
            for (int d = 0; d < AT_D; d += 8) {
                uint4 u;
                __half2 *h2 = reinterpret_cast<__half2 *>(&u);
#pragma unroll
                for (int k = 0; k < 4; ++k) h2[k] = __floats2half2_rn(o[d + 2 * k] * inv, o[d + 2 * k + 1] * inv);
                *reinterpret_cast<uint4 *>(dst + d) = u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, AT_TMEM_COLS);
}


// =====================================================================================================================
// Ping-pong variant (the default): one CTA owns TWO 128-query tiles of one (head, image) and runs two softmax
// warpgroups.  While warpgroup A exponentiates S_A(j) the tensor core computes S_B(j) = Q_B K_j^T and P_A V_j, so the
// MUFU-bound softmax of one tile hides behind the MMAs of the other, and every SM sub-partition hosts two softmax
// warps whose TMEM / shared-memory latencies overlap.  K_j / V_j are fetched once for both query tiles.
//   warp 0: TMA producer | warp 1: MMA issuer | warps 2-5: softmax group A | warps 6-9: softmax group B
// TMEM (512 columns): S_A [0,128) S_B [128,256) PV_A [256,320) PV_B [320,384).
// Bias modes: 0 none (DINOv2) | 1 dense fp16 [H,N,ld] | 2 BEiT relative-position table: the per-head table
// [nrd] (pre-multiplied by log2 e) and the per-key offset ky*(2gw-1)+kx live in shared memory and the bias of (q, k) is
// table[base_q - koff_k] — no [H,N,N] tensor is ever read (the reference materialises it per block per forward).
// =====================================================================================================================
struct Attn2Params {
    AttnParams a;
    const float *rel_table;   // [H, nrd] * log2(e)   (mode 2)
    const float *rel_rowmax;  // [H, N]: max_k bias(q, k) * log2(e)   (mode 2) — lets pass 1 skip the bias gather
    int nrd, gh, gw;
};

constexpr int A2_THREADS = 64 + 256;
constexpr int A2_Q_BYTES = 2 * AT_Q_BYTES;                 // 32 KB
constexpr int A2_KV_BYTES = AT_STAGES * 2 * AT_KV_BYTES;   // 64 KB
constexpr int A2_P_BYTES = 2 * AT_P_BYTES;                 // 64 KB
constexpr int A2_TAB_BYTES = 16384 + 4096;                 // fp32 table (<= 4096 entries) + u16 key offsets (<= 2048 keys)
constexpr int A2_SMEM_BASE = A2_Q_BYTES + A2_KV_BYTES + A2_P_BYTES + 256;

template <int BIAS_MODE>
__global__ void __launch_bounds__(A2_THREADS, 1) attention_fwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, Attn2Params pp) {
    const AttnParams &p = pp.a;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *sQ = smem_raw;                                  // tile A at +0, tile B at +16 KB
    uint8_t *sKV = sQ + A2_Q_BYTES;
    uint8_t *sP = sKV + A2_KV_BYTES;                         // P_A at +0, P_B at +32 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(sP + A2_P_BYTES);
    uint64_t *q_full = bars, *kv_full = bars + 1, *kv_empty = bars + 3;
    uint64_t *s_full = bars + 5, *p_full = bars + 7, *pv_full = bars + 9;   // [2] each, index = group
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 11);
    float *s_tab = reinterpret_cast<float *>(smem_raw + A2_SMEM_BASE);
    uint16_t *s_koff = reinterpret_cast<uint16_t *>(smem_raw + A2_SMEM_BASE + 16384);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qp = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qp * 2 * AT_BQ;
    const bool b_active = q0 + AT_BQ < p.N;                   // second tile may be entirely out of range
    const int num_kv = (p.N + AT_BKV - 1) / AT_BKV;
    const int row_base = b * p.N;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < AT_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int g = 0; g < 2; ++g) { mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 128); mbar_init(&pv_full[g], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    if (BIAS_MODE == 2) {
        const float *tab = pp.rel_table + (size_t)h * pp.nrd;
        for (int i = threadIdx.x; i < pp.nrd; i += A2_THREADS) s_tab[i] = __ldg(tab + i);
        for (int k = threadIdx.x; k < num_kv * AT_BKV; k += A2_THREADS) {
            const int t = k - 1;
            s_koff[k] = (k >= 1 && k < p.N) ? (uint16_t)((t / pp.gw) * (2 * pp.gw - 1) + (t % pp.gw)) : (uint16_t)0;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, b_active ? 2 * AT_Q_BYTES : AT_Q_BYTES);
            tma_load_2d(sQ, &tmQKV, q_full, h * AT_D, row_base + q0);
            if (b_active) tma_load_2d(sQ + AT_Q_BYTES, &tmQKV, q_full, h * AT_D, row_base + q0 + AT_BQ);
            for (int j = 0; j < num_kv; ++j) {
                const int s = j & 1;
                mbar_wait_backoff(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                uint8_t *sk = sKV + s * 2 * AT_KV_BYTES, *sv = sk + AT_KV_BYTES;
                mbar_arrive_expect_tx(&kv_full[s], 2 * AT_KV_BYTES);
                tma_load_2d(sk, &tmQKV, &kv_full[s], p.C + h * AT_D, row_base + j * AT_BKV);
                tma_load_2d(sv, &tmQKV, &kv_full[s], 2 * p.C + h * AT_D, row_base + j * AT_BKV);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_f16(AT_BQ, AT_BKV, 0, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_f16(AT_BQ, AT_D, 0, 0, 1);  // B (= V) is MN-major
            mbar_wait_backoff(q_full, 0);
            const int ngroups = b_active ? 2 : 1;
            for (int j = 0; j < num_kv; ++j) {
                const int s = j & 1;
                mbar_wait_backoff(&kv_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t sk = smem_u32(sKV + s * 2 * AT_KV_BYTES), sv = sk + AT_KV_BYTES;
                const uint64_t kdesc = make_desc_kmajor_sw128(sk);
                for (int g = 0; g < ngroups; ++g) {   // S_g = Q_g K_j^T  (the S buffer of group g was released by p_full[g](j-1))
                    const uint64_t qdesc = make_desc_kmajor_sw128(smem_u32(sQ + g * AT_Q_BYTES));
#pragma unroll
                    for (int k = 0; k < AT_D / 16; ++k)
                        umma_f16(tmem_base + g * 128, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_qk, k != 0);
                    umma_commit(&s_full[g]);
                }
                for (int g = 0; g < ngroups; ++g) {   // PV_g = P_g V_j once group g has written P_g
                    mbar_wait_backoff(&p_full[g], j & 1);
                    tc_fence_after();
                    const uint32_t sp = smem_u32(sP + g * AT_P_BYTES);
#pragma unroll
                    for (int k = 0; k < AT_BKV / 16; ++k) {
                        const uint64_t pdesc = make_desc_kmajor_sw128(sp + (k >> 2) * (AT_BQ * 128) + (k & 3) * 32);
                        const uint64_t vdesc = make_desc_mnmajor_sw128(sv + k * 16 * 128, 16 * 128);
                        umma_f16(tmem_base + 256 + g * 64, pdesc, vdesc, idesc_pv, k != 0);
                    }
                    umma_commit(&pv_full[g]);
                }
                umma_commit(&kv_empty[s]);
            }
        }
    } else {
        const int g = (warp - 2) >> 2;            // softmax group: 0 = tile A, 1 = tile B
        if (g == 1 && !b_active) goto done;
        {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t tmem_S = tmem_base + g * 128, tmem_PV = tmem_base + 256 + g * 64;
        uint8_t *sPg = sP + g * AT_P_BYTES;
        const int qi = q0 + g * AT_BQ + row;           // query index inside the image
        float m_run = -INFINITY, l_run = 0.f;
        float o[AT_D];
#pragma unroll
        for (int d = 0; d < AT_D; ++d) o[d] = 0.f;
        constexpr float LOG2E = 1.4426950408889634f;
        const __half *brow = BIAS_MODE == 1 ? p.bias + ((size_t)h * p.N + (qi < p.N ? qi : 0)) * p.bias_ld : nullptr;
        // relative-position table addressing: idx(q, k) = base_q - mult * koff_k; the class-token query uses the
        // constant entry nrd-3 (mult = 0); the class-token key (k = 0) is patched separately below
        int rp_base = 0, rp_mult = 1;
        float rp_k0 = 0.f;                                   // bias of (q, key 0)
        float rp_rowmax = 0.f;
        if (BIAS_MODE == 2) {
            const int qq = qi < p.N ? qi : 1;
            rp_rowmax = __ldg(pp.rel_rowmax + (size_t)h * p.N + qq);
            if (qq == 0) { rp_base = pp.nrd - 3; rp_mult = 0; rp_k0 = s_tab[pp.nrd - 1]; }
            else {
                const int t = qq - 1, qy = t / pp.gw, qx = t % pp.gw;
                rp_base = (qy + pp.gh - 1) * (2 * pp.gw - 1) + (qx + pp.gw - 1);
                rp_k0 = s_tab[pp.nrd - 2];
            }
        }

        for (int j = 0; j < num_kv; ++j) {
            const int kbase = j * AT_BKV;
            const int nvalid = min(AT_BKV, p.N - kbase);
            const bool full_tile = nvalid == AT_BKV;
            mbar_wait(&s_full[g], j & 1);
            tc_fence_after();
            // ---- first tile only: a real max pass (m_run starts at -inf).  Later tiles use LAZY rescaling: P is
            // formed against the running max as it stands, the tile max is tracked on the side, and only if it exceeds
            // the running max by more than 2^8 is the tile redone with the new max (rare; keeps P inside fp16 range).
            if (j == 0) {
                float mx = -INFINITY;
#pragma unroll 1
                for (int c0 = 0; c0 < AT_BKV; c0 += 32) {
                    if (c0 >= nvalid) break;
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_S + lane_off + c0, r);
                    tmem_ld_wait();
                    if (BIAS_MODE == 1) {
                        const uint4 *bp = reinterpret_cast<const uint4 *>(brow + kbase + c0);
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) {
                            const uint4 u = __ldg(bp + gq);
                            const __half2 *h2 = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 bf = __half22float2(h2[k]);
                                const int i = gq * 8 + 2 * k;
                                float s0 = fmaf(bf.x, LOG2E, __uint_as_float(r[i]) * p.scale_log2e);
                                float s1 = fmaf(bf.y, LOG2E, __uint_as_float(r[i + 1]) * p.scale_log2e);
                                if (!full_tile) { s0 = (c0 + i < nvalid) ? s0 : -INFINITY; s1 = (c0 + i + 1 < nvalid) ? s1 : -INFINITY; }
                                mx = fmaxf(mx, fmaxf(s0, s1));
                            }
                        }
                    } else if (full_tile) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c0 + i < nvalid) ? __uint_as_float(r[i]) : -INFINITY);
                    }
                }
                if (BIAS_MODE == 0) mx *= p.scale_log2e;
                if (BIAS_MODE == 2) mx = fmaf(mx, p.scale_log2e, rp_rowmax);   // upper bound: max(scale*s) + max(bias)
                m_run = mx;
            } else {
                // fold in PV of the previous tile (computed against the same m_run: no rescale needed yet)
                mbar_wait(&pv_full[g], (j - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < AT_D; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_PV + lane_off + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[c0 + i] += __uint_as_float(r[i]);
                }
            }
            // ---- P = exp2(scale*s [+ bias] - m_run) as fp16 into the swizzled K-major smem tile ----
            float lsum, mx_tile;
            for (int attempt = 0;; ++attempt) {
                lsum = 0.f;
                mx_tile = -INFINITY;
                const float m_use = m_run;
                // 16-column sub-chunks, ping-pong TMEM loads: sub-chunk c+1 is in flight while c is exponentiated
                uint32_t rn[16];
                tmem_ld_32x16(tmem_S + lane_off, rn);
#pragma unroll 1
                for (int c0 = 0; c0 < AT_BKV; c0 += 16) {
                    float bv[16];
                    // bias first: it does not depend on S, so its loads overlap the TMEM round trip
                    if (BIAS_MODE == 1) {
                        const uint4 *bp = reinterpret_cast<const uint4 *>(brow + kbase + c0);
#pragma unroll
                        for (int gq = 0; gq < 2; ++gq) {
                            const uint4 u = __ldg(bp + gq);
                            const __half2 *h2 = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
                            for (int k = 0; k < 4; ++k) { const float2 bf = __half22float2(h2[k]); bv[gq * 8 + 2 * k] = bf.x * LOG2E; bv[gq * 8 + 2 * k + 1] = bf.y * LOG2E; }
                        }
                    } else if (BIAS_MODE == 2) {
                        const uint4 *kp = reinterpret_cast<const uint4 *>(s_koff + kbase + c0);
#pragma unroll
                        for (int gq = 0; gq < 2; ++gq) {
                            const uint4 u = kp[gq];
                            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                bv[gq * 8 + 2 * k] = s_tab[rp_base - rp_mult * (int)(w[k] & 0xffffu)];
                                bv[gq * 8 + 2 * k + 1] = s_tab[rp_base - rp_mult * (int)(w[k] >> 16)];
                            }
                        }
                        if (kbase + c0 == 0) bv[0] = rp_k0;
                    }
                    uint32_t r[16];
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = rn[i];
                    if (c0 + 16 < AT_BKV) tmem_ld_32x16(tmem_S + lane_off + c0 + 16, rn);
                    uint32_t packed[8];
                    if (c0 < nvalid) {
                        float sv[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            sv[i] = BIAS_MODE == 0 ? __uint_as_float(r[i]) * p.scale_log2e : fmaf(__uint_as_float(r[i]), p.scale_log2e, bv[i]);
                            if (!full_tile) sv[i] = (c0 + i < nvalid) ? sv[i] : -INFINITY;
                            mx_tile = fmaxf(mx_tile, sv[i]);
                        }
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const __half2 h2 = __floats2half2_rn(ex2_approx(sv[i] - m_use), ex2_approx(sv[i + 1] - m_use));
                            const float2 f2 = __half22float2(h2);   // sum what the MMA will see
                            lsum += f2.x + f2.y;
                            packed[i >> 1] = *reinterpret_cast<const uint32_t *>(&h2);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) packed[i] = 0u;
                    }
                    // 16 keys = 32 B = two 16-byte chunks of atom (c0 / 64), chunk index ((c0 % 64) / 8 + t) ^ (row % 8)
                    uint8_t *atom = sPg + (c0 >> 6) * (AT_BQ * 128) + row * 128;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int chunk = (((c0 & 63) >> 3) + t) ^ (row & 7);
                        *reinterpret_cast<uint4 *>(atom + chunk * 16) = make_uint4(packed[4 * t], packed[4 * t + 1], packed[4 * t + 2], packed[4 * t + 3]);
                    }
                }
                const bool need = mx_tile > m_run + 8.0f;
                if (!__any_sync(0xffffffffu, need)) break;
                if (need) {   // raise this row's running max and rescale what has been accumulated so far
                    const float alpha = ex2_approx(m_run - mx_tile);
                    l_run *= alpha;
#pragma unroll
                    for (int d = 0; d < AT_D; ++d) o[d] *= alpha;
                    m_run = mx_tile;
                }
            }
            l_run += lsum;
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(&p_full[g]);
        }
        // ---- last PV, normalise, store ----
        mbar_wait(&pv_full[g], (num_kv - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < AT_D; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_PV + lane_off + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[c0 + i] += __uint_as_float(r[i]);
        }
        if (qi < p.N) {
            const float inv = 1.0f / l_run;
            __half *dst = p.out + (size_t)(row_base + qi) * p.C + h * AT_D;
#pragma unroll
            for (int d = 0; d < AT_D; d += 8) {
                uint4 u;
                __half2 *h2 = reinterpret_cast<__half2 *>(&u);
#pragma unroll
                for (int k = 0; k < 4; ++k) h2[k] = __floats2half2_rn(o[d + 2 * k] * inv, o[d + 2 * k + 1] * inv);
                *reinterpret_cast<uint4 *>(dst + d) = u;
            }
        }
        }
    }
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int MODE>
static int launch_attn2(const CUtensorMap &tm, const Attn2Params &pp, cudaStream_t stream) {
    const int smem = A2_SMEM_BASE + (MODE == 2 ? A2_TAB_BYTES : 0);
    static bool configured = false;
    if (!configured) {
        DM_CUDA_CHECK(cudaFuncSetAttribute(attention_fwd2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid((pp.a.N + 2 * AT_BQ - 1) / (2 * AT_BQ), pp.a.H, pp.a.B);
    attention_fwd2_kernel<MODE><<<grid, A2_THREADS, smem, stream>>>(tm, pp);
    DM_LAUNCH_CHECK("attention_fwd2_kernel");
    return DM_OK;
}

int attention_f16(const __half *qkv, const AttnParams &p, cudaStream_t stream, const float *rel_table = nullptr, const float *rel_rowmax = nullptr,
                  int nrd = 0, int gh = 0, int gw = 0) {
    if (p.C != p.H * AT_D) { set_error("attention_f16: head_dim must be 64"); return DM_E_UNSUPPORTED; }
    CUtensorMap tm;
    int rc = make_tmap_2d(&tm, qkv, (uint64_t)p.B * p.N, (uint64_t)3 * p.C, (uint64_t)3 * p.C, AT_BQ, AT_D);
    if (rc) return rc;
    if (p.bias && (p.bias_ld % 8 != 0 || p.bias_ld < ((p.N + AT_BKV - 1) / AT_BKV) * AT_BKV)) {
        set_error("attention_f16: bias row pitch must be a multiple of 8 and cover whole 128-key tiles (got %d)", p.bias_ld);
        return DM_E_INVALID;
    }
    static int use_v1 = -1;
    if (use_v1 < 0) { const char *e = getenv("DEPTHMAP_B200_ATTN_V1"); use_v1 = (e && e[0] == '1') ? 1 : 0; }
    if (!use_v1 || rel_table) {
        Attn2Params pp;
        pp.a = p; pp.rel_table = rel_table; pp.rel_rowmax = rel_rowmax; pp.nrd = nrd; pp.gh = gh; pp.gw = gw;
        if (rel_table) {
            if (nrd > 4096 || p.N > 2048 || gh * gw + 1 != p.N) { set_error("attention_f16: relative-position table mode supports nrd <= 4096, N <= 2048, N = gh*gw+1"); return DM_E_UNSUPPORTED; }
            return launch_attn2<2>(tm, pp, stream);
        }
        return p.bias ? launch_attn2<1>(tm, pp, stream) : launch_attn2<0>(tm, pp, stream);
    }
    static bool configured = false;
    if (!configured) {
        DM_CUDA_CHECK(cudaFuncSetAttribute(attention_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        DM_CUDA_CHECK(cudaFuncSetAttribute(attention_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        configured = true;
    }
    if (p.bias && (p.bias_ld % 8 != 0 || p.bias_ld < ((p.N + AT_BKV - 1) / AT_BKV) * AT_BKV)) {
        set_error("attention_f16: bias row pitch must be a multiple of 8 and cover whole 128-key tiles (got %d)", p.bias_ld);
        return DM_E_INVALID;
    }
    dim3 grid((p.N + AT_BQ - 1) / AT_BQ, p.H, p.B);
    if (p.bias) attention_fwd_kernel<true><<<grid, 192, AT_SMEM, stream>>>(tm, p);
    else attention_fwd_kernel<false><<<grid, 192, AT_SMEM, stream>>>(tm, p);
    DM_LAUNCH_CHECK("attention_fwd_kernel");
    return DM_OK;
}

}  // namespace dm

extern "C" __attribute__((visibility("default"))) int dm_attention_f16(const void *qkv, int B, int N, int H, float scale, const void *bias, int bias_ld,
                                                                    void *out, void *stream) {
    dm::AttnParams p;
    p.B = B; p.N = N; p.H = H; p.C = H * 64;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.bias = (const __half *)bias; p.bias_ld = bias_ld;
    p.out = (__half *)out;
    return dm::attention_f16((const __half *)qkv, p, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int dm_attention_relpos_f16(const void *qkv, int B, int gh, int gw, int H, float scale,
                                                                           const float *rel_table_log2e, const float *rel_rowmax_log2e, int nrd,
                                                                           void *out, void *stream) {
    dm::AttnParams p;
    p.B = B; p.N = gh * gw + 1; p.H = H; p.C = H * 64;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.bias = nullptr; p.bias_ld = 0;
    p.out = (__half *)out;
    if (!rel_table_log2e || !rel_rowmax_log2e) { dm::set_error("dm_attention_relpos_f16: table / rowmax is NULL"); return DM_E_INVALID; }
    return dm::attention_f16((const __half *)qkv, p, (cudaStream_t)stream, rel_table_log2e, rel_rowmax_log2e, nrd, gh, gw);
}
