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

#include "common.cuh"
#include "tc_common.cuh"

namespace dm {
using namespace tc;

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

__global__ void __launch_bounds__(192) attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, AttnParams p) {
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
                mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
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
            mbar_wait(q_full, 0);
            const uint64_t qdesc = make_desc_kmajor_sw128(smem_u32(sQ));
            const uint32_t sp = smem_u32(sP);
            for (int j = 0; j < num_kv; ++j) {
                const int s = j & 1;
                mbar_wait(&kv_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t sk = smem_u32(sKV + s * 2 * AT_KV_BYTES), sv = sk + AT_KV_BYTES;
                const uint64_t kdesc = make_desc_kmajor_sw128(sk);
#pragma unroll
                for (int k = 0; k < AT_D / 16; ++k) umma_f16(tmem_S, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_qk, k != 0);
                umma_commit(s_full);
                mbar_wait(p_full, j & 1);
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
        const __half *brow = p.bias ? p.bias + ((size_t)h * p.N + (qi < p.N ? qi : 0)) * p.bias_ld : nullptr;
        constexpr float LOG2E = 1.4426950408889634f;

        for (int j = 0; j < num_kv; ++j) {
            const int kbase = j * AT_BKV;
            const int nvalid = min(AT_BKV, p.N - kbase);
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            // ---- pass 1: row max of (scale*S + bias) in the log2 domain ----
            float mx = -INFINITY;
#pragma unroll 1
            for (int c0 = 0; c0 < AT_BKV; c0 += 32) {
                if (c0 >= nvalid) break;
                uint32_t r[32];
                tmem_ld_32x32(tmem_S + lane_off + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float sv = __uint_as_float(r[i]) * p.scale_log2e;
                    if (brow) sv = fmaf(__half2float(__ldg(brow + kbase + c0 + i)), LOG2E, sv);
                    if (c0 + i < nvalid) mx = fmaxf(mx, sv);
                }
            }
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
            const float alpha = exp2f(m_run - m_new);
            l_run *= alpha;
#pragma unroll
            for (int d = 0; d < AT_D; ++d) o[d] *= alpha;
            m_run = m_new;
            // ---- pass 2: P = exp2(s - m), fp16, into the swizzled K-major smem tile ----
            float lsum = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < AT_BKV; c0 += 32) {
                uint32_t r[32];
                if (c0 < nvalid) {
                    tmem_ld_32x32(tmem_S + lane_off + c0, r);
                    tmem_ld_wait();
                }
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = 0.f, p1 = 0.f;
                    if (c0 + i < nvalid) {
                        float sv = __uint_as_float(r[i]) * p.scale_log2e;
                        if (brow) sv = fmaf(__half2float(__ldg(brow + kbase + c0 + i)), LOG2E, sv);
                        p0 = exp2f(sv - m_new);
                    }
                    if (c0 + i + 1 < nvalid) {
                        float sv = __uint_as_float(r[i + 1]) * p.scale_log2e;
                        if (brow) sv = fmaf(__half2float(__ldg(brow + kbase + c0 + i + 1)), LOG2E, sv);
                        p1 = exp2f(sv - m_new);
                    }
                    const __half2 h2 = __floats2half2_rn(p0, p1);
                    // the sum must match what the MMA sees: accumulate the fp16-rounded probabilities
                    const float2 f2 = __half22float2(h2);
                    lsum += f2.x + f2.y;
                    packed[i >> 1] = *reinterpret_cast<const uint32_t *>(&h2);
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

int attention_f16(const __half *qkv, const AttnParams &p, cudaStream_t stream) {
    if (p.C != p.H * AT_D) { set_error("attention_f16: head_dim must be 64"); return DM_E_UNSUPPORTED; }
    CUtensorMap tm;
    int rc = make_tmap_2d(&tm, qkv, (uint64_t)p.B * p.N, (uint64_t)3 * p.C, (uint64_t)3 * p.C, AT_BQ, AT_D);
    if (rc) return rc;
    static bool configured = false;
    if (!configured) {
        DM_CUDA_CHECK(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        configured = true;
    }
    dim3 grid((p.N + AT_BQ - 1) / AT_BQ, p.H, p.B);
    attention_fwd_kernel<<<grid, 192, AT_SMEM, stream>>>(tm, p);
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
