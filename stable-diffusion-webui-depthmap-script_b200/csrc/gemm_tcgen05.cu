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
};

template <int BN>
struct GemmCfg {
    static constexpr int BM = 128, BK = 64;
    static constexpr int STAGES = BN >= 256 ? 3 : (BN == 128 ? 3 : 4);
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN, bool CONV>
__global__ void __launch_bounds__(192) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, GemmParams p) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
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
        float head_acc = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            const int n0 = n_blk * BN + c0;
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
                for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
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
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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
    static bool configured = false;
    if (!configured) {
        DM_CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    dim3 grid(m_tiles, (p.N + BN - 1) / BN);
    gemm_tcgen05_kernel<BN, CONV><<<grid, 192, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
    DM_LAUNCH_CHECK("gemm_tcgen05_kernel");
    return DM_OK;
}

static int pick_bn(int N) { return N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32); }

// Plain GEMM: A fp16 [M, K] (pitch lda), W fp16 [N, K] (pitch ldw)
int gemm_f16(const __half *A, int lda, const __half *W, int ldw, GemmParams p, cudaStream_t stream) {
    if (p.K % 64 != 0 || p.N % 32 != 0) { set_error("gemm_f16: K must be a multiple of 64 and N of 32 (K=%d N=%d)", p.K, p.N); return DM_E_INVALID; }
    if ((lda % 8) || (ldw % 8)) { set_error("gemm_f16: row pitches must be multiples of 8 elements"); return DM_E_INVALID; }
    const int bn = (p.epi == EPI_HEAD) ? p.N : pick_bn(p.N);
    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d(&tmA, A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)lda, 128, 64);
    if (rc) return rc;
    rc = make_tmap_2d(&tmB, W, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)ldw, (uint32_t)bn, 64);
    if (rc) return rc;
    const int m_tiles = (p.M + 127) / 128;
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
    const int bn = (p.epi == EPI_HEAD) ? p.N : pick_bn(p.N);
    CUtensorMap tmA, tmB;
    int rc = make_tmap_nhwc(&tmA, act, B, H, W, Cin, p.hbox, p.wbox);
    if (rc) return rc;
    rc = make_tmap_2d(&tmB, Wt, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K, (uint32_t)bn, 64);
    if (rc) return rc;
    const int m_tiles = B * p.tiles_x * p.tiles_y;
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
