// gemm_tc.cuh -- tcgen05 tensor-core GEMM for prefill / ViT / audio Linear layers on sm_100a:
//     C[M,N] = A[M,K] (fp32 activations) x W[N,K]^T (fp16 weights)  with fp32-grade accuracy.
//
// The 5th-gen tensor cores take fp16 operands, but the parity budget (1e-3 on logits) does not survive rounding
// activations to fp16.  So A is split once into A = hi + lo (both fp16, |lo| <= 2^-11 |hi|) and
// every weight tile is multiplied by both halves into the SAME fp32 TMEM accumulator: the weight products are
// exact (fp16 x fp16 -> fp32) and the dropped remainder of A is ~2^-22 relative.  Twice the MMA work, still
// tensor-pipe bound instead of CUDA-core bound.
//
// Structure (one 128x128 output tile per CTA, 192 threads, two CTAs resident per SM):
//   warp 0  : TMA producer -- cp.async.bulk.tensor.2d (SWIZZLE_128B) of A_hi, A_lo, W k-blocks (128 rows x 64 halfs
//             = 16 KB each) into a 2-stage shared-memory ring, completion on `full` mbarriers;
//   warp 1  : MMA issuer -- one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M128 N128 K16), 8 per
//             k-block (4 k-steps x {hi, lo}), accumulator = 128 lanes x 128 fp32 columns of TMEM; tcgen05.commit
//             releases the ring slot, and after the last k-block signals the epilogue;
//   warps 2-5: epilogue -- tcgen05.ld (32x32b.x32) TMEM -> registers, fused bias / residual / activation / SwiGLU,
//             fp32 stores.  Warp 2 also owns the TMEM allocation.
// Replaces candle Linear::forward at the reference call sites /root/reference/src/models/common/modules.rs:81-87,
// 538-577 and qwen3vl/model.rs:96-103,168-184,232-278 (prefill shapes); validated against gemm_simt.cuh.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "decode_fused.cuh"  // mbarrier helpers
#include "gemm_simt.cuh"

namespace aha {

#ifndef AHA_TC_STAGES
#define AHA_TC_STAGES 2      // 2-stage ring and TWO co-resident CTAs per SM (96 KB + 128 TMEM columns each): one CTA's epilogue runs under the
#endif                       // other's MMAs.  Measured against the 4-stage / 1-CTA build: 431 vs 398 TFLOP/s useful on 2560x4096x2048, VL2 prefill 76.6 vs 82.0 ms
constexpr int kTcBM = 128, kTcBN = 128, kTcBK = 64, kTcStages = AHA_TC_STAGES;
constexpr int kTcCtasPerSm = kTcStages <= 2 ? 2 : 1;
constexpr int kTcTileBytes = kTcBM * kTcBK * 2;             // 16 KB
constexpr int kTcStageBytes = 3 * kTcTileBytes;             // A_hi, A_lo, W
constexpr int kTcThreads = 192;

// ------------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_tmapEncodeTiled tmap_encode_fn() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        AHA_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        AHA_REQUIRE(q == cudaDriverEntryPointSuccess && p, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}
// fp16 row-major [rows, K] matrix, box = 128 rows x 64 halfs (128 bytes), 128-byte swizzle, OOB rows read as zero
inline CUtensorMap make_tmap_f16(const void* ptr, uint64_t rows, uint64_t K) {
    CUtensorMap m;
    const cuuint64_t gdim[2] = {K, rows};
    const cuuint64_t gstride[1] = {K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)kTcBM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = tmap_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AHA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4, LBO=1 (unused for
// swizzled K-major), SBO = 1024 B (8 rows x 128 B) >> 4, version 1 (Blackwell), layout_type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem_ptr) {
    const uint32_t a = smem_u32(smem_ptr);
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D=f32 (bit 4), A=B=f16 (0), both K-major, N>>3 at bit 17, M>>4 at bit 24.
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
          "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------------ A = hi + lo split
__global__ void split_f32_to_f16x2_kernel(const float* __restrict__ a, int lda, __half* __restrict__ hi, __half* __restrict__ lo, int M, int K) {
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // index of a float4 inside [M, K]
    const size_t per_row = K / 4;
    if (i4 >= (size_t)M * per_row) return;
    const size_t r = i4 / per_row, c = (i4 % per_row) * 4;
    const float4 v = *reinterpret_cast<const float4*>(a + r * lda + c);
    const float x[4] = {v.x, v.y, v.z, v.w};
    __half h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        h[e] = __float2half_rn(x[e]);
        l[e] = __float2half_rn(x[e] - __half2float(h[e]));
    }
    *reinterpret_cast<uint2*>(hi + r * K + c) = make_uint2(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[2]));
    *reinterpret_cast<uint2*>(lo + r * K + c) = make_uint2(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[2]));
}

// ------------------------------------------------------------------------------------------------ the GEMM kernel
struct GemmTcArgs {
    const float* bias;
    const float* resid; int ldr;
    float* C; int ldc;
    int M, N, K;
    int act;
    __half* out_hi = nullptr;   // when set, the result (after bias / activation / SwiGLU) is written as fp16 hi + lo halves [M, ldo] for the
    __half* out_lo = nullptr;   // next GEMM instead of fp32 C (EPI_STORE / EPI_ACT / EPI_SWIGLU)
    int ldo = 0;
    int group_m = 8;   // persistent kernel: tiles are walked in bands of group_m row blocks (n outer, m inner) so that one wave of CTAs shares few A and W tiles
};

// 8 floats -> 8 fp16 hi + 8 fp16 lo (one 16-byte store each)
__device__ __forceinline__ void tc_store_split8(const float* v, __half* hi, __half* lo) {
    __half h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        h[e] = __float2half_rn(v[e]);
        l[e] = __float2half_rn(v[e] - __half2float(h[e]));
    }
    *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(l);
}
// one thread's 32 consecutive accumulator columns [n, n + 32) of output row `row`: bias / residual / activation / SwiGLU, then fp32
// stores -- or, when g.out_hi is set, the fp16 hi + lo halves the next GEMM reads
template <int EPI>
__device__ __forceinline__ void tc_epilogue_store(const GemmTcArgs& g, float (&v)[32], int row, int n) {
    if (row >= g.M || n >= g.N) return;
    if (g.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += g.bias[n + j];
    }
    if (EPI == EPI_SWIGLU) {
        float r[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = silu_f(v[2 * j]) * v[2 * j + 1];
        if (g.out_hi) {
            const size_t o = (size_t)row * g.ldo + n / 2;
            tc_store_split8(r, g.out_hi + o, g.out_lo + o);
            tc_store_split8(r + 8, g.out_hi + o + 8, g.out_lo + o + 8);
        } else {
            float* out = g.C + (size_t)row * g.ldc + n / 2;
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
        return;
    }
    if (EPI == EPI_ACT) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = apply_act(g.act, v[j]);
    }
    if (EPI == EPI_RESID) {
        const float* rr = g.resid + (size_t)row * g.ldr + n;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 r4 = *reinterpret_cast<const float4*>(rr + j);
            v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
        }
    }
    if (EPI != EPI_RESID && g.out_hi) {
        const size_t o = (size_t)row * g.ldo + n;
#pragma unroll
        for (int j = 0; j < 32; j += 8) tc_store_split8(v + j, g.out_hi + o + j, g.out_lo + o + j);
        return;
    }
    float* out = g.C + (size_t)row * g.ldc + n;
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}

template <int EPI>
__global__ void __launch_bounds__(kTcThreads, kTcCtasPerSm) gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                                                               const __grid_constant__ CUtensorMap tm_w, GemmTcArgs g) {
    extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment in the shared window: align by hand (1 KB of slack is allocated)
    uint8_t* tiles = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);   // [stages][A_hi | A_lo | W]
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)kTcStages * kTcStageBytes);
    uint64_t* empty = full + kTcStages;
    uint64_t* acc_full = empty + kTcStages;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * kTcBM, n0 = blockIdx.x * kTcBN;
    const int nkb = g.K / kTcBK;

    if (tid == 0) {
        for (int i = 0; i < kTcStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // TMEM: 128 fp32 accumulator columns, allocated by one full warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "n"(kTcBN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_holder;

    if (warp == 0) {
        // ======================= TMA producer
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kTcStages;
                mbar_wait(&empty[s], ((kb / kTcStages) & 1) ^ 1);
                uint8_t* st = tiles + (size_t)s * kTcStageBytes;
                mbar_expect_tx(&full[s], kTcStageBytes);
                tma_load_2d(st, &tm_hi, kb * kTcBK, m0, &full[s]);
                tma_load_2d(st + kTcTileBytes, &tm_lo, kb * kTcBK, m0, &full[s]);
                tma_load_2d(st + 2 * kTcTileBytes, &tm_w, kb * kTcBK, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer
        const uint32_t idesc = umma_idesc_f16(kTcBM, kTcBN);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kTcStages;
            mbar_wait(&full[s], (kb / kTcStages) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint8_t* st = tiles + (size_t)s * kTcStageBytes;
                const uint64_t d_hi = umma_desc_sw128(st), d_lo = umma_desc_sw128(st + kTcTileBytes), d_w = umma_desc_sw128(st + 2 * kTcTileBytes);
#pragma unroll
                for (int ks = 0; ks < kTcBK / 16; ++ks) {
                    const uint64_t adv = (uint64_t)((ks * 16 * 2) >> 4);   // 32 bytes per K=16 step inside the 128-byte swizzle atom
                    umma_f16(tmem_acc, d_hi + adv, d_w + adv, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                    umma_f16(tmem_acc, d_lo + adv, d_w + adv, idesc, 1u);
                }
                umma_commit(&empty[s]);                  // frees the ring slot when these MMAs have read it
                if (kb == nkb - 1) umma_commit(acc_full);  // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ======================= epilogue warps 2..5: TMEM lane group = warp % 4
        const int lg = warp & 3;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int row = m0 + lg * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < kTcBN; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_acc + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, v);
            tc_epilogue_store<EPI>(g, v, row, n0 + c0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(kTcBN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ persistent 128 x 256 variant
// The 128 x 128 kernel above moves 48 KB through L2 -> SM per 2 x 128 x 128 x 64 MACs: at the measured 431 TFLOP/s useful that is
// 9.8 TB/s of the ~12 TB/s the L2 can deliver (B300_MICROARCH: LTS cap ~6300 B/clk) -- the kernel is L2-bandwidth-bound, not
// tensor-bound.  This variant makes the tile 128 x 256 (64 KB per 2 x 128 x 256 x 64 MACs: 1.5x the flops per byte), runs ONE
// persistent CTA per SM over a strided list of tiles, and keeps TWO 256-column accumulators in TMEM so that the epilogue warps drain
// tile i while the MMA warp is already accumulating tile i + 1 (acc_full / acc_empty mbarriers).
constexpr int kTc2BN = 256, kTc2Stages = 3;
constexpr int kTc2WBytes = kTc2BN * kTcBK * 2;                       // 32 KB
constexpr int kTc2StageBytes = 2 * kTcTileBytes + kTc2WBytes;        // A_hi, A_lo, W: 64 KB

// tile index -> (row block, column block): bands of `gm` row blocks, inside a band the column block is the slow index
__device__ __forceinline__ void tc2_tile(int t, int tiles_m, int tiles_n, int gm, int& mb, int& nb) {
    const int band = t / (gm * tiles_n), r = t - band * gm * tiles_n;
    const int rows = min(gm, tiles_m - band * gm);
    nb = r / rows;
    mb = band * gm + (r - nb * rows);
}

template <int EPI>
__global__ void __launch_bounds__(kTcThreads, 1) gemm_tc2_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                                                                const __grid_constant__ CUtensorMap tm_w, GemmTcArgs g) {
    extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
    uint8_t* tiles = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);   // [stages][A_hi | A_lo | W]
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)kTc2Stages * kTc2StageBytes);
    uint64_t* empty = full + kTc2Stages;
    uint64_t* acc_full = empty + kTc2Stages;       // [2]
    uint64_t* acc_empty = acc_full + 2;            // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = g.K / kTcBK;
    const int tiles_n = (g.N + kTc2BN - 1) / kTc2BN, tiles_m = (g.M + kTcBM - 1) / kTcBM, ntiles = tiles_n * tiles_m;

    if (tid == 0) {
        for (int i = 0; i < kTc2Stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }   // 4 epilogue warps release an accumulator
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // all 512 TMEM columns: two 128 x 256 fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ======================= TMA producer: one continuous ring across all tiles of this CTA
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                int mb, nb;
                tc2_tile(t, tiles_m, tiles_n, g.group_m, mb, nb);
                const int m0 = mb * kTcBM, n0 = nb * kTc2BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % kTc2Stages;
                    mbar_wait(&empty[s], ((it / kTc2Stages) & 1) ^ 1);
                    uint8_t* st = tiles + (size_t)s * kTc2StageBytes;
                    mbar_expect_tx(&full[s], kTc2StageBytes);
                    tma_load_2d(st, &tm_hi, kb * kTcBK, m0, &full[s]);
                    tma_load_2d(st + kTcTileBytes, &tm_lo, kb * kTcBK, m0, &full[s]);
                    tma_load_2d(st + 2 * kTcTileBytes, &tm_w, kb * kTcBK, n0, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer
        const uint32_t idesc = umma_idesc_f16(kTcBM, kTc2BN);
        int it = 0, ti = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
            const int acc = ti & 1;
            mbar_wait(&acc_empty[acc], (((ti >> 1) & 1) ^ 1));   // the epilogue has drained this accumulator (first use passes at once)
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * kTc2BN);
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % kTc2Stages;
                mbar_wait(&full[s], (it / kTc2Stages) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint8_t* st = tiles + (size_t)s * kTc2StageBytes;
                    const uint64_t d_hi = umma_desc_sw128(st), d_lo = umma_desc_sw128(st + kTcTileBytes), d_w = umma_desc_sw128(st + 2 * kTcTileBytes);
#pragma unroll
                    for (int ks = 0; ks < kTcBK / 16; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * 16 * 2) >> 4);
                        umma_f16(tmem_acc, d_hi + adv, d_w + adv, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                        umma_f16(tmem_acc, d_lo + adv, d_w + adv, idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                    if (kb == nkb - 1) umma_commit(&acc_full[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        // ======================= epilogue warps 2..5: TMEM lane group = warp % 4
        const int lg = warp & 3;
        int ti = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
            const int acc = ti & 1;
            int mb, nb;
            tc2_tile(t, tiles_m, tiles_n, g.group_m, mb, nb);
            const int m0 = mb * kTcBM, n0 = nb * kTc2BN;
            mbar_wait(&acc_full[acc], (ti >> 1) & 1);
            tc_fence_after();
            const int row = m0 + lg * 32 + lane;
            const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * kTc2BN);
#pragma unroll 1
            for (int c0 = 0; c0 < kTc2BN; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_acc + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, v);
                tc_epilogue_store<EPI>(g, v, row, n0 + c0);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);   // this warp's 32 TMEM lanes of the accumulator are drained
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}
inline size_t gemm_tc2_smem_bytes() { return (size_t)kTc2Stages * kTc2StageBytes + (2 * kTc2Stages + 4) * sizeof(uint64_t) + 16 + 1024; }

// ------------------------------------------------------------------------------------------------ CTA-pair (cta_group::2) variant
// Two CTAs of a cluster (the two SMs of one TPC) compute ONE 256 x 256 tile: tcgen05.mma.cta_group::2 with M = 256 takes rows
// 0-127 of A from the leader's shared memory and rows 128-255 from the peer's, and the two 128-row halves of the W tile the same
// way; each CTA's TMEM receives its own 128 rows x 256 columns of the accumulator.  Per CTA and k-block that is A_hi + A_lo
// (32 KB) + HALF a W tile (16 KB) = 48 KB for 128 x 256 x 64 x 2 MACs -- a third less L2 -> SM traffic per flop than the
// single-CTA 128 x 256 kernel and half the 128 x 128 kernel's, which is what bounds these GEMMs.  Roles per CTA as above (TMA warp,
// MMA warp -- only the leader's issues --, 4 epilogue warps); both CTAs' TMA transfers complete on the LEADER's `full` barrier
// (cp.async.bulk.tensor ... .cta_group::2), tcgen05.commit ... multicast::cluster releases the ring slot / publishes the
// accumulator in BOTH CTAs, and the peer's epilogue warps arrive remotely (mapa) on the leader's `acc_empty` barrier.
constexpr int kTc3Stages = 4;
constexpr int kTc3StageBytes = 3 * kTcTileBytes;   // per CTA: A_hi, A_lo (its 128 rows) and its 128-row half of the 256-row W tile

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes land on the mbarrier at the same offset in the LEADER CTA of the pair (bit 24 of a shared::cluster
// address selects the odd CTA of a pair; clearing it addresses the even one -- cute's Sm100MmaPeerBitMask)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar) & 0xFEFFFFFFu)
                 : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {   // arrive on the barrier at this offset in CTA `cta` of the cluster
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}

// bounded wait: a protocol error in the pair kernel traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
    gemm_tc3_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const __grid_constant__ CUtensorMap tm_w, GemmTcArgs g) {
    extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
    uint8_t* tiles = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)kTc3Stages * kTc3StageBytes);   // used in the leader only
    uint64_t* empty = full + kTc3Stages;
    uint64_t* acc_full = empty + kTc3Stages;       // [2]
    uint64_t* acc_empty = acc_full + 2;            // [2], used in the leader only: 4 epilogue warps x 2 CTAs
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int nkb = g.K / kTcBK;
    const int tiles_n = (g.N + 255) / 256, tiles_m = (g.M + 255) / 256, ntiles = tiles_n * tiles_m;

    if (tid == 0) {
        for (int i = 0; i < kTc3Stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // the same warp of both CTAs allocates: all 512 columns = two 128 x 256 fp32 accumulators per CTA
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ======================= TMA producer (both CTAs): my 128 rows of A, my 128-row half of the W tile
        if (lane == 0) {
            int it = 0;
            for (int t = pair; t < ntiles; t += npairs) {
                int mb, nb;
                tc2_tile(t, tiles_m, tiles_n, g.group_m, mb, nb);
                const int m0 = mb * 256 + (int)rank * 128, n0 = nb * 256 + (int)rank * 128;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % kTc3Stages;
                    mbar_wait_or_trap(&empty[s], ((it / kTc3Stages) & 1) ^ 1);
                    uint8_t* st = tiles + (size_t)s * kTc3StageBytes;
                    if (rank == 0) mbar_expect_tx(&full[s], 2 * kTc3StageBytes);   // both CTAs' bytes complete on the leader's barrier
                    tma_load_2d_pair(st, &tm_hi, kb * kTcBK, m0, &full[s]);
                    tma_load_2d_pair(st + kTcTileBytes, &tm_lo, kb * kTcBK, m0, &full[s]);
                    tma_load_2d_pair(st + 2 * kTcTileBytes, &tm_w, kb * kTcBK, n0, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer: the leader CTA only
        if (rank == 0) {
            const uint32_t idesc = umma_idesc_f16(256, 256);
            int it = 0, ti = 0;
            for (int t = pair; t < ntiles; t += npairs, ++ti) {
                const int acc = ti & 1;
                mbar_wait_or_trap(&acc_empty[acc], (((ti >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * 256);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % kTc3Stages;
                    mbar_wait_or_trap(&full[s], (it / kTc3Stages) & 1);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint8_t* st = tiles + (size_t)s * kTc3StageBytes;
                        const uint64_t d_hi = umma_desc_sw128(st), d_lo = umma_desc_sw128(st + kTcTileBytes), d_w = umma_desc_sw128(st + 2 * kTcTileBytes);
#pragma unroll
                        for (int ks = 0; ks < kTcBK / 16; ++ks) {
                            const uint64_t adv = (uint64_t)((ks * 16 * 2) >> 4);
                            umma_f16_pair(tmem_acc, d_hi + adv, d_w + adv, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                            umma_f16_pair(tmem_acc, d_lo + adv, d_w + adv, idesc, 1u);
                        }
                        umma_commit_pair(&empty[s]);
                        if (kb == nkb - 1) umma_commit_pair(&acc_full[acc]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ======================= epilogue warps 2..5 (both CTAs): my 128 rows x 256 columns of the tile
        const int lg = warp & 3;
        int ti = 0;
        for (int t = pair; t < ntiles; t += npairs, ++ti) {
            const int acc = ti & 1;
            int mb, nb;
            tc2_tile(t, tiles_m, tiles_n, g.group_m, mb, nb);
            const int m0 = mb * 256 + (int)rank * 128, n0 = nb * 256;
            mbar_wait_or_trap(&acc_full[acc], (ti >> 1) & 1);
            tc_fence_after();
            const int row = m0 + lg * 32 + lane;
            const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * 256);
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_acc + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, v);
                tc_epilogue_store<EPI>(g, v, row, n0 + c0);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&acc_empty[acc], 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the leader's MMAs and commits touch the peer's shared memory: nobody leaves early
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}
inline size_t gemm_tc3_smem_bytes() { return (size_t)kTc3Stages * kTc3StageBytes + (2 * kTc3Stages + 4) * sizeof(uint64_t) + 16 + 1024; }

inline size_t gemm_tc_smem_bytes() { return (size_t)kTcStages * kTcStageBytes + (2 * kTcStages + 1) * sizeof(uint64_t) + 16 + 1024; }

inline bool gemm_tc_supported(int M, int N, int K) { return K % kTcBK == 0 && N % 32 == 0 && M >= 1; }

// The > 48 KB dynamic shared memory opt-in is a per-device function attribute: aha_b200_create calls this after
// cudaSetDevice for every handle (a process-wide once-flag would leave a second device without it).
inline void gemm_tc_init() {
    const int smem = (int)gemm_tc_smem_bytes();
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<EPI_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<EPI_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<EPI_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int smem2 = (int)gemm_tc2_smem_bytes();
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc2_kernel<EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc2_kernel<EPI_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc2_kernel<EPI_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc2_kernel<EPI_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    const int smem3 = (int)gemm_tc3_smem_bytes();
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc3_kernel<EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc3_kernel<EPI_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc3_kernel<EPI_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc3_kernel<EPI_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
}

// A already split: hi/lo fp16 [M, K] row-major.  tm_w describes W [N, K].
inline void gemm_tc_launch(cudaStream_t st, int epi, const __half* a_hi, const __half* a_lo, const CUtensorMap& tm_w, const GemmTcArgs& g) {
    const CUtensorMap tm_hi = make_tmap_f16(a_hi, (uint64_t)g.M, (uint64_t)g.K);
    const CUtensorMap tm_lo = make_tmap_f16(a_lo, (uint64_t)g.M, (uint64_t)g.K);
    const size_t smem = gemm_tc_smem_bytes();
    dim3 grid(ceil_div(g.N, kTcBN), ceil_div(g.M, kTcBM));
#define AHA_TC_CASE(E)                                                                                                             \
    case E: gemm_tc_kernel<E><<<grid, kTcThreads, smem, st>>>(tm_hi, tm_lo, tm_w, g); break;
    switch (epi) {
        AHA_TC_CASE(EPI_STORE)
        AHA_TC_CASE(EPI_RESID)
        AHA_TC_CASE(EPI_ACT)
        AHA_TC_CASE(EPI_SWIGLU)
        default: AHA_REQUIRE(false, "gemm_tc: bad epilogue");
    }
#undef AHA_TC_CASE
    AHA_CUDA_CHECK(cudaGetLastError());
}

// persistent 128 x 256 launch: tm_w must have been built with 256-row boxes (make_tmap_f16_rows(w, N, K, 256))
inline CUtensorMap make_tmap_f16_rows(const void* ptr, uint64_t rows, uint64_t K, uint32_t box_rows) {
    CUtensorMap m;
    const cuuint64_t gdim[2] = {K, rows};
    const cuuint64_t gstride[1] = {K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kTcBK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = tmap_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AHA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}
inline void gemm_tc2_launch(cudaStream_t st, int epi, const __half* a_hi, const __half* a_lo, const CUtensorMap& tm_w256, const GemmTcArgs& g, int num_sms) {
    const CUtensorMap tm_hi = make_tmap_f16(a_hi, (uint64_t)g.M, (uint64_t)g.K);
    const CUtensorMap tm_lo = make_tmap_f16(a_lo, (uint64_t)g.M, (uint64_t)g.K);
    const size_t smem = gemm_tc2_smem_bytes();
    const int ntiles = ceil_div(g.N, kTc2BN) * ceil_div(g.M, kTcBM);
    dim3 grid(std::min(ntiles, num_sms));
#define AHA_TC2_CASE(E) case E: gemm_tc2_kernel<E><<<grid, kTcThreads, smem, st>>>(tm_hi, tm_lo, tm_w256, g); break;
    switch (epi) {
        AHA_TC2_CASE(EPI_STORE)
        AHA_TC2_CASE(EPI_RESID)
        AHA_TC2_CASE(EPI_ACT)
        AHA_TC2_CASE(EPI_SWIGLU)
        default: AHA_REQUIRE(false, "gemm_tc2: bad epilogue");
    }
#undef AHA_TC2_CASE
    AHA_CUDA_CHECK(cudaGetLastError());
}

// CTA-pair launch: tm_w is the ordinary 128-row-box map (each CTA of a pair loads its own 128-row half of the 256-row W tile)
inline void gemm_tc3_launch(cudaStream_t st, int epi, const __half* a_hi, const __half* a_lo, const CUtensorMap& tm_w, const GemmTcArgs& g, int num_sms) {
    const CUtensorMap tm_hi = make_tmap_f16(a_hi, (uint64_t)g.M, (uint64_t)g.K);
    const CUtensorMap tm_lo = make_tmap_f16(a_lo, (uint64_t)g.M, (uint64_t)g.K);
    const size_t smem = gemm_tc3_smem_bytes();
    const int ntiles = ceil_div(g.N, 256) * ceil_div(g.M, 256);
    dim3 grid(2 * std::min(ntiles, num_sms / 2));
#define AHA_TC3_CASE(E) case E: gemm_tc3_kernel<E><<<grid, kTcThreads, smem, st>>>(tm_hi, tm_lo, tm_w, g); break;
    switch (epi) {
        AHA_TC3_CASE(EPI_STORE)
        AHA_TC3_CASE(EPI_RESID)
        AHA_TC3_CASE(EPI_ACT)
        AHA_TC3_CASE(EPI_SWIGLU)
        default: AHA_REQUIRE(false, "gemm_tc3: bad epilogue");
    }
#undef AHA_TC3_CASE
    AHA_CUDA_CHECK(cudaGetLastError());
}

}  // namespace aha
