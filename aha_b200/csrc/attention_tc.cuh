// attention_tc.cuh -- flash attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) for head_dim 64: the ViT segments
// of Qwen3-VL (8160 tokens x 16 heads per layer for a 1080p image -- half of the whole prefill) and the audio encoder.
//
// Same contract as flash_attn_kernel / flash_attn_mma_kernel: softmax(Q K^T * scale) V per head, never materialising
// S x S (reference: Qwen3VLVisionAttention::forward, /root/reference/src/models/qwen3vl/model.rs:232-278;
// eager_attention_forward, src/models/common/modules.rs:757-813; NaiveAttention, modules.rs:201-242).
// Accuracy: like every tensor-core path here, operands are split x = hi + lo (fp16 each) and each product is formed as
// hi*hi + hi*lo + lo*hi into the same fp32 TMEM accumulator.
//
// One CTA = 128 queries of one head, KV tiles of 64 tokens, 160 threads, 2 CTAs per SM (96 KB of shared memory and 128
// TMEM columns each) so that one CTA's tensor work overlaps the other's softmax:
//   warp 4 (one elected lane): TMA loads (Q once; K and V^T tiles single-buffered, re-issued the moment the MMA that
//     read them has completed) and all tcgen05.mma: S[128 x 64] = Q K^T (12 UMMA M128 N64 K16 = 4 k-steps x 3 products)
//     into one of TWO S buffers in TMEM (S runs a tile ahead of the softmax), PV[128 x 64] = P V (12 more) into a third
//     64-column buffer, issued as soon as P is complete so that it executes under the next tile's softmax;
//     tcgen05.commit -> mbarriers;
//   warps 0-3: one thread per query row (= TMEM lane): tcgen05.ld its S row, scale / mask / online softmax in fp32, split
//     P into fp16 hi | lo and store it to shared memory in the K-major SWIZZLE_128B layout the MMA descriptors expect,
//     fence.proxy.async + arrive; later tcgen05.ld the PV row and fold it into the fp32 output row kept in registers.
// V is staged transposed ([head][d][token], written by the split pass) so that every MMA operand is K-major and the one
// descriptor form validated by gemm_tc.cuh serves all of them.
#pragma once
#include "attention_mma.cuh"
#include "gemm_tc.cuh"

namespace aha {

constexpr int kFaBQ = 128, kFaBKV = 64, kFaThreads = 160;

// fp16 row-major [rows, cols] matrix, box = box_rows x 64 halfs (128 bytes), 128-byte swizzle, OOB reads as zero
inline CUtensorMap make_tmap_f16_box(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    CUtensorMap m;
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {cols * 2};
    const cuuint32_t box[2] = {64u, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = tmap_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AHA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

// ---- split pass: q, k -> packed fp16 hi | lo [head][token][64]; v -> TRANSPOSED hi | lo [head][64][skv_pad] (zero padded) ----
struct SplitTcArgs {
    FlashArgs f;
    __half *q_hi, *q_lo, *k_hi, *k_lo, *vt_hi, *vt_lo;
    int nheads, nkv, skv_pad;
};
// blockIdx.y: 0 = q, 1 = k (one row per block, like split_qkv_kernel); 2 = v: 64-token tile per block, transposed through shared memory
template <int HD>
__global__ void split_qkv_tc_kernel(SplitTcArgs s) {
    const int which = blockIdx.y;
    if (which < 2) {
        const int S = which == 0 ? s.f.Sq : s.f.Skv, H = which == 0 ? s.nheads : s.nkv;
        // HD / 4 threads x float4 per row, 512 / HD rows per pass of the 128 threads
        constexpr int TPR = HD / 4, RPP = 128 / TPR;
        for (size_t row = (size_t)blockIdx.x * RPP + threadIdx.x / TPR; row < (size_t)S * H; row += (size_t)gridDim.x * RPP) {
            const int head = (int)(row / S), tok = (int)(row % S);
            const float* src = which == 0 ? s.f.q + (size_t)(s.f.q0 + tok) * s.f.q_tok_stride + (size_t)head * s.f.q_head_stride
                                          : s.f.kv.k + s.f.kv.off(s.f.kv0 + tok, head);
            __half* hi = (which == 0 ? s.q_hi : s.k_hi) + row * HD;
            __half* lo = (which == 0 ? s.q_lo : s.k_lo) + row * HD;
            {
                const int c4 = (threadIdx.x % TPR) * 4;
                const float4 v = *reinterpret_cast<const float4*>(src + c4);
                uint32_t h0, l0, h1, l1;
                split2(v.x, v.y, h0, l0);
                split2(v.z, v.w, h1, l1);
                *reinterpret_cast<uint2*>(hi + c4) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(lo + c4) = make_uint2(l0, l1);
            }
        }
        return;
    }
    __shared__ float tile[64][HD + 1];
    const int tiles_per_head = s.skv_pad / 64;
    for (int job = blockIdx.x; job < s.nkv * tiles_per_head; job += gridDim.x) {
        const int head = job / tiles_per_head, t0 = (job % tiles_per_head) * 64;
        __syncthreads();
        for (int idx = threadIdx.x; idx < 64 * HD; idx += blockDim.x) {
            const int r = idx / HD, c = idx % HD;
            tile[r][c] = (t0 + r < s.f.Skv) ? s.f.kv.v[s.f.kv.off(s.f.kv0 + t0 + r, head) + c] : 0.f;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < HD * 64; idx += blockDim.x) {
            const int d = idx / 64, r = idx % 64;
            const float x = tile[r][d];
            const __half h = __float2half_rn(x);
            const size_t o = ((size_t)head * HD + d) * s.skv_pad + t0 + r;
            s.vt_hi[o] = h;
            s.vt_lo[o] = __float2half_rn(x - __half2float(h));
        }
    }
}

struct FlashTcArgs {
    float* out; size_t o_tok_stride, o_head_stride;
    __half *out_hi, *out_lo;   // when set: the output goes out as fp16 hi + lo halves (operand format of the next GEMM) instead of fp32
    int Sq, Skv, q0, groups;
    float scaling;
};

__device__ __forceinline__ float ex2_approx(float x) {   // 2^x on the MUFU, flush-to-zero (inputs here are <= 0; -inf -> 0)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16 fp32 -> two 16-byte chunks (8 halfs hi, 8 halfs lo)
__device__ __forceinline__ void pack8_split(const float* p, uint4& hi, uint4& lo) {
    split2(p[0], p[1], hi.x, lo.x);
    split2(p[2], p[3], hi.y, lo.y);
    split2(p[4], p[5], hi.z, lo.z);
    split2(p[6], p[7], hi.w, lo.w);
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(kFaThreads, HD == 64 ? 2 : 1) flash_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_qh, const __grid_constant__ CUtensorMap tm_ql,
                                                                     const __grid_constant__ CUtensorMap tm_kh, const __grid_constant__ CUtensorMap tm_kl,
                                                                     const __grid_constant__ CUtensorMap tm_vh, const __grid_constant__ CUtensorMap tm_vl, FlashTcArgs a) {
    constexpr int BQ = kFaBQ, BKV = kFaBKV, KB = HD / 64;   // KB: 64-half k-blocks per head row (one TMA box each)
    constexpr int kQBytes = BQ * HD * 2, kKBytes = BKV * HD * 2, kVBytes = HD * BKV * 2, kPBytes = BQ * BKV * 2;   // head_dim 64: 16, 8, 8, 16 KB
    constexpr int kQBlk = BQ * 128, kKBlk = BKV * 128;       // bytes of one k-block of the Q / K tile
    extern __shared__ __align__(1024) uint8_t fa_tc_smem_raw[];
    uint8_t* base = fa_tc_smem_raw + ((1024u - (smem_u32(fa_tc_smem_raw) & 1023u)) & 1023u);
    uint8_t* Qh = base;               uint8_t* Ql = Qh + kQBytes;
    uint8_t* Kh = Ql + kQBytes;       uint8_t* Kl = Kh + kKBytes;
    uint8_t* Vh = Kl + kKBytes;       uint8_t* Vl = Vh + kVBytes;
    uint8_t* Ph = Vl + kVBytes;       uint8_t* Pl = Ph + kPBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Pl + kPBytes);
    uint64_t *bar_q = bars, *bar_k = bars + 1, *bar_v = bars + 2, *bar_o = bars + 3, *bar_p = bars + 4, *bar_s = bars + 5;   // bar_s[2]: one per S buffer
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 7);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int head = blockIdx.y, kvh = head / a.groups;
    const int qt0 = blockIdx.x * BQ;
    const int causal_shift = a.Skv - a.Sq;
    int kv_end = a.Skv;
    if (CAUSAL) kv_end = min(a.Skv, qt0 + BQ + causal_shift);
    const int ntiles = (kv_end + BKV - 1) / BKV;

    if (tid == 0) {
        mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1); mbar_init(&bar_s[0], 1); mbar_init(&bar_s[1], 1); mbar_init(bar_o, 1); mbar_init(bar_p, BQ);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {   // TMEM: 2 x 64 columns for the S buffers + 64 for the PV tile (256 allocated: a power of two; two CTAs fill the SM's 512)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_s = *tmem_holder, tmem_o = tmem_s + 128u;

    if (warp == 4) {
        // ======================= control lane: TMA + MMA issue.  S runs ONE TILE AHEAD of the softmax (two S buffers in TMEM) and
        // PV(t - 1) is issued the moment P(t - 1) is complete, so both MMAs execute under the softmax threads' arithmetic.
        if (lane == 0) {
            const int qrow = head * a.Sq + qt0, krow = kvh * a.Skv, vrow = kvh * HD;
            auto load_k = [&](int t) {
                mbar_expect_tx(bar_k, 2 * kKBytes);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    tma_load_2d(Kh + kb * kKBlk, &tm_kh, kb * 64, krow + t * BKV, bar_k);
                    tma_load_2d(Kl + kb * kKBlk, &tm_kl, kb * 64, krow + t * BKV, bar_k);
                }
            };
            mbar_expect_tx(bar_q, 2 * kQBytes);
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                tma_load_2d(Qh + kb * kQBlk, &tm_qh, kb * 64, qrow, bar_q);
                tma_load_2d(Ql + kb * kQBlk, &tm_ql, kb * 64, qrow, bar_q);
            }
            load_k(0);
            mbar_expect_tx(bar_v, 2 * kVBytes);
            tma_load_2d(Vh, &tm_vh, 0, vrow, bar_v);
            tma_load_2d(Vl, &tm_vl, 0, vrow, bar_v);
            const uint32_t idesc = umma_idesc_f16(BQ, 64), idesc_pv = umma_idesc_f16(BQ, HD);
            const uint64_t d_qh = umma_desc_sw128(Qh), d_ql = umma_desc_sw128(Ql), d_kh = umma_desc_sw128(Kh), d_kl = umma_desc_sw128(Kl);
            const uint64_t d_vh = umma_desc_sw128(Vh), d_vl = umma_desc_sw128(Vl), d_ph = umma_desc_sw128(Ph), d_pl = umma_desc_sw128(Pl);
            auto issue_s = [&](int t) {                          // S(t) = Qh Kh^T + Qh Kl^T + Ql Kh^T into S buffer t & 1
                const uint32_t dst = tmem_s + (uint32_t)((t & 1) * 64);
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks) {   // k-block ks / 4 of the tiles, 32 bytes per K = 16 step inside its 128-byte swizzle atom
                    const uint64_t qa = (uint64_t)(((ks >> 2) * kQBlk + (ks & 3) * 32) >> 4), ka = (uint64_t)(((ks >> 2) * kKBlk + (ks & 3) * 32) >> 4);
                    umma_f16(dst, d_qh + qa, d_kh + ka, idesc, ks > 0 ? 1u : 0u);
                    umma_f16(dst, d_qh + qa, d_kl + ka, idesc, 1u);
                    umma_f16(dst, d_ql + qa, d_kh + ka, idesc, 1u);
                }
                umma_commit(&bar_s[t & 1]);
            };
            auto issue_pv = [&](int t) {                         // PV(t) = Ph Vh + Ph Vl + Pl Vh   (B = V^T tile, K-major over the kv tokens)
                mbar_wait(bar_p, (uint32_t)(t & 1));             // P(t) is in shared memory; S buffer (t + 1) & 1 ... and PV(t - 1) have been read
                mbar_wait(bar_v, (uint32_t)(t & 1));
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < BKV / 16; ++ks) {
                    const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                    umma_f16(tmem_o, d_ph + adv, d_vh + adv, idesc_pv, ks > 0 ? 1u : 0u);
                    umma_f16(tmem_o, d_ph + adv, d_vl + adv, idesc_pv, 1u);
                    umma_f16(tmem_o, d_pl + adv, d_vh + adv, idesc_pv, 1u);
                }
                umma_commit(bar_o);
            };
            mbar_wait(bar_q, 0);
            mbar_wait(bar_k, 0);
            tc_fence_after();
            issue_s(0);
            for (int t = 0; t < ntiles; ++t) {
                mbar_wait(&bar_s[t & 1], (uint32_t)((t >> 1) & 1));   // S(t) done: the K tile is free, fetch K(t + 1)
                if (t + 1 < ntiles) load_k(t + 1);
                if (t >= 1) issue_pv(t - 1);                     // runs under the softmax of tile t
                if (t + 1 < ntiles) {                            // S(t + 1) into the other buffer (its last reader, softmax(t - 1), has arrived at bar_p(t - 1))
                    mbar_wait(bar_k, (uint32_t)((t + 1) & 1));
                    tc_fence_after();
                    issue_s(t + 1);
                }
                if (t >= 1) {
                    mbar_wait(bar_o, (uint32_t)((t - 1) & 1));    // PV(t - 1) done: the V tile is free, fetch V(t)
                    mbar_expect_tx(bar_v, 2 * kVBytes);
                    tma_load_2d(Vh, &tm_vh, t * BKV, vrow, bar_v);
                    tma_load_2d(Vl, &tm_vl, t * BKV, vrow, bar_v);
                }
            }
            issue_pv(ntiles - 1);
        }
        __syncwarp();
    } else {
        // ======================= softmax threads: row = tid (TMEM lane tid)
        const int row = tid, qi = qt0 + row;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        float o[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) o[c] = 0.f;
        float m = -INFINITY, l = 0.f;                            // m in the log2 domain: exp(x) = 2^(x * log2 e)
        const float scale_log2 = a.scaling * 1.4426950408889634f;
        uint8_t* prow_h = Ph + (size_t)row * 128;
        uint8_t* prow_l = Pl + (size_t)row * 128;
        for (int t = 0; t < ntiles; ++t) {
            const int kt0 = t * BKV;
            mbar_wait(&bar_s[t & 1], (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            float s[BKV];
            {
                const uint32_t src = tmem_s + (uint32_t)((t & 1) * 64) + lane_base;
                float v0[32], v1[32];
                tmem_ld32(src, v0);
                tmem_ld32(src + 32u, v1);
#pragma unroll
                for (int c = 0; c < 32; ++c) { s[c] = v0[c]; s[32 + c] = v1[c]; }
            }
            // The softmax arithmetic is what bounds this kernel (the tensor pipe idles under it), so it is kept minimal: masking only on
            // the tiles that can need it (the last one / the causal diagonal), the max on raw scores (scale > 0 commutes with max), scale
            // and max subtraction in ONE fma feeding ex2.approx, and the rescale of O skipped while no row of the warp moved its max.
            const bool mask_tile = (kt0 + BKV > a.Skv) || (CAUSAL && kt0 + BKV - 1 > qt0 + (int)(warp * 32) + causal_shift);
            float rmax = -INFINITY;
            if (mask_tile) {
#pragma unroll
                for (int c = 0; c < BKV; ++c) {
                    const int kj = kt0 + c;
                    if (kj >= a.Skv || (CAUSAL && kj > qi + causal_shift)) s[c] = -INFINITY;
                    rmax = fmaxf(rmax, s[c]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < BKV; ++c) rmax = fmaxf(rmax, s[c]);
            }
            const float mn = fmaxf(m, rmax * scale_log2);
            const float mu = (mn == -INFINITY) ? 0.f : mn;
            const float alpha = ex2_approx(m - mu);
            float rs = 0.f;
#pragma unroll
            for (int c = 0; c < BKV; ++c) { s[c] = ex2_approx(fmaf(s[c], scale_log2, -mu)); rs += s[c]; }
            l = l * alpha + rs;
            m = mn;
            if (t >= 1) {   // PV(t - 1) ran under the arithmetic above; it also has to be done before P(t) may overwrite P(t - 1)
                mbar_wait(bar_o, (uint32_t)((t - 1) & 1));
                tc_fence_after();
                const bool rescale = __any_sync(0xffffffffu, alpha != 1.0f);
#pragma unroll
                for (int c0 = 0; c0 < HD; c0 += 32) {
                    float v0[32];
                    tmem_ld32(tmem_o + lane_base + (uint32_t)c0, v0);
                    if (rescale) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[c0 + c] = (o[c0 + c] + v0[c]) * alpha;
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[c0 + c] += v0[c];
                    }
                }
            }
            // P row -> shared memory, K-major SWIZZLE_128B: 16-byte chunk j of row r sits at chunk (j ^ (r & 7))
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 hi, lo;
                pack8_split(s + 8 * j, hi, lo);
                const int sw = (j ^ (row & 7)) << 4;
                *reinterpret_cast<uint4*>(prow_h + sw) = hi;
                *reinterpret_cast<uint4*>(prow_l + sw) = lo;
            }
            fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core's async-proxy reads
            tc_fence_before();
            mbar_arrive(bar_p);
        }
        {
            mbar_wait(bar_o, (uint32_t)((ntiles - 1) & 1));
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < HD; c0 += 32) {
                float v0[32];
                tmem_ld32(tmem_o + lane_base + (uint32_t)c0, v0);
#pragma unroll
                for (int c = 0; c < 32; ++c) o[c0 + c] += v0[c];
            }
            tc_fence_before();
        }
        if (qi < a.Sq) {
            const float inv = 1.0f / l;
            const size_t off = (size_t)(a.q0 + qi) * a.o_tok_stride + (size_t)head * a.o_head_stride;
            if (a.out_hi) {
#pragma unroll
                for (int c = 0; c < HD; c += 8) {
                    float r[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) r[e] = o[c + e] * inv;
                    uint4 hi, lo;
                    pack8_split(r, hi, lo);
                    *reinterpret_cast<uint4*>(a.out_hi + off + c) = hi;
                    *reinterpret_cast<uint4*>(a.out_lo + off + c) = lo;
                }
            } else {
                float* dst = a.out + off;
#pragma unroll
                for (int c = 0; c < HD; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_s), "n"(256) : "memory");
    }
}

template <int HD>
inline size_t flash_tc_smem_bytes() { return (size_t)(2 * kFaBQ * HD * 2 + 2 * kFaBKV * HD * 2 + 2 * HD * kFaBKV * 2 + 2 * kFaBQ * kFaBKV * 2) + 8 * sizeof(uint64_t) + 1024; }
inline int flash_tc_skv_pad(int Skv) { return ceil_div(Skv, 64) * 64; }
// workspace halfs: q hi|lo + k hi|lo ([head][token][HD]) + v^T hi|lo ([head][HD][skv_pad])
template <int HD>
inline size_t flash_tc_ws_halfs(const FlashArgs& a, int nheads) {
    const int nkv = nheads / a.groups;
    return 2 * ((size_t)nheads * a.Sq + (size_t)nkv * a.Skv + (size_t)nkv * flash_tc_skv_pad(a.Skv)) * HD;
}
inline void flash_attn_tc_init() {
    AHA_CUDA_CHECK(cudaFuncSetAttribute(flash_attn_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flash_tc_smem_bytes<64>()));
    AHA_CUDA_CHECK(cudaFuncSetAttribute(flash_attn_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flash_tc_smem_bytes<64>()));
    AHA_CUDA_CHECK(cudaFuncSetAttribute(flash_attn_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flash_tc_smem_bytes<128>()));
    AHA_CUDA_CHECK(cudaFuncSetAttribute(flash_attn_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flash_tc_smem_bytes<128>()));
}
// head_dim 64 (ViT, audio encoder: 2 CTAs per SM) and 128 (LLM prefill, causal, GQA, paged KV: 1 CTA per SM).  `ws`: at least flash_tc_ws_halfs<HD>() halfs.
template <int HD>
inline void flash_attn_tc(cudaStream_t st, const FlashArgs& a, int nheads, bool causal, __half* ws) {
    if (a.Sq == 0) return;
    const int nkv = nheads / a.groups, pad = flash_tc_skv_pad(a.Skv);
    SplitTcArgs sp;
    sp.f = a; sp.nheads = nheads; sp.nkv = nkv; sp.skv_pad = pad;
    const size_t nq = (size_t)nheads * a.Sq * HD, nk = (size_t)nkv * a.Skv * HD, nv = (size_t)nkv * HD * pad;
    sp.q_hi = ws; sp.q_lo = ws + nq; sp.k_hi = ws + 2 * nq; sp.k_lo = sp.k_hi + nk; sp.vt_hi = sp.k_lo + nk; sp.vt_lo = sp.vt_hi + nv;
    const unsigned blocks = (unsigned)std::min<size_t>(std::max<size_t>(ceil_div((int)std::max((size_t)nheads * a.Sq, (size_t)nkv * a.Skv), 512 / HD), (size_t)nkv * (pad / 64)), 148 * 32);
    split_qkv_tc_kernel<HD><<<dim3(blocks, 3), 128, 0, st>>>(sp);
    const CUtensorMap tqh = make_tmap_f16_box(sp.q_hi, (uint64_t)nheads * a.Sq, HD, kFaBQ), tql = make_tmap_f16_box(sp.q_lo, (uint64_t)nheads * a.Sq, HD, kFaBQ);
    const CUtensorMap tkh = make_tmap_f16_box(sp.k_hi, (uint64_t)nkv * a.Skv, HD, kFaBKV), tkl = make_tmap_f16_box(sp.k_lo, (uint64_t)nkv * a.Skv, HD, kFaBKV);
    const CUtensorMap tvh = make_tmap_f16_box(sp.vt_hi, (uint64_t)nkv * HD, (uint64_t)pad, HD), tvl = make_tmap_f16_box(sp.vt_lo, (uint64_t)nkv * HD, (uint64_t)pad, HD);
    FlashTcArgs m;
    m.out = a.out; m.o_tok_stride = a.o_tok_stride; m.o_head_stride = a.o_head_stride; m.out_hi = a.out_hi; m.out_lo = a.out_lo;
    m.Sq = a.Sq; m.Skv = a.Skv; m.q0 = a.q0; m.groups = a.groups; m.scaling = a.scaling;
    dim3 grid(ceil_div(a.Sq, kFaBQ), nheads);
    const size_t smem = flash_tc_smem_bytes<HD>();
    if (causal) flash_attn_tc_kernel<HD, true><<<grid, kFaThreads, smem, st>>>(tqh, tql, tkh, tkl, tvh, tvl, m);
    else flash_attn_tc_kernel<HD, false><<<grid, kFaThreads, smem, st>>>(tqh, tql, tkh, tkl, tvh, tvl, m);
    AHA_CUDA_CHECK(cudaGetLastError());
}

}  // namespace aha
