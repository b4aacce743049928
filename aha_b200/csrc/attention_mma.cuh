// attention_mma.cuh -- flash attention on the tensor cores for prefill shapes (ViT segments, audio encoder, causal LLM
// prefill over the paged KV cache) with fp32-grade accuracy.
//
// Same contract as flash_attn_kernel (attention.cuh): softmax(Q K^T * scale [+ causal mask]) V, GQA head mapping,
// online softmax, never materialising S x S (reference: eager_attention_forward,
// /root/reference/src/models/common/modules.rs:757-813; ViT per-segment attention qwen3vl/model.rs:258-277).
// fp16 tensor-core operands alone would break the 1e-3 logit budget, so every operand is split
// x = hi + lo (fp16 each) when its tile is staged in shared memory and each product is formed as
// hi*hi + hi*lo + lo*hi with fp32 accumulation (the dropped lo*lo term is ~2^-22 relative):
//   S = Q K^T : mma.sync m16n8k16, A = Q fragments (ldmatrix), B = K rows (ldmatrix, K is [kv][d] = col-major B);
//   P = softmax tile kept in registers in the accumulator layout, re-packed as A fragments (hi, lo);
//   O += P V  : B = V^T fragments (ldmatrix.trans on V [kv][d]).
// One CTA = 64 queries of one head (4 warps x 16 rows), KV tiles of 64.
#pragma once
#include "attention.cuh"

namespace aha {

__device__ __forceinline__ uint32_t fa_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(fa_smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(fa_smem_u32(p)));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(128) flash_attn_mma_kernel(FlashArgs a) {
    constexpr int BQ = 64, BKV = 64, LD = HD + 8;   // row stride in halfs: +16 bytes keeps ldmatrix conflict-free
    extern __shared__ __align__(16) __half fa_smem[];
    __half* Qh = fa_smem;            __half* Ql = Qh + BQ * LD;
    __half* Kh = Ql + BQ * LD;       __half* Kl = Kh + BKV * LD;
    __half* Vh = Kl + BKV * LD;      __half* Vl = Vh + BKV * LD;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int head = blockIdx.y, kvh = head / a.groups;
    const int qt0 = blockIdx.x * BQ;

    // stage one [64 x HD] fp32 tile as hi/lo fp16 (rows past `valid` are zero)
    auto stage = [&](const float* base, size_t row_off_fn_dummy, __half* hi, __half* lo, int valid, auto row_ptr) {
        (void)base; (void)row_off_fn_dummy;
        for (int idx = tid; idx < 64 * (HD / 4); idx += 128) {
            const int r = idx / (HD / 4), c4 = (idx % (HD / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < valid) v = *reinterpret_cast<const float4*>(row_ptr(r) + c4);
            uint32_t h0, l0, h1, l1;
            split2(v.x, v.y, h0, l0);
            split2(v.z, v.w, h1, l1);
            *reinterpret_cast<uint2*>(hi + r * LD + c4) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(lo + r * LD + c4) = make_uint2(l0, l1);
        }
    };
    stage(nullptr, 0, Qh, Ql, min(BQ, a.Sq - qt0),
          [&](int r) { return a.q + (size_t)(a.q0 + qt0 + r) * a.q_tok_stride + (size_t)head * a.q_head_stride; });

    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows g and g+8 of this warp's 16
    const int causal_shift = a.Skv - a.Sq;
    int kv_end = a.Skv;
    if (CAUSAL) kv_end = min(a.Skv, qt0 + BQ + causal_shift);
    const int ntiles = (kv_end + BKV - 1) / BKV;
    const int qrow0 = qt0 + warp * 16 + g, qrow1 = qrow0 + 8;

    for (int tile = 0; tile < ntiles; ++tile) {
        const int kt0 = tile * BKV;
        __syncthreads();   // previous tile consumed (and the Q tile staged, first iteration)
        const int valid = min(BKV, a.Skv - kt0);
        stage(nullptr, 0, Kh, Kl, valid, [&](int r) { return a.kv.k + a.kv.off(a.kv0 + kt0 + r, kvh); });
        stage(nullptr, 0, Vh, Vl, valid, [&](int r) { return a.kv.v + a.kv.off(a.kv0 + kt0 + r, kvh); });
        __syncthreads();

        // ---- S = Q K^T (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
            uint32_t qh[4], ql[4];
            const int qr = warp * 16 + (lane & 15), qc = ks * 16 + (lane >> 4) * 8;
            ldsm_x4(qh, Qh + qr * LD + qc);
            ldsm_x4(ql, Ql + qr * LD + qc);
#pragma unroll
            for (int np = 0; np < 4; ++np) {   // two 8-wide kv tiles per ldmatrix.x4
                uint32_t kh[4], kl[4];
                const int kr = np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, kc = ks * 16 + ((lane >> 3) & 1) * 8;
                ldsm_x4(kh, Kh + kr * LD + kc);
                ldsm_x4(kl, Kl + kr * LD + kc);
                mma_f16(s[2 * np], qh, kh[0], kh[1]);
                mma_f16(s[2 * np], qh, kl[0], kl[1]);
                mma_f16(s[2 * np], ql, kh[0], kh[1]);
                mma_f16(s[2 * np + 1], qh, kh[2], kh[3]);
                mma_f16(s[2 * np + 1], qh, kl[2], kl[3]);
                mma_f16(s[2 * np + 1], ql, kh[2], kh[3]);
            }
        }
        // ---- scale, mask, online softmax (rows g / g+8; this lane holds cols nt*8 + 2t, +1)
        float rmax0 = -INFINITY, rmax1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kj = kt0 + nt * 8 + 2 * t + (e & 1);
                const int qi = (e < 2) ? qrow0 : qrow1;
                float v = s[nt][e] * a.scaling;
                if (kj >= a.Skv || (CAUSAL && kj > qi + causal_shift)) v = -INFINITY;
                s[nt][e] = v;
                if (e < 2) rmax0 = fmaxf(rmax0, v); else rmax1 = fmaxf(rmax1, v);
            }
        }
        rmax0 = fmaxf(rmax0, __shfl_xor_sync(0xffffffffu, rmax0, 1)); rmax0 = fmaxf(rmax0, __shfl_xor_sync(0xffffffffu, rmax0, 2));
        rmax1 = fmaxf(rmax1, __shfl_xor_sync(0xffffffffu, rmax1, 1)); rmax1 = fmaxf(rmax1, __shfl_xor_sync(0xffffffffu, rmax1, 2));
        const float mn0 = fmaxf(m0, rmax0), mn1 = fmaxf(m1, rmax1);
        const float mu0 = (mn0 == -INFINITY) ? 0.f : mn0, mu1 = (mn1 == -INFINITY) ? 0.f : mn1;
        const float al0 = expf(m0 - mu0), al1 = expf(m1 - mu1);
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = expf(s[nt][0] - mu0); s[nt][1] = expf(s[nt][1] - mu0);
            s[nt][2] = expf(s[nt][2] - mu1); s[nt][3] = expf(s[nt][3] - mu1);
            rs0 += s[nt][0] + s[nt][1];
            rs1 += s[nt][2] + s[nt][3];
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        l0 = l0 * al0 + rs0; l1 = l1 * al1 + rs1;
        m0 = mn0; m1 = mn1;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) { o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1; }
        // ---- O += P V   (P re-packed from the accumulator layout into A fragments, hi and lo)
#pragma unroll
        for (int j = 0; j < BKV / 16; ++j) {
            uint32_t ph[4], pl[4];
            split2(s[2 * j][0], s[2 * j][1], ph[0], pl[0]);
            split2(s[2 * j][2], s[2 * j][3], ph[1], pl[1]);
            split2(s[2 * j + 1][0], s[2 * j + 1][1], ph[2], pl[2]);
            split2(s[2 * j + 1][2], s[2 * j + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int dp = 0; dp < HD / 16; ++dp) {   // two 8-wide d tiles per ldmatrix.x4.trans
                uint32_t vh[4], vl[4];
                const int vr = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, vc = dp * 16 + ((lane >> 4) & 1) * 8;
                ldsm_x4_t(vh, Vh + vr * LD + vc);
                ldsm_x4_t(vl, Vl + vr * LD + vc);
                mma_f16(o[2 * dp], ph, vh[0], vh[1]);
                mma_f16(o[2 * dp], ph, vl[0], vl[1]);
                mma_f16(o[2 * dp], pl, vh[0], vh[1]);
                mma_f16(o[2 * dp + 1], ph, vh[2], vh[3]);
                mma_f16(o[2 * dp + 1], ph, vl[2], vl[3]);
                mma_f16(o[2 * dp + 1], pl, vh[2], vh[3]);
            }
        }
    }
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        const int d = i * 8 + 2 * t;
        if (qrow0 < a.Sq)
            *reinterpret_cast<float2*>(a.out + (size_t)(a.q0 + qrow0) * a.o_tok_stride + (size_t)head * a.o_head_stride + d) = make_float2(o[i][0] * inv0, o[i][1] * inv0);
        if (qrow1 < a.Sq)
            *reinterpret_cast<float2*>(a.out + (size_t)(a.q0 + qrow1) * a.o_tok_stride + (size_t)head * a.o_head_stride + d) = make_float2(o[i][2] * inv1, o[i][3] * inv1);
    }
}

template <int HD>
inline void flash_attn_mma(cudaStream_t st, const FlashArgs& a, int nheads, bool causal) {
    if (a.Sq == 0) return;
    const size_t smem = (size_t)6 * 64 * (HD + 8) * sizeof(__half);
    dim3 grid(ceil_div(a.Sq, 64), nheads);
    if (causal) {
        AHA_CUDA_CHECK(cudaFuncSetAttribute(flash_attn_mma_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        flash_attn_mma_kernel<HD, true><<<grid, 128, smem, st>>>(a);
    } else {
        AHA_CUDA_CHECK(cudaFuncSetAttribute(flash_attn_mma_kernel<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        flash_attn_mma_kernel<HD, false><<<grid, 128, smem, st>>>(a);
    }
    AHA_CUDA_CHECK(cudaGetLastError());
}

}  // namespace aha
