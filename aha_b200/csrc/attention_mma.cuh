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

// ---- pre-pass: split fp32 q / k / v rows into packed fp16 hi | lo buffers, layout [head][token][HD] ------------------
// (one pass per layer instead of one conversion per CTA per KV tile)
struct SplitQKVArgs {
    FlashArgs f;
    __half *q_hi, *q_lo, *k_hi, *k_lo, *v_hi, *v_lo;
    int nheads, nkv;
};
template <int HD>
__global__ void split_qkv_kernel(SplitQKVArgs s) {
    // blockIdx.y: 0 = q, 1 = k, 2 = v; blockIdx.x: row index (head * S + token)
    const int which = blockIdx.y;
    const int S = which == 0 ? s.f.Sq : s.f.Skv, H = which == 0 ? s.nheads : s.nkv;
    const size_t row = blockIdx.x;
    if (row >= (size_t)S * H) return;
    const int head = (int)(row / S), tok = (int)(row % S);
    const float* src = which == 0 ? s.f.q + (size_t)(s.f.q0 + tok) * s.f.q_tok_stride + (size_t)head * s.f.q_head_stride
                                  : (which == 1 ? s.f.kv.k : s.f.kv.v) + s.f.kv.off(s.f.kv0 + tok, head);
    __half* hi = (which == 0 ? s.q_hi : which == 1 ? s.k_hi : s.v_hi) + row * HD;
    __half* lo = (which == 0 ? s.q_lo : which == 1 ? s.k_lo : s.v_lo) + row * HD;
    for (int c4 = threadIdx.x * 4; c4 < HD; c4 += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + c4);
        uint32_t h0, l0, h1, l1;
        split2(v.x, v.y, h0, l0);
        split2(v.z, v.w, h1, l1);
        *reinterpret_cast<uint2*>(hi + c4) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(lo + c4) = make_uint2(l0, l1);
    }
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(fa_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct FlashMmaArgs {
    const __half *q_hi, *q_lo, *k_hi, *k_lo, *v_hi, *v_lo;   // packed [head][token][HD]
    float* out; size_t o_tok_stride, o_head_stride;
    int Sq, Skv, q0, groups;
    float scaling;
};

// One CTA = 128 queries of one head (8 warps x 16 rows); K/V tiles of 64 tokens, cp.async double buffered.
template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(256) flash_attn_mma_kernel(FlashMmaArgs a) {
    constexpr int BQ = 128, BKV = 64, LD = HD + 8;   // row stride in halfs: +16 bytes keeps ldmatrix conflict-free
    constexpr int CH = HD / 8;                       // 16-byte chunks per row
    extern __shared__ __align__(16) __half fa_smem[];
    __half* Qh = fa_smem;            __half* Ql = Qh + BQ * LD;
    __half* KV = Ql + BQ * LD;       // [2 buffers][Kh | Kl | Vh | Vl][BKV * LD]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int head = blockIdx.y, kvh = head / a.groups;
    const int qt0 = blockIdx.x * BQ;
    const size_t qbase = (size_t)head * a.Sq, kbase = (size_t)kvh * a.Skv;

    auto load_kv = [&](int tile, int buf) {
        __half* dst = KV + (size_t)buf * 4 * BKV * LD;
        const int kt0 = tile * BKV;
        for (int idx = tid; idx < 4 * BKV * CH; idx += 256) {
            const int arr = idx / (BKV * CH), rem = idx % (BKV * CH), r = rem / CH, c = rem % CH;
            const int tok = min(kt0 + r, a.Skv - 1);   // rows past Skv are masked; clamp keeps the data finite
            const __half* src = (arr == 0 ? a.k_hi : arr == 1 ? a.k_lo : arr == 2 ? a.v_hi : a.v_lo) + (kbase + tok) * HD + c * 8;
            cp_async16(dst + (size_t)arr * BKV * LD + r * LD + c * 8, src);
        }
    };
    // Q tile (once) + first K/V tile
    for (int idx = tid; idx < 2 * BQ * CH; idx += 256) {
        const int arr = idx / (BQ * CH), rem = idx % (BQ * CH), r = rem / CH, c = rem % CH;
        const int tok = min(qt0 + r, a.Sq - 1);
        cp_async16((arr == 0 ? Qh : Ql) + r * LD + c * 8, (arr == 0 ? a.q_hi : a.q_lo) + (qbase + tok) * HD + c * 8);
    }
    const int causal_shift = a.Skv - a.Sq;
    int kv_end = a.Skv;
    if (CAUSAL) kv_end = min(a.Skv, qt0 + BQ + causal_shift);
    const int ntiles = (kv_end + BKV - 1) / BKV;
    load_kv(0, 0);
    cp_async_commit();

    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows g and g+8 of this warp's 16
    const int qrow0 = qt0 + warp * 16 + g, qrow1 = qrow0 + 8;

    for (int tile = 0; tile < ntiles; ++tile) {
        const int kt0 = tile * BKV, buf = tile & 1;
        cp_async_wait<0>();
        __syncthreads();                       // tile `tile` landed for everyone; buffer buf^1 is free again
        if (tile + 1 < ntiles) load_kv(tile + 1, buf ^ 1);
        cp_async_commit();
        const __half* Kh = KV + (size_t)buf * 4 * BKV * LD;
        const __half* Kl = Kh + BKV * LD;
        const __half* Vh = Kl + BKV * LD;
        const __half* Vl = Vh + BKV * LD;
        // a warp whose 16 rows are entirely above the causal frontier of this tile has nothing to add
        const bool warp_active = !CAUSAL || (kt0 <= qt0 + warp * 16 + 15 + causal_shift);
        if (warp_active) {
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
        }   // warp_active
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

// `ws`: workspace of at least flash_mma_ws_halfs() halfs (hi/lo copies of q, k, v).
template <int HD>
inline size_t flash_mma_ws_halfs(const FlashArgs& a, int nheads) {
    const int nkv = nheads / a.groups;
    return 2 * ((size_t)nheads * a.Sq + 2 * (size_t)nkv * a.Skv) * HD;
}
template <int HD>
inline void flash_attn_mma(cudaStream_t st, const FlashArgs& a, int nheads, bool causal, __half* ws) {
    if (a.Sq == 0) return;
    const int nkv = nheads / a.groups;
    SplitQKVArgs sp;
    sp.f = a; sp.nheads = nheads; sp.nkv = nkv;
    const size_t nq = (size_t)nheads * a.Sq * HD, nk = (size_t)nkv * a.Skv * HD;
    sp.q_hi = ws; sp.q_lo = ws + nq; sp.k_hi = ws + 2 * nq; sp.k_lo = sp.k_hi + nk; sp.v_hi = sp.k_lo + nk; sp.v_lo = sp.v_hi + nk;
    const unsigned rows = (unsigned)std::max((size_t)nheads * a.Sq, (size_t)nkv * a.Skv);
    split_qkv_kernel<HD><<<dim3(rows, 3), HD / 4, 0, st>>>(sp);
    FlashMmaArgs m;
    m.q_hi = sp.q_hi; m.q_lo = sp.q_lo; m.k_hi = sp.k_hi; m.k_lo = sp.k_lo; m.v_hi = sp.v_hi; m.v_lo = sp.v_lo;
    m.out = a.out; m.o_tok_stride = a.o_tok_stride; m.o_head_stride = a.o_head_stride;
    m.Sq = a.Sq; m.Skv = a.Skv; m.q0 = a.q0; m.groups = a.groups; m.scaling = a.scaling;
    const size_t smem = (size_t)(2 * 128 + 8 * 64) * (HD + 8) * sizeof(__half);
    dim3 grid(ceil_div(a.Sq, 128), nheads);
    if (causal) {
        AHA_CUDA_CHECK(cudaFuncSetAttribute(flash_attn_mma_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        flash_attn_mma_kernel<HD, true><<<grid, 256, smem, st>>>(m);
    } else {
        AHA_CUDA_CHECK(cudaFuncSetAttribute(flash_attn_mma_kernel<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        flash_attn_mma_kernel<HD, false><<<grid, 256, smem, st>>>(m);
    }
    AHA_CUDA_CHECK(cudaGetLastError());
}

}  // namespace aha
