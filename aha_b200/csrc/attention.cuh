// attention.cuh -- q/k RMSNorm + RoPE + paged-KV write, flash-style prefill attention (fp32 SIMT, exact),
// and split-KV decode attention with the per-head norm/RoPE/KV-append fused into its prologue.
//
// Reference semantics: QKNormAttention::forward (/root/reference/src/models/common/modules.rs:530-579):
//   q,k per-head RMSNorm over head_dim BEFORE RoPE; apply_rotary_pos_emb = x*cos + rotate_half(x)*sin
//   (/root/reference/src/position_embed/rope.rs:15-22,96-132); K (rotated) and V appended to the cache;
//   eager_attention_forward (modules.rs:757-813): softmax(QK^T * 1/sqrt(hd) + causal mask) V with GQA
//   head h <-> kv head h / n_rep (/root/reference/src/utils/tensor_utils.rs:108-124).
// The reference re-concatenates the whole cache every step and materialises S x S scores; here the cache
// is paged and written in place, and scores never leave the SM.
#pragma once
#include "common.cuh"
#include "kernels_common.cuh"

namespace aha {

constexpr int kPageShift = 5;            // 32-token KV pages
constexpr int kPage = 1 << kPageShift;

// ---------------------------------------------------------------------------------------------------
// KV addressing.  Paged pool layout (fp32): [layer][page][K|V][kv_head][PAGE tokens][hd].
// A contiguous source (ViT / audio encoder, no cache) sets page_table = nullptr.
struct KVSrc {
    const float* k;          // base of K (layer offset applied)
    const float* v;          // base of V
    const int* page_table;   // token>>page_shift -> physical page, or nullptr
    int page_shift;          // log2(PAGE)
    size_t page_stride;      // floats between consecutive physical pages
    size_t tok_stride;       // floats between consecutive tokens (inside a page, or globally if contiguous)
    size_t head_stride;      // floats between kv heads
    __device__ __forceinline__ size_t off(int tok, int head) const {
        size_t o = (size_t)head * head_stride;
        if (page_table) {
            o += (size_t)page_table[tok >> page_shift] * page_stride + (size_t)(tok & ((1 << page_shift) - 1)) * tok_stride;
        } else {
            o += (size_t)tok * tok_stride;
        }
        return o;
    }
};

// ---------------------------------------------------------------------------------------------------
// RoPE parameters.  inv_freq[j] = 1 / theta^(2j/hd) (f32 powf on the host, rope.rs:7-13).
// sel[j] in {0,1,2}: which of the 3 M-RoPE position rows drives frequency j (all 0 for 1-D RoPE;
// interleaved rule of rope.rs:454-476 for Qwen3-VL).  pos3: [3][S] int32.
struct RopeArgs {
    const float* inv_freq;
    const uint8_t* sel;
    const int* pos3;
    int S;
};

// Prefill: one block per (token, head slot).  Slots [0,nh) are q heads, [nh, nh+nkv) k heads, then v heads.
// In: qkv [S, (nh+2nkv)*hd] raw projections.  Out: q normalised+rotated in place; K,V written to the cache at
// token index pos0 + s.
template <int HD>
__global__ void __launch_bounds__(HD) qk_norm_rope_kv_kernel(float* __restrict__ qkv, const float* __restrict__ qw,
                                                            const float* __restrict__ kw, float eps, RopeArgs rp,
                                                            float* __restrict__ kdst, float* __restrict__ vdst,
                                                            KVSrc kv, int nh, int nkv, int pos0) {
    __shared__ float red[32];
    __shared__ float xs[HD];
    const int s = blockIdx.x, slot = blockIdx.y, d = threadIdx.x;
    const int row = (nh + 2 * nkv) * HD;
    float* src = qkv + (size_t)s * row + (size_t)slot * HD;
    const float x = src[d];
    if (slot >= nh + nkv) {  // V: plain copy into the cache
        vdst[kv.off(pos0 + s, slot - nh - nkv) + d] = kv_store_round(x);
        return;
    }
    const bool is_q = slot < nh;
    float ss = block_sum(x * x, red);
    const float inv = 1.0f / sqrtf(ss / (float)HD + eps);
    const float n = x * inv * (is_q ? qw[d] : kw[d]);
    xs[d] = n;
    __syncthreads();
    const int j = d % (HD / 2);
    const float p = (float)rp.pos3[(int)rp.sel[j] * rp.S + s];
    const float ang = p * rp.inv_freq[j];
    const float c = cosf(ang), sn = sinf(ang);
    const float rot = (d < HD / 2) ? -xs[d + HD / 2] : xs[d - HD / 2];
    const float o = n * c + rot * sn;
    if (is_q) src[d] = o;
    else kdst[kv.off(pos0 + s, slot - nh) + d] = kv_store_round(o);
}

// ---------------------------------------------------------------------------------------------------
// Flash-style attention, fp32 on the CUDA cores (exact path / tcgen05 validation baseline).
struct FlashArgs {
    const float* q; size_t q_tok_stride, q_head_stride;   // q[(tok)*q_tok_stride + head*q_head_stride + d]
    KVSrc kv;
    float* out; size_t o_tok_stride, o_head_stride;
    int Sq, Skv;       // queries and keys in this segment
    int q0, kv0;       // first query / kv token index of the segment (varlen segments, ViT cu_seqlens)
    int groups;        // q heads per kv head
    float scaling;
    __half* out_hi = nullptr;   // tcgen05 kernel only: write the output as fp16 hi + lo halves (same strides as `out`) for the GEMM
    __half* out_lo = nullptr;   // that follows, instead of fp32 `out`
};

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(256) flash_attn_kernel(FlashArgs a) {
    constexpr int BQ = 64, BKV = 64, LD = HD + 4, LDP = BKV + 4, NH = HD / 64;
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;
    float* Ks = Qs + BQ * LD;
    float* Vs = Ks + BKV * LD;
    float* Ps = Vs + BKV * LD;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int head = blockIdx.y, kvh = head / a.groups;
    const int qt0 = blockIdx.x * BQ;

    // Q tile
    for (int idx = tid; idx < BQ * (HD / 4); idx += 256) {
        const int r = idx / (HD / 4), c4 = idx % (HD / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qt0 + r < a.Sq)
            v = *reinterpret_cast<const float4*>(a.q + (size_t)(a.q0 + qt0 + r) * a.q_tok_stride + (size_t)head * a.q_head_stride + c4 * 4);
        *reinterpret_cast<float4*>(Qs + r * LD + c4 * 4) = v;
    }

    float m[4], l[4], o[4][NH][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
            for (int e = 0; e < 4; ++e) o[i][h][e] = 0.f;
    }

    const int causal_shift = a.Skv - a.Sq;  // query i sees keys j <= i + shift
    int kv_end = a.Skv;
    if (CAUSAL) kv_end = min(a.Skv, qt0 + BQ + causal_shift);
    const int ntiles = (kv_end + BKV - 1) / BKV;

    for (int t = 0; t < ntiles; ++t) {
        const int kt0 = t * BKV;
        __syncthreads();  // previous tile fully consumed (also orders the Q tile stores before first use)
        for (int idx = tid; idx < BKV * (HD / 4); idx += 256) {
            const int r = idx / (HD / 4), c4 = idx % (HD / 4);
            float4 kx = make_float4(0.f, 0.f, 0.f, 0.f), vx = kx;
            if (kt0 + r < a.Skv) {
                const size_t off = a.kv.off(a.kv0 + kt0 + r, kvh) + c4 * 4;
                kx = *reinterpret_cast<const float4*>(a.kv.k + off);
                vx = *reinterpret_cast<const float4*>(a.kv.v + off);
            }
            *reinterpret_cast<float4*>(Ks + r * LD + c4 * 4) = kx;
            *reinterpret_cast<float4*>(Vs + r * LD + c4 * 4) = vx;
        }
        __syncthreads();

        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
        for (int d4 = 0; d4 < HD / 4; ++d4) {
            float4 qv[4], kvv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * LD + d4 * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) kvv[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * LD + d4 * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[i][j] = fmaf(qv[i].x, kvv[j].x, s[i][j]);
                    s[i][j] = fmaf(qv[i].y, kvv[j].y, s[i][j]);
                    s[i][j] = fmaf(qv[i].z, kvv[j].z, s[i][j]);
                    s[i][j] = fmaf(qv[i].w, kvv[j].w, s[i][j]);
                }
        }
        // scale, mask, online softmax
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int qi = qt0 + ty + 16 * i;
            float rmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kj = kt0 + tx + 16 * j;
                float v = s[i][j] * a.scaling;
                if (kj >= a.Skv || (CAUSAL && kj > qi + causal_shift)) v = -INFINITY;
                s[i][j] = v;
                rmax = fmaxf(rmax, v);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, off));
            const float mnew = fmaxf(m[i], rmax);
            const float muse = (mnew == -INFINITY) ? 0.f : mnew;
            const float alpha = expf(m[i] - muse);  // m = -inf -> 0
            float rsum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = expf(s[i][j] - muse);
                rsum += p;
                Ps[(ty + 16 * i) * LDP + tx + 16 * j] = p;
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, off);
            l[i] = l[i] * alpha + rsum;
            m[i] = mnew;
#pragma unroll
            for (int h = 0; h < NH; ++h)
#pragma unroll
                for (int e = 0; e < 4; ++e) o[i][h][e] *= alpha;
        }
        __syncthreads();
        // O += P V
#pragma unroll 2
        for (int j4 = 0; j4 < BKV / 4; ++j4) {
            float4 pv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * LDP + j4 * 4);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float4 vv[NH];
#pragma unroll
                for (int h = 0; h < NH; ++h) vv[h] = *reinterpret_cast<const float4*>(Vs + (j4 * 4 + jj) * LD + tx * 4 + 64 * h);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float p = jj == 0 ? pv[i].x : jj == 1 ? pv[i].y : jj == 2 ? pv[i].z : pv[i].w;
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        o[i][h][0] = fmaf(p, vv[h].x, o[i][h][0]);
                        o[i][h][1] = fmaf(p, vv[h].y, o[i][h][1]);
                        o[i][h][2] = fmaf(p, vv[h].z, o[i][h][2]);
                        o[i][h][3] = fmaf(p, vv[h].w, o[i][h][3]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int qi = qt0 + ty + 16 * i;
        if (qi >= a.Sq) continue;
        const float inv = 1.0f / l[i];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            float4 r = make_float4(o[i][h][0] * inv, o[i][h][1] * inv, o[i][h][2] * inv, o[i][h][3] * inv);
            *reinterpret_cast<float4*>(a.out + (size_t)(a.q0 + qi) * a.o_tok_stride + (size_t)head * a.o_head_stride + tx * 4 + 64 * h) = r;
        }
    }
}

template <int HD>
inline size_t flash_smem_bytes() { return (size_t)(3 * 64 * (HD + 4) + 64 * 68) * sizeof(float); }

template <int HD>
inline void flash_attn(cudaStream_t st, const FlashArgs& a, int nheads, bool causal) {
    if (a.Sq == 0) return;
    const size_t smem = flash_smem_bytes<HD>();
    dim3 grid(ceil_div(a.Sq, 64), nheads);
    if (causal) {
        AHA_CUDA_CHECK(cudaFuncSetAttribute(flash_attn_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        flash_attn_kernel<HD, true><<<grid, 256, smem, st>>>(a);
    } else {
        AHA_CUDA_CHECK(cudaFuncSetAttribute(flash_attn_kernel<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        flash_attn_kernel<HD, false><<<grid, 256, smem, st>>>(a);
    }
    AHA_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------
// Decode attention (one new token), split over the KV length.  grid = (nsplit, nkv), 256 threads.
// Prologue per CTA: q heads of this kv group and the new k are RMS-normalised + rotated from the raw
// projections; the CTA whose token range holds the new position appends K,V to the paged cache.
// Each warp streams tokens of its range (one 512-byte K row + V row per token for hd=128), keeps an
// online softmax per q head, then the CTA publishes an (m, l, o[hd]) partial.  The last CTA to finish
// for a kv head (atomic ticket) merges the partials and writes the attention output row.
struct DecodeAttnArgs {
    const float* qkv;        // [(nh+2nkv)*hd] raw projections of the current token
    const float* qw; const float* kw; float eps;
    const float* inv_freq;
    const DecodeState* st;
    float* kbase; float* vbase;  // layer bases into the pool (writable)
    KVSrc kv;
    float* partial;          // [nh][nsplit][hd+2]
    int* counters;           // [nkv], zero on entry, reset on exit
    float* out;              // [nh*hd]
    int nh, nkv, nsplit;
    float scaling;
};

template <int HD, int G>
__device__ __forceinline__ void decode_attn_body(const DecodeAttnArgs& a, const int split, const int kvh) {
    static_assert(HD == 128, "decode attention is specialised for head_dim 128");
    constexpr int NW = 8;
    __shared__ __align__(16) float qs[G][HD];
    __shared__ __align__(16) float knew[HD];
    __shared__ __align__(16) float vnew[HD];
    __shared__ float sm_m[NW][G], sm_l[NW][G];
    __shared__ __align__(16) float sm_acc[NW][G][HD];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t_new = a.st->pos;            // cache index of the current token
    const int ctx = t_new + 1;
    const float rpos = (float)(t_new + a.st->rope_delta);

    // ---- prologue: norm + rope for q heads (warps 0..G-1), new k (warp G); v copy (warp G+1)
    if (warp <= G) {
        const bool is_q = warp < G;
        const float* src = a.qkv + (size_t)(is_q ? (kvh * G + warp) : (a.nh + kvh)) * HD;
        const float4 x = *reinterpret_cast<const float4*>(src + lane * 4);
        float ss = x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        ss = warp_sum(ss);
        const float inv = 1.0f / sqrtf(ss / (float)HD + a.eps);
        const float4 w = *reinterpret_cast<const float4*>((is_q ? a.qw : a.kw) + lane * 4);
        float n[4] = {x.x * inv * w.x, x.y * inv * w.y, x.z * inv * w.z, x.w * inv * w.w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float partner = __shfl_xor_sync(0xffffffffu, n[e], 16);  // element d +- 64
            const int d = lane * 4 + e;
            const int j = d & (HD / 2 - 1);
            const float ang = rpos * a.inv_freq[j];
            const float rot = (d < HD / 2) ? -partner : partner;
            o[e] = n[e] * cosf(ang) + rot * sinf(ang);
        }
        float* dst = is_q ? qs[warp] : knew;
        if (!is_q) { for (int e = 0; e < 4; ++e) o[e] = kv_store_round(o[e]); }
        *reinterpret_cast<float4*>(dst + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
    } else if (warp == G + 1) {
        float4 v = *reinterpret_cast<const float4*>(a.qkv + (size_t)(a.nh + a.nkv + kvh) * HD + lane * 4);
        v = make_float4(kv_store_round(v.x), kv_store_round(v.y), kv_store_round(v.z), kv_store_round(v.w));
        *reinterpret_cast<float4*>(vnew + lane * 4) = v;
    }
    __syncthreads();

    const int chunk = (ctx + a.nsplit - 1) / a.nsplit;
    const int lo = split * chunk, hi = min(ctx, lo + chunk);
    if (t_new >= lo && t_new < hi && tid < HD) {  // append to the cache
        const size_t off = a.kv.off(t_new, kvh) + tid;
        a.kbase[off] = knew[tid];
        a.vbase[off] = vnew[tid];
    }

    float4 q[G];
#pragma unroll
    for (int g = 0; g < G; ++g) q[g] = *reinterpret_cast<const float4*>(qs[g] + lane * 4);
    float m[G], l[G];
    float4 acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { m[g] = -INFINITY; l[g] = 0.f; acc[g] = make_float4(0.f, 0.f, 0.f, 0.f); }

    for (int t = lo + warp; t < hi; t += NW * 2) {
        float4 k0, v0, k1 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = k1;
        const int t1 = t + NW;
        const bool has1 = t1 < hi;
        if (t == t_new) { k0 = *reinterpret_cast<const float4*>(knew + lane * 4); v0 = *reinterpret_cast<const float4*>(vnew + lane * 4); }
        else { const size_t off = a.kv.off(t, kvh) + lane * 4; k0 = *reinterpret_cast<const float4*>(a.kv.k + off); v0 = *reinterpret_cast<const float4*>(a.kv.v + off); }
        if (has1) {
            if (t1 == t_new) { k1 = *reinterpret_cast<const float4*>(knew + lane * 4); v1 = *reinterpret_cast<const float4*>(vnew + lane * 4); }
            else { const size_t off = a.kv.off(t1, kvh) + lane * 4; k1 = *reinterpret_cast<const float4*>(a.kv.k + off); v1 = *reinterpret_cast<const float4*>(a.kv.v + off); }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float s0 = q[g].x * k0.x + q[g].y * k0.y + q[g].z * k0.z + q[g].w * k0.w;
            float s1 = q[g].x * k1.x + q[g].y * k1.y + q[g].z * k1.z + q[g].w * k1.w;
            s0 = warp_sum(s0) * a.scaling;
            s1 = has1 ? warp_sum(s1) * a.scaling : -INFINITY;
            const float mnew = fmaxf(m[g], fmaxf(s0, s1));
            const float alpha = expf(m[g] - mnew);
            const float p0 = expf(s0 - mnew), p1 = expf(s1 - mnew);
            l[g] = l[g] * alpha + p0 + p1;
            acc[g].x = acc[g].x * alpha + p0 * v0.x + p1 * v1.x;
            acc[g].y = acc[g].y * alpha + p0 * v0.y + p1 * v1.y;
            acc[g].z = acc[g].z * alpha + p0 * v0.z + p1 * v1.z;
            acc[g].w = acc[g].w * alpha + p0 * v0.w + p1 * v1.w;
            m[g] = mnew;
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (lane == 0) { sm_m[warp][g] = m[g]; sm_l[warp][g] = l[g]; }
        *reinterpret_cast<float4*>(&sm_acc[warp][g][lane * 4]) = acc[g];
    }
    __syncthreads();
    // merge the 8 warps -> partial for this split
    for (int idx = tid; idx < G * HD; idx += 256) {
        const int g = idx / HD, d = idx % HD;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; ++w) M = fmaxf(M, sm_m[w][g]);
        float L = 0.f, O = 0.f;
        if (M != -INFINITY) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float e = expf(sm_m[w][g] - M);
                L += sm_l[w][g] * e;
                O += sm_acc[w][g][d] * e;
            }
        }
        float* p = a.partial + ((size_t)(kvh * G + g) * a.nsplit + split) * (HD + 2);
        p[d] = O;
        if (d == 0) { p[HD] = M; p[HD + 1] = L; }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&a.counters[kvh], 1) == a.nsplit - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int idx = tid; idx < G * HD; idx += 256) {
        const int g = idx / HD, d = idx % HD;
        const float* pb = a.partial + (size_t)(kvh * G + g) * a.nsplit * (HD + 2);
        float M = -INFINITY;
        for (int s = 0; s < a.nsplit; ++s) M = fmaxf(M, __ldcg(pb + (size_t)s * (HD + 2) + HD));
        float L = 0.f, O = 0.f;
        for (int s = 0; s < a.nsplit; ++s) {
            const float ms = __ldcg(pb + (size_t)s * (HD + 2) + HD);
            if (ms == -INFINITY) continue;
            const float e = expf(ms - M);
            L += __ldcg(pb + (size_t)s * (HD + 2) + HD + 1) * e;
            O += __ldcg(pb + (size_t)s * (HD + 2) + d) * e;
        }
        a.out[(size_t)(kvh * G + g) * HD + d] = O / L;
    }
    if (tid == 0) a.counters[kvh] = 0;
}

template <int HD, int G>
__global__ void __launch_bounds__(256) decode_attn_kernel(DecodeAttnArgs a) { decode_attn_body<HD, G>(a, blockIdx.x, blockIdx.y); }

// The same for a batch of sequences decoded in lockstep (batch_decode.cuh): grid = (nsplit, nkv, sequences); sequence z takes its
// projections row, DecodeState, page table, partial / counter / output buffers from table[z].
template <int HD, int G>
__global__ void __launch_bounds__(256) decode_attn_batch_kernel(const DecodeAttnArgs* __restrict__ table) {
    const DecodeAttnArgs a = table[blockIdx.z];
    decode_attn_body<HD, G>(a, blockIdx.x, blockIdx.y);
}

}  // namespace aha
