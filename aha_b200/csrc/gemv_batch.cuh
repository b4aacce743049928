// gemv_batch.cuh -- decode projections for a BATCH of sequences: Y[b][N] = W[N,K] (fp16) . X[b][K] (fp32), b < nb <= 8.
//
// The reference serves one request at a time (/root/reference/src/server/api.rs:117) and its Linear::forward
// (/root/reference/src/models/common/modules.rs:81-87,538-577) sees one row per decode step.  At batch 1 a decode step is
// pure weight streaming (gemv.cuh, decode_fused.cuh); decoding nb sequences in lockstep reads every weight ONCE for all of
// them, so the bytes per generated token fall by nb.  This kernel is gemv_kernel with nb activation rows:
//   * every warp streams RPW weight rows (16-byte L1-bypassing loads), each lane owning the same 8-element chunks c = lane,
//     lane + 32, ... as in gemv_kernel, and keeps RPW x NB fp32 accumulators;
//   * the activations are staged in shared memory KC = 2048 elements at a time (NB x 8 KB), so K = 6144 (down projection)
//     needs the same 64 KB as K = 2048 and several CTAs stay resident per SM;
//   * per (row, sequence) the products are added in exactly the order of gemv_kernel (chunks ascending per lane, then the
//     warp tree), so a batched step reproduces the single-sequence per-op step bit for bit;
//   * fused RMSNorm prologue (per sequence) and bias / residual / SwiGLU epilogues as in gemv.cuh.
#pragma once
#include "common.cuh"
#include "gemv.cuh"

namespace aha {

constexpr int kGemvBatchMax = 8;      // sequences per step
constexpr int kGemvBatchKC = 2048;    // activation elements per sequence staged per pass

struct GemvBatchArgs {
    const __half* W;       // [N, K]
    const float* x; int ldx;        // [nb][K] rows, ldx floats apart
    const float* norm_w;   // [K] (PRO_RMSNORM)
    float eps;
    const float* bias;     // [N] or nullptr
    const float* resid; int ldr;    // [nb][N] (GEPI_RESID; may alias out)
    float* out; int ldo;            // [nb][N] (or [nb][N/2] for SWIGLU)
    int N, K, nb;
};

template <int NB, int RPW, int PRO, int EPI>
__global__ void __launch_bounds__(256) gemv_batch_kernel(GemvBatchArgs a) {
    extern __shared__ __align__(16) float xs[];     // [NB][KC]
    __shared__ float red[32];
    __shared__ float s_inv[NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = a.K, nb = a.nb;

    if (PRO == PRO_RMSNORM) {   // 1 / rms of every sequence's row, summed in gemv_kernel's order (thread-strided float4s, then block_sum)
        for (int b = 0; b < NB; ++b) {
            if (b >= nb) break;
            const float* xr = a.x + (size_t)b * a.ldx;
            float ss = 0.f;
            for (int i = tid * 4; i < K; i += 256 * 4) {
                const float4 v = *reinterpret_cast<const float4*>(xr + i);
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
            ss = block_sum(ss, red);
            if (tid == 0) s_inv[b] = 1.0f / sqrtf(ss / (float)K + a.eps);
        }
        __syncthreads();
    }

    const int row0 = (blockIdx.x * 8 + warp) * RPW;
    float acc[RPW][NB];
    const __half* wr[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        wr[r] = a.W + (size_t)min(row0 + r, a.N - 1) * K;
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;
    }

    for (int k0 = 0; k0 < K; k0 += kGemvBatchKC) {
        const int kc = min(kGemvBatchKC, K - k0);
        __syncthreads();   // the previous pass has been consumed
#pragma unroll
        for (int b = 0; b < NB; ++b) {   // rows beyond nb are staged as zeros: the inner loop then needs no predicate (their results are never stored)
            const float* xr = a.x + (size_t)(b < nb ? b : 0) * a.ldx + k0;
            const float inv = (PRO == PRO_RMSNORM && b < nb) ? s_inv[b] : 1.0f;
            for (int i = tid * 4; i < kc; i += 256 * 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (b < nb) {
                    v = *reinterpret_cast<const float4*>(xr + i);
                    if (PRO == PRO_RMSNORM) {
                        const float4 w = *reinterpret_cast<const float4*>(a.norm_w + k0 + i);
                        v.x = v.x * inv * w.x; v.y = v.y * inv * w.y; v.z = v.z * inv * w.z; v.w = v.w * inv * w.w;
                    }
                }
                *reinterpret_cast<float4*>(xs + (size_t)b * kGemvBatchKC + i) = v;
            }
        }
        __syncthreads();
        const int nchunk = kc >> 3;
        for (int c = lane; c < nchunk; c += 32) {
            // the 8 weights of every row are widened to fp32 ONCE and then meet all NB sequences (dot8 would convert them again per sequence);
            // per (row, sequence) the eight fmas run in dot8's order, so the sums are those of gemv_kernel bit for bit
            float wf[RPW][8];
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const uint4 w = ldg_stream(wr[r] + k0 + c * 8);
                const float2 p0 = h2_to_f2(w.x), p1 = h2_to_f2(w.y), p2 = h2_to_f2(w.z), p3 = h2_to_f2(w.w);
                wf[r][0] = p0.x; wf[r][1] = p0.y; wf[r][2] = p1.x; wf[r][3] = p1.y; wf[r][4] = p2.x; wf[r][5] = p2.y; wf[r][6] = p3.x; wf[r][7] = p3.y;
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const float4 x0 = *reinterpret_cast<const float4*>(xs + (size_t)b * kGemvBatchKC + c * 8);
                const float4 x1 = *reinterpret_cast<const float4*>(xs + (size_t)b * kGemvBatchKC + c * 8 + 4);
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    float s = acc[r][b];
                    s = fmaf(wf[r][0], x0.x, s); s = fmaf(wf[r][1], x0.y, s); s = fmaf(wf[r][2], x0.z, s); s = fmaf(wf[r][3], x0.w, s);
                    s = fmaf(wf[r][4], x1.x, s); s = fmaf(wf[r][5], x1.y, s); s = fmaf(wf[r][6], x1.z, s); s = fmaf(wf[r][7], x1.w, s);
                    acc[r][b] = s;
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[r][b] = warp_sum(acc[r][b]);

    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b >= nb) break;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int row = row0 + r;
                if (row >= a.N) break;
                float v = acc[r][b] + (a.bias ? a.bias[row] : 0.f);
                if (EPI == GEPI_SWIGLU) {
                    if ((r & 1) == 0) {   // rows (2i, 2i+1) = (gate_i, up_i)
                        const float up = acc[r + 1 < RPW ? r + 1 : r][b] + (a.bias ? a.bias[row + 1] : 0.f);
                        a.out[(size_t)b * a.ldo + (row >> 1)] = silu_f(v) * up;
                    }
                } else {
                    if (EPI == GEPI_RESID) v += a.resid[(size_t)b * a.ldr + row];
                    a.out[(size_t)b * a.ldo + row] = v;
                }
            }
        }
    }
}

// ---- version 2: the same arithmetic with the WEIGHTS prefetched through a per-thread shared-memory ring (cp.async, 16 bytes per row and stage) ----
// Version 1 holds the weights of one iteration in registers, so a warp has RPW 16-byte loads in flight and a step of 8 sequences is bound by bytes in
// flight per SM (call 22: 4.3 ms per step).  Here every thread runs STAGES - 1 iterations ahead: it copies the 16 bytes IT will consume (row r, chunk
// c = t * 32 + lane) straight into its own slot of the ring and reads them back after cp.async.wait_group -- no cross-thread dependency, so no barrier
// guards the ring, and 2 CTAs x 256 threads x 3 stages x 4 rows x 16 B = 96 KB are in flight per SM instead of 32.  Activations are staged 1024
// elements per sequence at a time (32 KB for 8 sequences) so that two CTAs fit one SM.  Per (row, sequence) the fmas run in version 1's order.
constexpr int kGemvRingKC = 1024;
constexpr int kGemvRingStages = 4;

__device__ __forceinline__ void gb_cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void gb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void gb_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NB, int RPW, int PRO, int EPI>
__global__ void __launch_bounds__(256, 2) gemv_batch_ring_kernel(GemvBatchArgs a) {
    constexpr int ST = kGemvRingStages, KC = kGemvRingKC;
    extern __shared__ __align__(16) float dsm[];
    float* xs = dsm;                                                              // [NB][KC]
    uint4* ring = reinterpret_cast<uint4*>(dsm + (size_t)NB * KC);               // [ST][RPW][256]
    __shared__ float red[32];
    __shared__ float s_inv[NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = a.K, nb = a.nb;

    const int row0 = (blockIdx.x * 8 + warp) * RPW;
    const __half* wr[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) wr[r] = a.W + (size_t)min(row0 + r, a.N - 1) * K;
    const int nch = K >> 3;                      // 8-element chunks per row
    const int T = (nch + 31) >> 5;               // iterations: chunk c = t * 32 + lane
    auto issue = [&](int t) {
        const int c = t * 32 + lane;
        if (t < T && c < nch) {
#pragma unroll
            for (int r = 0; r < RPW; ++r) gb_cp_async16(&ring[((t % ST) * RPW + r) * 256 + tid], wr[r] + (size_t)c * 8);
        }
        gb_cp_async_commit();                    // one group per iteration for every thread, copies or not: the group arithmetic stays uniform
    };
#pragma unroll
    for (int t = 0; t < ST - 1; ++t) issue(t);   // the weight stream starts before the activations are even looked at

    if (PRO == PRO_RMSNORM) {
        for (int b = 0; b < NB; ++b) {
            if (b >= nb) break;
            const float* xr = a.x + (size_t)b * a.ldx;
            float ss = 0.f;
            for (int i = tid * 4; i < K; i += 256 * 4) {
                const float4 v = *reinterpret_cast<const float4*>(xr + i);
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
            ss = block_sum(ss, red);
            if (tid == 0) s_inv[b] = 1.0f / sqrtf(ss / (float)K + a.eps);
        }
        __syncthreads();
    }

    float acc[RPW][NB];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;

    constexpr int IT_PER_PASS = KC / 256;        // iterations that share one staged activation pass
    for (int t = 0; t < T; ++t) {
        issue(t + ST - 1);
        if (t % IT_PER_PASS == 0) {              // new pass: stage the next KC activations of every sequence (uniform across the CTA)
            const int k0 = (t / IT_PER_PASS) * KC, kc = min(KC, K - k0);
            __syncthreads();                     // the previous pass has been consumed
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const float* xr = a.x + (size_t)(b < nb ? b : 0) * a.ldx + k0;
                const float inv = (PRO == PRO_RMSNORM && b < nb) ? s_inv[b] : 1.0f;
                for (int i = tid * 4; i < kc; i += 256 * 4) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (b < nb) {
                        v = *reinterpret_cast<const float4*>(xr + i);
                        if (PRO == PRO_RMSNORM) {
                            const float4 w = *reinterpret_cast<const float4*>(a.norm_w + k0 + i);
                            v.x = v.x * inv * w.x; v.y = v.y * inv * w.y; v.z = v.z * inv * w.z; v.w = v.w * inv * w.w;
                        }
                    }
                    *reinterpret_cast<float4*>(xs + (size_t)b * KC + i) = v;
                }
            }
            __syncthreads();
        }
        gb_cp_async_wait<ST - 1>();              // this thread's copies of iteration t have landed (groups complete in order)
        const int c = t * 32 + lane;
        if (c < nch) {
            const int xo = (c * 8) % KC;         // offset of the chunk inside the staged pass
            float wf[RPW][8];
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const uint4 w = ring[((t % ST) * RPW + r) * 256 + tid];
                const float2 p0 = h2_to_f2(w.x), p1 = h2_to_f2(w.y), p2 = h2_to_f2(w.z), p3 = h2_to_f2(w.w);
                wf[r][0] = p0.x; wf[r][1] = p0.y; wf[r][2] = p1.x; wf[r][3] = p1.y; wf[r][4] = p2.x; wf[r][5] = p2.y; wf[r][6] = p3.x; wf[r][7] = p3.y;
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const float4 x0 = *reinterpret_cast<const float4*>(xs + (size_t)b * KC + xo);
                const float4 x1 = *reinterpret_cast<const float4*>(xs + (size_t)b * KC + xo + 4);
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    float s = acc[r][b];
                    s = fmaf(wf[r][0], x0.x, s); s = fmaf(wf[r][1], x0.y, s); s = fmaf(wf[r][2], x0.z, s); s = fmaf(wf[r][3], x0.w, s);
                    s = fmaf(wf[r][4], x1.x, s); s = fmaf(wf[r][5], x1.y, s); s = fmaf(wf[r][6], x1.z, s); s = fmaf(wf[r][7], x1.w, s);
                    acc[r][b] = s;
                }
            }
        }
    }
    gb_cp_async_wait<0>();
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[r][b] = warp_sum(acc[r][b]);

    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b >= nb) break;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int row = row0 + r;
                if (row >= a.N) break;
                float v = acc[r][b] + (a.bias ? a.bias[row] : 0.f);
                if (EPI == GEPI_SWIGLU) {
                    if ((r & 1) == 0) {
                        const float up = acc[r + 1 < RPW ? r + 1 : r][b] + (a.bias ? a.bias[row + 1] : 0.f);
                        a.out[(size_t)b * a.ldo + (row >> 1)] = silu_f(v) * up;
                    }
                } else {
                    if (EPI == GEPI_RESID) v += a.resid[(size_t)b * a.ldr + row];
                    a.out[(size_t)b * a.ldo + row] = v;
                }
            }
        }
    }
}

template <int NB, int RPW, int PRO, int EPI>
inline void gemv_batch_ring_launch_one(cudaStream_t st, const GemvBatchArgs& a) {
    static bool attr_set[64] = {};
    const size_t smem = (size_t)NB * kGemvRingKC * sizeof(float) + (size_t)kGemvRingStages * RPW * 256 * sizeof(uint4);
    int dev = 0;
    AHA_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        AHA_CUDA_CHECK(cudaFuncSetAttribute(gemv_batch_ring_kernel<NB, RPW, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    gemv_batch_ring_kernel<NB, RPW, PRO, EPI><<<ceil_div(a.N, 8 * RPW), 256, smem, st>>>(a);
    AHA_CUDA_CHECK(cudaGetLastError());
}

template <int NB, int RPW, int PRO, int EPI>
inline void gemv_batch_launch_one(cudaStream_t st, const GemvBatchArgs& a) {
    // 64 KB of dynamic shared memory for NB = 8: opt in once per (instantiation, device) -- the attribute is per device, so the flag is too
    static bool attr_set[64] = {};
    const size_t smem = (size_t)NB * kGemvBatchKC * sizeof(float);
    if (smem > 48 * 1024) {
        int dev = 0;
        AHA_CUDA_CHECK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            AHA_CUDA_CHECK(cudaFuncSetAttribute(gemv_batch_kernel<NB, RPW, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    gemv_batch_kernel<NB, RPW, PRO, EPI><<<ceil_div(a.N, 8 * RPW), 256, smem, st>>>(a);
    AHA_CUDA_CHECK(cudaGetLastError());
}

template <int PRO, int EPI>
inline void gemv_batch_launch(cudaStream_t st, const GemvBatchArgs& a, bool ring) {
    if (ring) {   // version 2 (weights through the cp.async ring): same rows-per-warp rule, two batch widths
        const bool small = a.nb <= 4;
        if (a.N >= 4096) { if (small) gemv_batch_ring_launch_one<4, 4, PRO, EPI>(st, a); else gemv_batch_ring_launch_one<8, 4, PRO, EPI>(st, a); }
        else { if (small) gemv_batch_ring_launch_one<4, 2, PRO, EPI>(st, a); else gemv_batch_ring_launch_one<8, 2, PRO, EPI>(st, a); }
        return;
    }
    // rows per warp: with nb sequences every staged activation chunk costs nb shared-memory reads per weight chunk, so reuse across rows matters
    // more than at batch 1 (call 21: 5.4 ms per step of 8 sequences with gemv.cuh's choice): 4 rows per warp from N = 4096 (128 CTAs), else 2
    const int rpw = a.N >= 4096 ? 4 : 2;
    const bool small = a.nb <= 4;
    if (rpw >= 4) { if (small) gemv_batch_launch_one<4, 4, PRO, EPI>(st, a); else gemv_batch_launch_one<8, 4, PRO, EPI>(st, a); }
    else if (rpw == 2 || EPI == GEPI_SWIGLU) { if (small) gemv_batch_launch_one<4, 2, PRO, EPI>(st, a); else gemv_batch_launch_one<8, 2, PRO, EPI>(st, a); }
    else { if (small) gemv_batch_launch_one<4, 1, PRO, EPI>(st, a); else gemv_batch_launch_one<8, 1, PRO, EPI>(st, a); }
}

inline void gemv_batch(cudaStream_t st, int pro, int epi, const GemvBatchArgs& a, bool ring = false) {
    AHA_REQUIRE(a.nb >= 1 && a.nb <= kGemvBatchMax, "gemv_batch: 1..8 sequences");
    AHA_REQUIRE(a.K % 8 == 0 && a.ldx % 4 == 0, "gemv_batch: K must be a multiple of 8 and rows 16-byte aligned");
    AHA_REQUIRE(epi != GEPI_SWIGLU || a.N % 2 == 0, "gemv_batch: SwiGLU needs an even row count");
#define AHA_GEMVB_CASE(P, E) if (pro == P && epi == E) { gemv_batch_launch<P, E>(st, a, ring); return; }
    AHA_GEMVB_CASE(PRO_NONE, GEPI_STORE)
    AHA_GEMVB_CASE(PRO_NONE, GEPI_RESID)
    AHA_GEMVB_CASE(PRO_RMSNORM, GEPI_STORE)
    AHA_GEMVB_CASE(PRO_RMSNORM, GEPI_SWIGLU)
#undef AHA_GEMVB_CASE
    AHA_REQUIRE(false, "gemv_batch: unsupported prologue/epilogue combination");
}

}  // namespace aha
