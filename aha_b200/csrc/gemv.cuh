// gemv.cuh -- batch-1 decode projections: y[N] = W[N,K] (fp16) . x[K] (fp32), fp32 accumulate.
// HBM-bound (1 FLOP per weight byte); per-op version with the RMSNorm prologue and the residual /
// SwiGLU / argmax epilogues fused so that a decoder layer is 5 launches:
//   [rmsnorm+qkv] [attention] [o_proj+residual] [rmsnorm+gate/up+SwiGLU] [down+residual]
// Reference ops replaced: RmsNorm + 3 Linear (modules.rs:538-553), o_proj (:577), GateUpDownMLP
// (modules.rs:81-87), final norm + lm_head (qwen3/model.rs:142,186-187), ArgMax sampler (sample.rs:13-22).
#pragma once
#include "common.cuh"
#include "kernels_common.cuh"

namespace aha {

enum GemvPro { PRO_NONE = 0, PRO_RMSNORM = 1 };
enum GemvEpi { GEPI_STORE = 0, GEPI_RESID = 1, GEPI_SWIGLU = 2, GEPI_ARGMAX = 3 };

struct GemvArgs {
    const __half* W;       // [N, K]
    const float* x;        // [K]
    const float* norm_w;   // [K] (PRO_RMSNORM)
    float eps;
    const float* bias;     // [N] or nullptr
    const float* resid;    // [N] (GEPI_RESID; may alias out)
    float* out;            // [N] (or [N/2] for SWIGLU)
    float* pmax; int* pidx;  // per-block argmax candidates (GEPI_ARGMAX)
    int N, K;
};

template <int RPW, int PRO, int EPI>
__global__ void __launch_bounds__(256) gemv_kernel(GemvArgs a) {
    extern __shared__ __align__(16) float xs[];
    __shared__ float red[32];
    __shared__ float smax[8];
    __shared__ int sidx[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = a.K;
    float ss = 0.f;
    for (int i = tid * 4; i < K; i += 256 * 4) {
        float4 v = *reinterpret_cast<const float4*>(a.x + i);
        *reinterpret_cast<float4*>(xs + i) = v;
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (PRO == PRO_RMSNORM) {
        ss = block_sum(ss, red);
        const float inv = 1.0f / sqrtf(ss / (float)K + a.eps);
        for (int i = tid * 4; i < K; i += 256 * 4) {  // each thread rescales exactly what it loaded
            float4 v = *reinterpret_cast<float4*>(xs + i);
            const float4 w = *reinterpret_cast<const float4*>(a.norm_w + i);
            v.x = v.x * inv * w.x; v.y = v.y * inv * w.y; v.z = v.z * inv * w.z; v.w = v.w * inv * w.w;
            *reinterpret_cast<float4*>(xs + i) = v;
        }
    }
    __syncthreads();

    const int row0 = (blockIdx.x * 8 + warp) * RPW;
    float acc[RPW];
    const __half* wr[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        acc[r] = 0.f;
        wr[r] = a.W + (size_t)min(row0 + r, a.N - 1) * K;
    }
    const int nchunk = K >> 3;
#pragma unroll 4
    for (int c = lane; c < nchunk; c += 32) {
        const float4 x0 = *reinterpret_cast<const float4*>(xs + c * 8);
        const float4 x1 = *reinterpret_cast<const float4*>(xs + c * 8 + 4);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const uint4 w = ldg_stream(wr[r] + c * 8);
            acc[r] = dot8(w, x0, x1, acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) acc[r] = warp_sum(acc[r]);

    float best = -INFINITY;
    int bi = 0x7fffffff;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int row = row0 + r;
            if (row >= a.N) break;
            float v = acc[r] + (a.bias ? a.bias[row] : 0.f);
            if (EPI == GEPI_SWIGLU) {
                if ((r & 1) == 0) {  // rows (2i, 2i+1) = (gate_i, up_i)
                    const float up = acc[r + 1 < RPW ? r + 1 : r] + (a.bias ? a.bias[row + 1] : 0.f);
                    a.out[row >> 1] = silu_f(v) * up;
                }
            } else {
                if (EPI == GEPI_RESID) v += a.resid[row];
                a.out[row] = v;
                if (EPI == GEPI_ARGMAX && (v > best)) { best = v; bi = row; }
            }
        }
    }
    if (EPI == GEPI_ARGMAX) {
        if (lane == 0) { smax[warp] = best; sidx[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; ++w)
                if (smax[w] > best || (smax[w] == best && sidx[w] < bi)) { best = smax[w]; bi = sidx[w]; }
            a.pmax[blockIdx.x] = best;
            a.pidx[blockIdx.x] = bi;
        }
    }
}

// Rows per warp: enough CTAs to cover the 148 SMs at >= 2 CTAs each, otherwise as many rows per warp as
// possible so the activation chunk read from shared memory is reused.
inline int gemv_pick_rpw(int N, bool need_pairs) {
    if (N / 32 >= 296) return 4;
    if (N / 16 >= 296 || need_pairs) return 2;
    return 1;
}

template <int PRO, int EPI>
inline void gemv_launch_rpw(cudaStream_t st, const GemvArgs& a, int rpw) {
    const size_t smem = (size_t)a.K * sizeof(float);
    const int grid = ceil_div(a.N, 8 * rpw);
    switch (rpw) {
        case 4: gemv_kernel<4, PRO, EPI><<<grid, 256, smem, st>>>(a); break;
        case 2: gemv_kernel<2, PRO, EPI><<<grid, 256, smem, st>>>(a); break;
        default:
            if (EPI == GEPI_SWIGLU) { gemv_kernel<2, PRO, EPI><<<ceil_div(a.N, 16), 256, smem, st>>>(a); }
            else gemv_kernel<1, PRO, EPI><<<grid, 256, smem, st>>>(a);
    }
    AHA_CUDA_CHECK(cudaGetLastError());
}

template <int RPW, int PRO, int EPI>
inline void gemv_set_attr() {
    AHA_CUDA_CHECK(cudaFuncSetAttribute(gemv_kernel<RPW, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
}
template <int PRO, int EPI>
inline void gemv_set_attr_all() { gemv_set_attr<1, PRO, EPI>(); gemv_set_attr<2, PRO, EPI>(); gemv_set_attr<4, PRO, EPI>(); }
// Opt in to > 48 KB of shared memory (K up to 16384 activations staged per CTA).  Call once per device.
inline void gemv_init() {
    gemv_set_attr_all<PRO_NONE, GEPI_STORE>();
    gemv_set_attr_all<PRO_NONE, GEPI_RESID>();
    gemv_set_attr_all<PRO_RMSNORM, GEPI_STORE>();
    gemv_set_attr_all<PRO_RMSNORM, GEPI_SWIGLU>();
    gemv_set_attr_all<PRO_RMSNORM, GEPI_ARGMAX>();
}

inline int gemv_grid(int N, int epi) { return ceil_div(N, 8 * gemv_pick_rpw(N, epi == GEPI_SWIGLU)); }

inline void gemv(cudaStream_t st, int pro, int epi, const GemvArgs& a) {
    AHA_REQUIRE(a.K % 8 == 0 && a.K * sizeof(float) <= 64 * 1024, "gemv: K must be a multiple of 8 and <= 16384");
    AHA_REQUIRE(epi != GEPI_SWIGLU || a.N % 2 == 0, "gemv: SwiGLU needs an even row count");
    const int rpw = gemv_pick_rpw(a.N, epi == GEPI_SWIGLU);
#define AHA_GEMV_CASE(P, E) if (pro == P && epi == E) { gemv_launch_rpw<P, E>(st, a, rpw); return; }
    AHA_GEMV_CASE(PRO_NONE, GEPI_STORE)
    AHA_GEMV_CASE(PRO_NONE, GEPI_RESID)
    AHA_GEMV_CASE(PRO_RMSNORM, GEPI_STORE)
    AHA_GEMV_CASE(PRO_RMSNORM, GEPI_SWIGLU)
    AHA_GEMV_CASE(PRO_RMSNORM, GEPI_ARGMAX)
#undef AHA_GEMV_CASE
    AHA_REQUIRE(false, "gemv: unsupported prologue/epilogue combination");
}

}  // namespace aha
