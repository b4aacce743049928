// common.cuh -- shared host/device helpers for libaha_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#define AHA_CUDA_CHECK(expr)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) +      \
                                     " at " + __FILE__ + ":" + std::to_string(__LINE__) + " (" + \
                                     #expr + ")");                                                \
        }                                                                                         \
    } while (0)

#define AHA_REQUIRE(cond, msg)                                                      \
    do {                                                                            \
        if (!(cond)) throw std::runtime_error(std::string("aha_b200: ") + (msg));   \
    } while (0)

namespace aha {

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; `red` must hold >= 32 floats of shared memory.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` from a previous use
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}

// 16-byte streaming load of weights (read once per step): bypass L1.
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float2 h2_to_f2(uint32_t u) {
    __half2 h = *reinterpret_cast<__half2*>(&u);
    return __half22float2(h);
}

// dot of 8 fp16 weights (one uint4) with 8 fp32 activations, fp32 accumulate, exact products.
__device__ __forceinline__ float dot8(const uint4& w, const float4& x0, const float4& x1, float acc) {
    float2 a = h2_to_f2(w.x), b = h2_to_f2(w.y), c = h2_to_f2(w.z), d = h2_to_f2(w.w);
    acc = fmaf(a.x, x0.x, acc); acc = fmaf(a.y, x0.y, acc);
    acc = fmaf(b.x, x0.z, acc); acc = fmaf(b.y, x0.w, acc);
    acc = fmaf(c.x, x1.x, acc); acc = fmaf(c.y, x1.y, acc);
    acc = fmaf(d.x, x1.z, acc); acc = fmaf(d.y, x1.w, acc);
    return acc;
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_tanh_f(float x) {
    const float c = 0.7978845608028654f;  // sqrt(2/pi)
    return 0.5f * x * (1.f + tanhf(c * (x + 0.044715f * x * x * x)));
}

enum Act { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU_ERF = 2, ACT_GELU_TANH = 3 };
__device__ __forceinline__ float apply_act(int act, float x) {
    switch (act) {
        case ACT_SILU: return silu_f(x);
        case ACT_GELU_ERF: return gelu_erf_f(x);
        case ACT_GELU_TANH: return gelu_tanh_f(x);
        default: return x;
    }
}

// Measurement build only (-DAHA_KV_ROUND_FP16, variants/kv16.so): every K / V value is rounded to fp16 precision on its way into the
// (still fp32) cache, which is numerically what an fp16 KV cache would hold.  The default build compiles this to the identity.
#ifdef AHA_KV_ROUND_FP16
__device__ __forceinline__ float kv_store_round(float x) { return __half2float(__float2half_rn(x)); }
#else
__device__ __forceinline__ float kv_store_round(float x) { return x; }
#endif

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace aha
