// decode_fused.cuh -- the whole per-token decode step as ONE persistent, warp-specialised kernel.
//
// Why: at batch 1 the step is pure HBM streaming (3.4 GB of fp16 weights + the KV cache per token for
// Qwen3-VL-2B, ~1 FLOP/byte) cut into ~140 dependent pieces of 8-50 MB; as separate launches every piece pays
// launch + pipeline ramp and the memory system idles between them.  Here one CTA per SM stays resident for
// the whole step:
//   * the last warp (one elected lane) is the PRODUCER: it walks this CTA's static schedule of transfers for all
//     layers -- weight row slabs and paged-KV half pages -- and issues TMA bulk copies
//     (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes, SASS UBLKCP, L2 evict-first policy) into
//     a ring of 16 KB shared-memory stages guarded by full/empty mbarriers.  It never waits for activations, so
//     weights for the next phases keep streaming while the consumers sit in a grid barrier;
//   * the other 11 warps are CONSUMERS: per phase they stage the activation vector in shared memory (RMSNorm
//     fused), and each warp consumes the ring slots it owns: a whole stage of weight rows is reduced by one warp
//     with fp32 dot products (fp16 weights converted exactly) and the fused epilogue (residual add from the rows
//     of the residual stream the CTA keeps on chip, SwiGLU, argmax) is applied by its lanes; attention is split-KV
//     with per-warp online softmax over the half pages the warp owns, q/k RMSNorm + RoPE + KV append fused in, and
//     the merge of the splits folded into the o_proj input load;
//   * phases are separated by a grid-wide barrier (one counter in L2, release/acquire) that only the consumers join.
// Phases per layer: [rmsnorm+qkv] -> [attention] -> [o_proj+residual] -> [rmsnorm+gate/up+SwiGLU] ->
// [down+residual]; then [final norm + lm_head + argmax] and the on-device token feedback.
//
// Variant KS (template flag, opt-in through aha_options.decode_impl = 3; NOT the default): the down projection is
// computed K-split right behind gate/up -- CTA c multiplies the k-rows of Wdown^T that match its own slice of
// silu(gate)*up (still in shared memory) into all H outputs and adds them into a global fp32 accumulator with
// red.global.add.f32 -- so the grid barrier and the h round trip between the two disappear (4 barriers per layer).
// The next consumer of the residual stream adds the accumulator while it loads x.  Summation order across CTAs is
// then run-dependent (fp32 atomics).
//
// Variant KO (template flag, decode_impl = 4, or 5 together with KS; NOT the default): o_proj is computed K-split per
// kv group right behind the attention.  The grid barrier between the two becomes a barrier over the nsplit CTAs of
// the group, each CTA merges only its own group's partials (G heads instead of all of them) and multiplies its slice
// of the rows of that group's column block of Wo, adding the result into the residual stream with fp32 reductions.
//
// Reference semantics are those of Qwen3DecoderLayer::forward / QKNormAttention::forward / GateUpDownMLP
// (/root/reference/src/models/qwen3/model.rs:71-87, src/models/common/modules.rs:81-87,530-579,757-813)
// with seq_len = 1; arithmetic is fp32 throughout, identical to the per-op kernels in gemv.cuh/attention.cuh.
#pragma once
#include <cooperative_groups.h>

#include "attention.cuh"
#include "common.cuh"
#include "kernels_common.cuh"

namespace aha {

constexpr int kFusedStages = 11;                         // ring slots (all in use by default: one per consumer warp)
constexpr int kFusedStageBytes = 16384;
constexpr int kFusedConsumers = 11;                      // consumer warps (+1 producer = 12 warps: register allocation granularity)
constexpr int kFusedThreads = (kFusedConsumers + 1) * 32;
constexpr int kFusedMaxRows = 8;                         // weight rows per stage (<= 16 KB)
constexpr int kFusedMaxK = 8192;                         // activation vector staged in shared memory (32 KB)
constexpr int kHalfPage = 16;                            // tokens per attention stage (K 8 KB + V 8 KB)
constexpr int kFusedTraceWords = 2 * 4096 + 256 * 256;     // timing traces: CTA 0 consumer/producer stamps + [cta][256] barrier arrivals
constexpr int kFusedMaxOwnRows = 64;                     // residual-stream rows owned by one CTA (H / grid, rounded up)
constexpr int kFusedPartialStride = 128 + 4;             // floats per (head, split) attention partial: acc[128], m, l, pad (16-byte rows)
constexpr int kFusedMergeChunk = 20;                     // splits merged per pass when the o_proj input is assembled
constexpr int kFusedMaxHs = 128;                         // SwiGLU outputs one CTA keeps in shared memory (K-split variant): I / grid, rounded up
constexpr int kFusedMaxPages = 512;                      // page-table entries staged in shared memory (16K tokens)

struct FusedLayer {
    const __half *qkv, *o, *gu, *down;
    const float *qkv_b, *o_b;
    const float *ln1, *ln2, *qn, *kn;
    const __half* down_t;   // [I][H] transposed copy of `down` (variant KS only, else nullptr)
    const __half* o_g;      // [nkv][H][G*hd] column blocks of `o`, one per kv group (variant KO only, else nullptr)
};

struct FusedArgs {
    const FusedLayer* layers;   // device array [L]
    int L, H, I, nh, nkv, hd, V, qkv_dim;
    float eps, scaling;
    const __half* embed;
    const __half* lm_head;
    const float* final_norm;
    const float* inv_freq;
    DecodeState* st;
    float* x;          // [H]   residual stream
    float* qkv1;       // [qkv_dim]
    float* h1;         // [I]
    float* logits;     // [V]
    float* partial;    // [nh][nsplit][kFusedPartialStride]
    float* xo;         // [2][H] variant KO: residual stream the K-split o_proj accumulates into, by layer parity (owners write x, CTAs add)
    unsigned* gsync;   // [nkv] variant KO: arrival counters of the per-kv-group barriers (zeroed by the host before launch)
    float* acc2;       // [2][H] fp32 accumulators of the K-split down projection (variant KS), zeroed by the host before launch
    unsigned* sync;    // [2] grid-barrier counter, final ticket   (zeroed by the host before launch)
    float* pmax; int* pidx;  // [grid] per-CTA argmax candidates
    uint32_t* argmax_out;
    uint32_t* history; int hist_cap;
    float* kv_pool; size_t layer_stride, page_stride;
    const int* page_table;
    int nsplit;
    int stages;        // ring slots in use (<= kFusedStages)
    int dbg;           // timing experiments only: bit0 = skip grid barriers, bit1 = skip the GEMV math, bit2 = timestamps, bit3 = old full-fence grid barrier, bit6 = per-CTA barrier arrival stamps, bit4 / bit5 = drop the L2 evict-first hint on weight / KV copies
    unsigned long long* trace;  // [2][4096] globaltimer stamps of CTA 0 (consumer thread 0 / producer), dbg bit2
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, 16-byte aligned).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define AHA_STAMP(a, who, idx) do { if (((a).dbg & 4) && blockIdx.x == 0 && (idx) < 4096) (a).trace[(who) * 4096 + (idx)++] = gtime(); } while (0)
// Same copy with an L2 evict-first policy: weights are read exactly once per step, so they should not displace the
// activations / partials / barrier words that every CTA re-reads from L2.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kFusedConsumers * 32) : "memory"); }
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Lighter synchronisation primitives than __threadfence() (= MEMBAR.SC.GPU + CCTL.IVALL per call on sm_100): a release
// reduction / acq_rel atomic issued by ONE thread after a CTA barrier publishes the whole CTA's writes (release is
// cumulative over the bar.sync), and a relaxed poll + fence.acq_rel is the matching acquire.  Cross-CTA data is always
// read with ld.global.cg, so no L1 line can be stale.
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned atom_acq_rel_add(unsigned* p, unsigned v) {
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Grid barrier joined by the consumer threads only: one arrival counter in L2 (zeroed by the host before every
// launch), polled by consumer thread 0 with relaxed loads; activations are always read with ld.global.cg.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& seq, int dbg = 0, unsigned long long* trace = nullptr) {
    if (dbg & 1) { consumer_bar_sync(); return; }
    consumer_bar_sync();                       // every consumer thread of this CTA has issued its writes
    if (threadIdx.x == 0) {
        if ((dbg & 64) && seq < 255) trace[8192 + blockIdx.x * 256 + seq] = gtime();   // per-CTA arrival times (skew analysis)
        if (dbg & 8) {                         // A/B: the old full-fence barrier
            __threadfence();
            atomicAdd(counter, 1u);
            const unsigned target = (seq + 1u) * gridDim.x;
            while (ld_relaxed_u32(counter) < target) {}
            __threadfence();
        } else {
            red_release_add(counter, 1u);      // publish the CTA's writes at gpu scope (cumulative through the bar.sync)
            const unsigned target = (seq + 1u) * gridDim.x;
            while (ld_relaxed_u32(counter) < target) {}
            fence_acq_rel_gpu();
        }
    }
    seq += 1u;
    consumer_bar_sync();
}

// Barrier over the CTAs that share one arrival counter (variant KO: the nsplit CTAs of a kv group), same protocol as
// grid_barrier: release reduction, relaxed poll, acquire fence; consumer threads only.
__device__ __forceinline__ void group_barrier(unsigned* counter, unsigned target) {
    consumer_bar_sync();
    if (threadIdx.x == 0) {
        red_release_add(counter, 1u);
        while (ld_relaxed_u32(counter) < target) {}
        fence_acq_rel_gpu();
    }
    consumer_bar_sync();
}

// ------------------------------------------------------------------------------------------------ schedule
// Row slab of a [N,K] fp16 matrix owned by this CTA: contiguous rows [r0, r1); `unit` = 2 keeps SwiGLU pairs together.
__device__ __forceinline__ void cta_rows(int N, int unit, int& r0, int& r1) {
    const long long units = N / unit;
    r0 = (int)((units * blockIdx.x) / gridDim.x) * unit;
    r1 = (int)((units * (blockIdx.x + 1)) / gridDim.x) * unit;
}
// Rows per ring stage.  Short phases (few rows per CTA) use smaller stages so that the slab spreads over more consumer
// warps -- the per-stage latency of one warp is on the critical path of every phase.
__host__ __device__ __forceinline__ int rows_per_stage(int K, int N, int grid) {
    int r = kFusedStageBytes / (2 * K);
    r = r > kFusedMaxRows ? kFusedMaxRows : (r < 1 ? 1 : r);
    if (r > 1) r &= ~1;            // even, so interleaved gate/up pairs never straddle a stage
    const int per_cta = N / grid;
    while (r > 2 && per_cta < r * 5) r >>= 1;
    return r;
}
// variant KO: rows of o_proj computed by split `split` of a kv group, and rows per stage for its K' = G*hd columns
__device__ __forceinline__ void ko_rows(int H, int nsplit, int split, int& r0, int& r1) {
    r0 = (int)(((long long)H * split) / nsplit);
    r1 = (int)(((long long)H * (split + 1)) / nsplit);
}
__device__ __forceinline__ int ko_rows_per_stage(int Kp) {
    const int r = kFusedStageBytes / (2 * Kp);
    return r > kFusedMaxRows ? kFusedMaxRows : (r < 1 ? 1 : r);
}
// attention work item of this CTA: (kv head, [hp0, hp1) half pages); empty when the CTA has no item
__device__ __forceinline__ bool attn_item(const FusedArgs& a, int ctx, int& kvh, int& split, int& hp0, int& hp1) {
    const int items = a.nkv * a.nsplit;
    if ((int)blockIdx.x >= items) { kvh = 0; split = 0; hp0 = hp1 = 0; return false; }
    kvh = blockIdx.x / a.nsplit;
    split = blockIdx.x % a.nsplit;
    const int nhp = (ctx + kHalfPage - 1) / kHalfPage;
    const int per = (nhp + a.nsplit - 1) / a.nsplit;
    hp0 = split * per;
    hp1 = min(nhp, hp0 + per);
    if (hp1 < hp0) hp1 = hp0;
    return true;
}

struct Ring {
    uint8_t* buf;
    uint64_t* full;
    uint64_t* empty;
};

// ------------------------------------------------------------------------------------------------ producer
struct Producer {
    Ring ring;
    unsigned it = 0;
    unsigned ns = kFusedStages;
    bool use_hint = false, hint_kv = false;
    uint64_t policy = 0;
    __device__ __forceinline__ void acquire(int& slot) {
        slot = it % ns;
        mbar_wait(&ring.empty[slot], ((it / ns) & 1u) ^ 1u);
    }
    __device__ void rows(const __half* W, int N, int K, int unit) {
        int r0, r1;
        cta_rows(N, unit, r0, r1);
        const int R = rows_per_stage(K, N, gridDim.x);
        for (int r = r0; r < r1; r += R) {
            const int nr = min(R, r1 - r);
            int slot;
            acquire(slot);
            const uint32_t bytes = (uint32_t)nr * K * 2u;
            mbar_expect_tx(&ring.full[slot], bytes);
            if (use_hint) tma_bulk_g2s_hint(ring.buf + (size_t)slot * kFusedStageBytes, W + (size_t)r * K, bytes, &ring.full[slot], policy);
            else tma_bulk_g2s(ring.buf + (size_t)slot * kFusedStageBytes, W + (size_t)r * K, bytes, &ring.full[slot]);
            ++it;
        }
    }
    // K-split down projection: k-rows [r0/2, r1/2) of Wdown^T [I][H] -- the rows that match this CTA's slice of h.
    __device__ void rows_t(const __half* Wt, int I2, int H) {
        int r0, r1;
        cta_rows(I2, 2, r0, r1);
        const int k0 = r0 >> 1, k1 = r1 >> 1;
        const int R = max(1, kFusedStageBytes / (2 * H));
        for (int k = k0; k < k1; k += R) {
            const int nr = min(R, k1 - k);
            int slot;
            acquire(slot);
            const uint32_t bytes = (uint32_t)nr * H * 2u;
            mbar_expect_tx(&ring.full[slot], bytes);
            if (use_hint) tma_bulk_g2s_hint(ring.buf + (size_t)slot * kFusedStageBytes, Wt + (size_t)k * H, bytes, &ring.full[slot], policy);
            else tma_bulk_g2s(ring.buf + (size_t)slot * kFusedStageBytes, Wt + (size_t)k * H, bytes, &ring.full[slot]);
            ++it;
        }
    }
    // variant KO: this CTA's row slice of its kv group's column block of Wo
    __device__ void rows_o_group(const FusedArgs& a, const __half* o_g, int ctx) {
        int kvh, split, hp0, hp1;
        if (!attn_item(a, ctx, kvh, split, hp0, hp1)) return;
        const int Kp = (a.nh / a.nkv) * a.hd;
        int r0, r1;
        ko_rows(a.H, a.nsplit, split, r0, r1);
        const int R = ko_rows_per_stage(Kp);
        const __half* W = o_g + (size_t)kvh * a.H * Kp;
        for (int r = r0; r < r1; r += R) {
            const int nr = min(R, r1 - r);
            int slot;
            acquire(slot);
            const uint32_t bytes = (uint32_t)nr * Kp * 2u;
            mbar_expect_tx(&ring.full[slot], bytes);
            if (use_hint) tma_bulk_g2s_hint(ring.buf + (size_t)slot * kFusedStageBytes, W + (size_t)r * Kp, bytes, &ring.full[slot], policy);
            else tma_bulk_g2s(ring.buf + (size_t)slot * kFusedStageBytes, W + (size_t)r * Kp, bytes, &ring.full[slot]);
            ++it;
        }
    }
    const int* pages = nullptr;   // page table staged in shared memory
    __device__ void attn(const FusedArgs& a, int layer, int ctx) {
        int kvh, split, hp0, hp1;
        if (!attn_item(a, ctx, kvh, split, hp0, hp1)) return;
        const float* kbase = a.kv_pool + (size_t)layer * a.layer_stride;
        const size_t vofs = (size_t)a.nkv * kPage * a.hd;
        for (int hp = hp0; hp < hp1; ++hp) {
            int slot;
            acquire(slot);
            const int page = pages[(hp * kHalfPage) >> kPageShift];
            const size_t off = (size_t)page * a.page_stride + (size_t)kvh * kPage * a.hd + (size_t)((hp * kHalfPage) & (kPage - 1)) * a.hd;
            const uint32_t half_bytes = kHalfPage * a.hd * 4u;  // 8 KB for hd = 128
            uint8_t* dst = ring.buf + (size_t)slot * kFusedStageBytes;
            mbar_expect_tx(&ring.full[slot], 2u * half_bytes);
            if (hint_kv) {
                tma_bulk_g2s_hint(dst, kbase + off, half_bytes, &ring.full[slot], policy);
                tma_bulk_g2s_hint(dst + half_bytes, kbase + vofs + off, half_bytes, &ring.full[slot], policy);
            } else {
                tma_bulk_g2s(dst, kbase + off, half_bytes, &ring.full[slot]);
                tma_bulk_g2s(dst + half_bytes, kbase + vofs + off, half_bytes, &ring.full[slot]);
            }
            ++it;
        }
    }
};

// ------------------------------------------------------------------------------------------------ consumer
// Ring slots are OWNED by consumer warps: slot s is only ever consumed by warp s % 8, so each warp sees the uses
// of its slots strictly in order (no cross-warp mbarrier phase hazards) and 8 stages are processed concurrently.
// A GEMV stage (R rows x K fp16) is reduced entirely by its owner warp against the activation vector staged in
// shared memory; the epilogue is applied by the lanes of that warp -- no cross-warp partial sums.
enum FusedEpi { FE_QKV = 0, FE_RESID = 1, FE_SWIGLU = 2, FE_LOGITS = 3 };

struct Consumer {
    Ring ring;
    unsigned it = 0;    // global stage counter (same sequence as the producer's)
    float* xs;          // [K] activation vector (shared), also aliased by the attention scratch
    float* red;         // [32] scratch
    float* xown;        // [kFusedMaxOwnRows] this CTA's rows of the residual stream (never re-read from global memory)
    int warp, lane;

    unsigned ns = kFusedStages;
    __device__ __forceinline__ bool owns(unsigned i) const { return (int)((i % ns) % kFusedConsumers) == warp; }
    __device__ __forceinline__ const uint8_t* wait_full(unsigned i) {
        const int slot = i % ns;
        mbar_wait(&ring.full[slot], (i / ns) & 1u);
        return ring.buf + (size_t)slot * kFusedStageBytes;
    }
    __device__ __forceinline__ void release(unsigned i) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.empty[i % ns]);
    }

    // Shared-memory layout of the activation vector: element e = 8c + j sits at 4c + j (j < 4) or K/2 + 4c + (j - 4), so
    // the two float4 halves of chunk c are read by lane c at a 16-byte lane stride -- conflict-free LDS.128 (a plain
    // [K] layout puts the lanes 32 bytes apart and every x read pays a 2-way bank conflict).
    static __device__ __forceinline__ int xs_pos(int e, int K) { return ((e >> 2) & 1) * (K >> 1) + (e >> 3) * 4; }
    // Stage the activation vector of a GEMV in shared memory (all consumer threads), optional RMSNorm.
    // src is fp32 global written by other CTAs (ld.global.cg) or an fp16 embedding row.
    // ADD2 (K-split variant): x = src32 + src32b (the residual stream plus the accumulated down projection); the rows
    // [own_r0, own_r1) this CTA owns are captured into xown on the way.
    template <bool ADD2 = false>
    __device__ void load_x(int K, const float* src32, const __half* src16, const float* norm_w, float eps, const float* src32b = nullptr,
                           int own_r0 = 0, int own_r1 = 0) {
        const int tid = threadIdx.x;
        constexpr int kPer = (kFusedMaxK + kFusedConsumers * 32 * 4 - 1) / (kFusedConsumers * 32 * 4);   // float4 per thread
        float4 v[kPer], w[kPer];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int e = (tid + j * kFusedConsumers * 32) * 4;
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            w[j] = v[j];
            if (e < K) {
                if (norm_w) w[j] = *reinterpret_cast<const float4*>(norm_w + e);   // static data: in flight with x
                if (src16) {
                    const uint2 u = *reinterpret_cast<const uint2*>(src16 + e);
                    const float2 a = h2_to_f2(u.x), b = h2_to_f2(u.y);
                    v[j] = make_float4(a.x, a.y, b.x, b.y);
                } else {
                    v[j] = __ldcg(reinterpret_cast<const float4*>(src32 + e));
                    if (ADD2) {
                        const float4 b = __ldcg(reinterpret_cast<const float4*>(src32b + e));
                        v[j].x += b.x; v[j].y += b.y; v[j].z += b.z; v[j].w += b.w;
                        if (e + 3 >= own_r0 && e < own_r1) {
                            const float t[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) if (e + q >= own_r0 && e + q < own_r1) xown[e + q - own_r0] = t[q];
                        }
                    }
                }
            }
        }
        if (norm_w) {
#pragma unroll
            for (int j = 0; j < kPer; ++j) ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
            ss = warp_sum(ss);
            if (lane == 0) red[warp] = ss;
            consumer_bar_sync();
            float tot = 0.f;
#pragma unroll
            for (int q = 0; q < kFusedConsumers; ++q) tot += red[q];
            const float inv = 1.0f / sqrtf(tot / (float)K + eps);
#pragma unroll
            for (int j = 0; j < kPer; ++j) { v[j].x *= inv * w[j].x; v[j].y *= inv * w[j].y; v[j].z *= inv * w[j].z; v[j].w *= inv * w[j].w; }
        }
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int e = (tid + j * kFusedConsumers * 32) * 4;
            if (e < K) *reinterpret_cast<float4*>(xs + xs_pos(e, K)) = v[j];
        }
        consumer_bar_sync();
    }

    // Input of o_proj: merge the split-KV attention partials of every head straight into xs -- each CTA does the (tiny)
    // merge redundantly from L2 instead of one "last" CTA per kv head publishing it behind a fence + atomic + two more
    // dependent L2 round trips.  One warp per head; all loads of a pass are in flight at once.
    __device__ void load_attn(const FusedArgs& a) {
        constexpr int HD = 128, PS = kFusedPartialStride, CH = kFusedMergeChunk;
        const int K = a.nh * HD;
        for (int h0 = warp; h0 < a.nh; h0 += kFusedConsumers) {
            const int h = (h0 + (int)blockIdx.x) % a.nh;   // all CTAs read the same lines: rotate the head order so they spread over the L2 slices
            const float* pb = a.partial + (size_t)h * a.nsplit * PS;
            float M = -INFINITY, L = 0.f;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s0 = 0; s0 < a.nsplit; s0 += CH) {
                float4 v[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j)
                    v[j] = (s0 + j < a.nsplit) ? __ldcg(reinterpret_cast<const float4*>(pb + (size_t)(s0 + j) * PS) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                float m = -INFINITY, l = 0.f;
                if (lane < CH && s0 + lane < a.nsplit) {
                    const float2 ml = __ldcg(reinterpret_cast<const float2*>(pb + (size_t)(s0 + lane) * PS + HD));
                    m = ml.x; l = ml.y;
                }
                float cm = m;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
                const float Mn = fmaxf(M, cm);
                const float so = (M == -INFINITY) ? 0.f : expf(M - Mn);
                const float e = (m == -INFINITY) ? 0.f : expf(m - Mn);
                L = L * so + warp_sum(l * e);
                acc.x *= so; acc.y *= so; acc.z *= so; acc.w *= so;
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const float ej = __shfl_sync(0xffffffffu, e, j);
                    acc.x += v[j].x * ej; acc.y += v[j].y * ej; acc.z += v[j].z * ej; acc.w += v[j].w * ej;
                }
                M = Mn;
            }
            *reinterpret_cast<float4*>(xs + xs_pos(h * HD + 4 * lane, K)) = make_float4(acc.x / L, acc.y / L, acc.z / L, acc.w / L);
        }
        consumer_bar_sync();
    }

    // dot products of a FULL stage of R rows against xs; NA independent accumulators per row break the FMA chain
    template <int R, int NA>
    __device__ __forceinline__ void stage_dots(const uint8_t* st, int K, float (&v)[kFusedMaxRows]) const {
        float acc[R][NA];
#pragma unroll
        for (int q = 0; q < R; ++q)
#pragma unroll
            for (int n = 0; n < NA; ++n) acc[q][n] = 0.f;
        const int nchunk = K >> 3;
        int c = lane;
        for (; c + 32 * (NA - 1) < nchunk; c += 32 * NA) {
#pragma unroll
            for (int n = 0; n < NA; ++n) {
                const int cc = c + 32 * n;
                const float4 x0 = *reinterpret_cast<const float4*>(xs + cc * 4);
                const float4 x1 = *reinterpret_cast<const float4*>(xs + (K >> 1) + cc * 4);
#pragma unroll
                for (int q = 0; q < R; ++q)
                    acc[q][n] = dot8(*reinterpret_cast<const uint4*>(st + (size_t)cc * 16 + (size_t)q * K * 2), x0, x1, acc[q][n]);
            }
        }
        for (; c < nchunk; c += 32) {
            const float4 x0 = *reinterpret_cast<const float4*>(xs + c * 4);
            const float4 x1 = *reinterpret_cast<const float4*>(xs + (K >> 1) + c * 4);
#pragma unroll
            for (int q = 0; q < R; ++q) acc[q][0] = dot8(*reinterpret_cast<const uint4*>(st + (size_t)c * 16 + (size_t)q * K * 2), x0, x1, acc[q][0]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) {
            float t = 0.f;
#pragma unroll
            for (int n = 0; n < NA; ++n) t += acc[q][n];
            v[q] = t;
        }
    }

    // variant KO, input of the K-split o_proj: merge the split partials of the G heads of THIS CTA's kv group into
    // xs[0 .. G*hd) (same arithmetic as load_attn, a G-th of the heads).
    __device__ void load_attn_group(const FusedArgs& a, int kvh, int G) {
        constexpr int HD = 128, PS = kFusedPartialStride, CH = kFusedMergeChunk;
        const int Kp = G * HD;
        if (warp < G) {
            const float* pb = a.partial + (size_t)(kvh * G + warp) * a.nsplit * PS;
            float M = -INFINITY, L = 0.f;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s0 = 0; s0 < a.nsplit; s0 += CH) {
                float4 v[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j)
                    v[j] = (s0 + j < a.nsplit) ? __ldcg(reinterpret_cast<const float4*>(pb + (size_t)(s0 + j) * PS) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                float m = -INFINITY, l = 0.f;
                if (lane < CH && s0 + lane < a.nsplit) {
                    const float2 ml = __ldcg(reinterpret_cast<const float2*>(pb + (size_t)(s0 + lane) * PS + HD));
                    m = ml.x; l = ml.y;
                }
                float cm = m;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
                const float Mn = fmaxf(M, cm);
                const float so = (M == -INFINITY) ? 0.f : expf(M - Mn);
                const float e = (m == -INFINITY) ? 0.f : expf(m - Mn);
                L = L * so + warp_sum(l * e);
                acc.x *= so; acc.y *= so; acc.z *= so; acc.w *= so;
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const float ej = __shfl_sync(0xffffffffu, e, j);
                    acc.x += v[j].x * ej; acc.y += v[j].y * ej; acc.z += v[j].z * ej; acc.w += v[j].w * ej;
                }
                M = Mn;
            }
            *reinterpret_cast<float4*>(xs + xs_pos(warp * HD + 4 * lane, Kp)) = make_float4(acc.x / L, acc.y / L, acc.z / L, acc.w / L);
        }
        consumer_bar_sync();
    }

    // variant KO: xo[r] += Wo_g[kvh][r, :] . xs for this CTA's rows of the group's column block (fp32 reductions into the
    // residual stream; the 8 kv groups add into the same element).
    __device__ void o_ksplit(const FusedArgs& a, int split, int G, float* xo) {
        const int Kp = G * a.hd;
        int r0, r1;
        ko_rows(a.H, a.nsplit, split, r0, r1);
        const int R = ko_rows_per_stage(Kp);
        unsigned i = it;
        for (int r = r0; r < r1; r += R, ++i) {
            if (!owns(i)) continue;
            const int nr = min(R, r1 - r);
            const uint8_t* st = wait_full(i);
            float v[kFusedMaxRows];
#pragma unroll
            for (int q = 0; q < kFusedMaxRows; ++q) v[q] = 0.f;
            if (nr == R && R == 8) stage_dots<8, 1>(st, Kp, v);
            else if (nr == R && R == 4) stage_dots<4, 2>(st, Kp, v);
            else {
                for (int q = 0; q < nr; ++q) {
                    float t[kFusedMaxRows];
                    stage_dots<1, 4>(st + (size_t)q * Kp * 2, Kp, t);
#pragma unroll
                    for (int z = 0; z < kFusedMaxRows; ++z) if (z == q) v[z] = t[0];
                }
            }
            release(i);
            float mine = 0.f;
#pragma unroll
            for (int q = 0; q < kFusedMaxRows; ++q) {
                if (q < nr) {
                    const float t = warp_sum(v[q]);
                    if (lane == q) mine = t;
                }
            }
            if (lane < nr) atomicAdd(xo + r + lane, mine);   // result unused: RED.E.ADD.F32
        }
        it = i;
        consumer_bar_sync();
    }

    // K-split down projection (variant KS): acc[n] += sum_{k in this CTA's slice} Wdown^T[k][n] * h[k].  Every consumer
    // warp reads every stage (thread t owns the 8-column output chunks t, t + 352, ...); the warp that finishes a stage
    // last hands the slot back to the producer.  NCH = output chunks per thread.
    template <int NCH>
    __device__ void down_ksplit(const FusedArgs& a, const float* hs, int* slot_done, float* acc) {
        int r0, r1;
        cta_rows(2 * a.I, 2, r0, r1);
        const int k0 = r0 >> 1, k1 = r1 >> 1;
        const int H = a.H, R = max(1, kFusedStageBytes / (2 * H));
        const int nchunk = H >> 3;
        float av[NCH][8];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) av[j][e] = 0.f;
        unsigned i = it;
        for (int k = k0; k < k1; k += R, ++i) {
            const int nr = min(R, k1 - k);
            const uint8_t* st = wait_full(i);
            for (int q = 0; q < nr; ++q) {
                const float hv = hs[k - k0 + q];
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int ch = (int)threadIdx.x + j * kFusedConsumers * 32;
                    if (ch < nchunk) {
                        const uint4 w = *reinterpret_cast<const uint4*>(st + (size_t)q * H * 2 + (size_t)ch * 16);
                        const float2 w0 = h2_to_f2(w.x), w1 = h2_to_f2(w.y), w2 = h2_to_f2(w.z), w3 = h2_to_f2(w.w);
                        av[j][0] = fmaf(w0.x, hv, av[j][0]); av[j][1] = fmaf(w0.y, hv, av[j][1]);
                        av[j][2] = fmaf(w1.x, hv, av[j][2]); av[j][3] = fmaf(w1.y, hv, av[j][3]);
                        av[j][4] = fmaf(w2.x, hv, av[j][4]); av[j][5] = fmaf(w2.y, hv, av[j][5]);
                        av[j][6] = fmaf(w3.x, hv, av[j][6]); av[j][7] = fmaf(w3.y, hv, av[j][7]);
                    }
                }
            }
            __syncwarp();   // every lane's shared-memory reads of the stage have fed its FMAs
            if (lane == 0) {
                const int slot = i % ns;
                if (atomicAdd(&slot_done[slot], 1) == kFusedConsumers - 1) {
                    slot_done[slot] = 0;   // ordered before the slot's next use by the mbarrier release / acquire chain
                    mbar_arrive(&ring.empty[slot]);
                }
            }
        }
        it = i;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int ch = (int)threadIdx.x + j * kFusedConsumers * 32;
            if (ch < nchunk) {
#pragma unroll
                for (int e = 0; e < 8; ++e) atomicAdd(acc + ch * 8 + e, av[j][e]);   // result unused: RED.E.ADD.F32
            }
        }
        consumer_bar_sync();
    }

    // One GEMV phase: y[r0:r1) = W[r0:r1, :] . xs  with the fused epilogue.  Ends with a consumer barrier so
    // that xs may be overwritten by the next phase.
    template <int EPI>
    __device__ void gemv(const FusedArgs& a, int N, int K, const float* bias, float* out, float& best, int& bi) {
        int r0, r1;
        cta_rows(N, EPI == FE_SWIGLU ? 2 : 1, r0, r1);
        const int R = rows_per_stage(K, N, gridDim.x);
        unsigned i = it;
        for (int r = r0; r < r1; r += R, ++i) {
            if (!owns(i)) continue;
            const int nr = min(R, r1 - r);
            const uint8_t* st = wait_full(i);
            float v[kFusedMaxRows];
#pragma unroll
            for (int q = 0; q < kFusedMaxRows; ++q) v[q] = 0.f;
            if (!(a.dbg & 2)) {
                if (nr == R) {
                    switch (R) {
                        case 8: stage_dots<8, 1>(st, K, v); break;
                        case 4: stage_dots<4, 2>(st, K, v); break;
                        case 2: stage_dots<2, 2>(st, K, v); break;
                        default: stage_dots<1, 4>(st, K, v); break;
                    }
                } else {   // tail stage of the slab: row by row
                    for (int q = 0; q < nr; ++q) {
                        float t[kFusedMaxRows];
                        stage_dots<1, 4>(st + (size_t)q * K * 2, K, t);
#pragma unroll
                        for (int z = 0; z < kFusedMaxRows; ++z) if (z == q) v[z] = t[0];
                    }
                }
            }
            release(i);   // every shared-memory read of the stage fed v[] above; __syncwarp orders the lanes
            float mine = 0.f, mate = 0.f;   // lane q keeps row q (and row q^1 for the SwiGLU pair)
#pragma unroll
            for (int q = 0; q < kFusedMaxRows; ++q) {
                if (q < nr) {
                    const float t = warp_sum(v[q]);
                    if (lane == q) mine = t;
                    if (lane == (q ^ 1)) mate = t;
                }
            }
            if (lane < nr) {
                const int row = r + lane;
                if (EPI == FE_SWIGLU) {
                    if ((lane & 1) == 0) out[row >> 1] = silu_f(mine) * mate;   // rows (2i, 2i+1) = (gate_i, up_i)
                } else {
                    float y = mine;
                    if (bias) y += bias[row];
                    if (EPI == FE_RESID) { y += xown[row - r0]; xown[row - r0] = y; }   // o_proj and down own the same rows of x
                    out[row] = y;
                    if (EPI == FE_LOGITS && (y > best || (y == best && row < bi))) { best = y; bi = row; }
                }
            }
        }
        it = i;
        consumer_bar_sync();
    }
};

// Attention scratch in shared memory (consumer side; aliases the xs region, unused during this phase)
template <int G>
struct AttnSmem {
    float qs[G][128];
    float knew[128];
    float vnew[128];
    float m[kFusedConsumers][G];
    float l[kFusedConsumers][G];
    float acc[kFusedConsumers][G][128];

};

template <int G>
__device__ void fused_attention(const FusedArgs& a, Consumer& c, AttnSmem<G>& s, const float* cs, const int* spages, int layer, const FusedLayer& Ly, int t_new) {
    constexpr int HD = 128;
    const int ctx = t_new + 1;
    const int lane = c.lane, warp = c.warp, tid = threadIdx.x;
    int kvh, split, hp0, hp1;
    const bool has_item = attn_item(a, ctx, kvh, split, hp0, hp1);
    if (!has_item) return;  // this CTA's producer issued nothing for the phase either
    // ---- prologue: q heads (warps 0..G-1) and the new k (warp G): RMSNorm over hd then RoPE; v copy (warp G+1)
    if (warp <= G) {
        const bool is_q = warp < G;
        const float* src = a.qkv1 + (size_t)(is_q ? (kvh * G + warp) : (a.nh + kvh)) * HD;
        const float4 x = __ldcg(reinterpret_cast<const float4*>(src + lane * 4));
        float ss = x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        ss = warp_sum(ss);
        const float inv = 1.0f / sqrtf(ss / (float)HD + a.eps);
        const float4 w = *reinterpret_cast<const float4*>((is_q ? Ly.qn : Ly.kn) + lane * 4);
        const float n[4] = {x.x * inv * w.x, x.y * inv * w.y, x.z * inv * w.z, x.w * inv * w.w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float partner = __shfl_xor_sync(0xffffffffu, n[e], 16);
            const int d = lane * 4 + e, j = d & (HD / 2 - 1);
            const float rot = (d < HD / 2) ? -partner : partner;
            o[e] = n[e] * cs[j] + rot * cs[HD / 2 + j];
        }
        float* dst = is_q ? s.qs[warp] : s.knew;
        *reinterpret_cast<float4*>(dst + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
    } else if (warp == G + 1) {
        *reinterpret_cast<float4*>(s.vnew + lane * 4) = __ldcg(reinterpret_cast<const float4*>(a.qkv1 + (size_t)(a.nh + a.nkv + kvh) * HD + lane * 4));
    }
    consumer_bar_sync();
    const int hp_new = t_new / kHalfPage;
    if (hp_new >= hp0 && hp_new < hp1 && tid < HD) {  // append K,V of the current token to the paged cache
        const int page = spages[t_new >> kPageShift];
        const size_t off = (size_t)layer * a.layer_stride + (size_t)page * a.page_stride + (size_t)kvh * kPage * HD + (size_t)(t_new & (kPage - 1)) * HD + tid;
        a.kv_pool[off] = s.knew[tid];
        a.kv_pool[off + (size_t)a.nkv * kPage * HD] = s.vnew[tid];
    }
    // Lane layout inside the owner warp: group = lane / 8 handles token (4*itr + group) of the half page,
    // sub = lane % 8 handles dims {4*sub + 32*e + 0..3 : e = 0..3} (conflict-free 128-byte rows per quarter warp).
    const int grp = lane >> 3, sub = lane & 7;
    float4 q[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e) q[g][e] = *reinterpret_cast<const float4*>(s.qs[g] + 4 * sub + 32 * e);
    float m[G], l[G];
    float4 acc[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[g][e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    unsigned i = c.it;
    for (int hp = hp0; hp < hp1; ++hp, ++i) {
        if (!c.owns(i)) continue;
        const uint8_t* st = c.wait_full(i);
        const float* ks = reinterpret_cast<const float*>(st);
        const float* vs = reinterpret_cast<const float*>(st + kHalfPage * HD * 4);
        const int tbase = hp * kHalfPage;
#pragma unroll
        for (int itr = 0; itr < kHalfPage / 4; ++itr) {
            const int tl = 4 * itr + grp;          // token inside the half page
            const int t = tbase + tl;
            const bool valid = t < ctx;
            const float* kr = (t == t_new) ? s.knew : ks + tl * HD;
            const float* vr = (t == t_new) ? s.vnew : vs + tl * HD;
            float4 kk[4], vv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                kk[e] = *reinterpret_cast<const float4*>(kr + 4 * sub + 32 * e);
                vv[e] = *reinterpret_cast<const float4*>(vr + 4 * sub + 32 * e);
                if (!valid) vv[e] = make_float4(0.f, 0.f, 0.f, 0.f);   // slots past ctx hold stale data: 0 * x must stay 0
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float sc = 0.f;
#pragma unroll
                for (int e = 0; e < 4; ++e) sc += q[g][e].x * kk[e].x + q[g][e].y * kk[e].y + q[g][e].z * kk[e].z + q[g][e].w * kk[e].w;
                sc += __shfl_xor_sync(0xffffffffu, sc, 1);
                sc += __shfl_xor_sync(0xffffffffu, sc, 2);
                sc += __shfl_xor_sync(0xffffffffu, sc, 4);
                sc = valid ? sc * a.scaling : -INFINITY;
                const float mnew = fmaxf(m[g], sc);
                const float muse = (mnew == -INFINITY) ? 0.f : mnew;
                const float alpha = expf(m[g] - muse);
                const float pp = expf(sc - muse);
                l[g] = l[g] * alpha + pp;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc[g][e].x = acc[g][e].x * alpha + pp * vv[e].x;
                    acc[g][e].y = acc[g][e].y * alpha + pp * vv[e].y;
                    acc[g][e].z = acc[g][e].z * alpha + pp * vv[e].z;
                    acc[g][e].w = acc[g][e].w * alpha + pp * vv[e].w;
                }
                m[g] = mnew;
            }
        }
        c.release(i);  // after the math: every shared-memory read of this stage has been consumed
    }
    c.it = i;
    // merge the 4 token groups of the warp (lanes differing in bits 3,4), then publish the per-warp state
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float M = fmaxf(m[g], __shfl_xor_sync(0xffffffffu, m[g], 8));
        M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 16));
        const float sc = (m[g] == -INFINITY) ? 0.f : expf(m[g] - M);
        float L = l[g] * sc;
        L += __shfl_xor_sync(0xffffffffu, L, 8);
        L += __shfl_xor_sync(0xffffffffu, L, 16);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float4 t = make_float4(acc[g][e].x * sc, acc[g][e].y * sc, acc[g][e].z * sc, acc[g][e].w * sc);
            t.x += __shfl_xor_sync(0xffffffffu, t.x, 8); t.x += __shfl_xor_sync(0xffffffffu, t.x, 16);
            t.y += __shfl_xor_sync(0xffffffffu, t.y, 8); t.y += __shfl_xor_sync(0xffffffffu, t.y, 16);
            t.z += __shfl_xor_sync(0xffffffffu, t.z, 8); t.z += __shfl_xor_sync(0xffffffffu, t.z, 16);
            t.w += __shfl_xor_sync(0xffffffffu, t.w, 8); t.w += __shfl_xor_sync(0xffffffffu, t.w, 16);
            if (grp == 0) *reinterpret_cast<float4*>(&s.acc[warp][g][4 * sub + 32 * e]) = t;
        }
        if (lane == 0) { s.m[warp][g] = M; s.l[warp][g] = L; }
    }
    consumer_bar_sync();
    for (int idx = tid; idx < G * HD; idx += kFusedConsumers * 32) {
        const int g = idx / HD, d = idx % HD;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < kFusedConsumers; ++w) M = fmaxf(M, s.m[w][g]);
        float L = 0.f, O = 0.f;
        if (M != -INFINITY) {
#pragma unroll
            for (int w = 0; w < kFusedConsumers; ++w) {
                const float e = expf(s.m[w][g] - M);
                L += s.l[w][g] * e;
                O += s.acc[w][g][d] * e;
            }
        }
        float* p = a.partial + ((size_t)(kvh * G + g) * a.nsplit + split) * kFusedPartialStride;
        p[d] = O;
        if (d == 0) { p[HD] = M; p[HD + 1] = L; }
    }
    // the split partials are merged by the consumers of the next phase (Consumer::load_attn), after the grid barrier
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int G, bool KS = false, bool KO = false>
__global__ void __launch_bounds__(kFusedThreads, 1) decode_step_fused_kernel(FusedArgs a) {
    extern __shared__ __align__(1024) uint8_t fused_smem_raw[];
    uint8_t* ringbuf = fused_smem_raw;
    uint64_t* full = reinterpret_cast<uint64_t*>(fused_smem_raw + (size_t)kFusedStages * kFusedStageBytes);
    uint64_t* empty = full + kFusedStages;
    float* red = reinterpret_cast<float*>(empty + kFusedStages);
    float* xown = red + 32;                                     // [kFusedMaxOwnRows]
    float* cs = xown + kFusedMaxOwnRows;                        // [128] cos | sin of the step's rotary angles
    float* xs = cs + 128;                                       // [kFusedMaxK] activations / attention scratch
    int* spages = reinterpret_cast<int*>(xs + kFusedMaxK);      // [kFusedMaxPages] page table copy
    float* hs = reinterpret_cast<float*>(spages + kFusedMaxPages);   // [kFusedMaxHs] this CTA's slice of silu(gate)*up (variant KS)
    int* slot_done = reinterpret_cast<int*>(hs + kFusedMaxHs);       // [kFusedStages] warps done with a shared stage (variant KS)
    AttnSmem<G>* as = reinterpret_cast<AttnSmem<G>*>(xs);
    static_assert(sizeof(AttnSmem<G>) <= kFusedMaxK * sizeof(float), "attention scratch must fit in the xs region");

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < kFusedStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (KS && tid < kFusedStages) slot_done[tid] = 0;
    if ((a.dbg & 64) && tid == 0) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); a.trace[8192 + blockIdx.x * 256 + 255] = smid; }
    const int t_new = a.st->pos;
    const int rope_delta = a.st->rope_delta;
    const uint32_t token = a.st->token;
    const int ctx = t_new + 1;
    for (int i = tid; i < (ctx + kPage - 1) / kPage && i < kFusedMaxPages; i += kFusedThreads) spages[i] = a.page_table[i];
    if (tid < 64) {   // RoPE angle of this step (same for every layer): cos/sin once per kernel
        const float ang = (float)(t_new + rope_delta) * a.inv_freq[tid];
        cs[tid] = cosf(ang);
        cs[64 + tid] = sinf(ang);
    }
    __syncthreads();
    Ring ring{ringbuf, full, empty};

    if (warp == kFusedConsumers) {
        // =================================================== PRODUCER (one elected lane)
        if (lane == 0) {
            Producer p;
            p.ring = ring;
            p.ns = (unsigned)a.stages;
            p.use_hint = (a.dbg & 16) == 0;      // dbg bit4 switches the L2 evict-first hint off (A/B runs)
            p.hint_kv = (a.dbg & 32) == 0;       // bit5: same switch for the KV half-page copies
            p.policy = l2_evict_first_policy();
            p.pages = spages;
            int pe = 0;
            AHA_STAMP(a, 1, pe);
            FusedLayer nxt = a.layers[0];
            for (int l = 0; l < a.L; ++l) {
                const FusedLayer Ly = nxt;
                if (l + 1 < a.L) nxt = a.layers[l + 1];   // pointer table one layer ahead: off the issue path
                p.rows(Ly.qkv, a.qkv_dim, a.H, 1); AHA_STAMP(a, 1, pe);
                p.attn(a, l, ctx); AHA_STAMP(a, 1, pe);
                if (KO) p.rows_o_group(a, Ly.o_g, ctx); else p.rows(Ly.o, a.H, a.nh * a.hd, 1);
                AHA_STAMP(a, 1, pe);
                p.rows(Ly.gu, 2 * a.I, a.H, 2); AHA_STAMP(a, 1, pe);
                if (KS) p.rows_t(Ly.down_t, 2 * a.I, a.H); else p.rows(Ly.down, a.H, a.I, 1);
                AHA_STAMP(a, 1, pe);
            }
            p.rows(a.lm_head, a.V, a.H, 1); AHA_STAMP(a, 1, pe);
        }
        return;
    }
    // ======================================================= CONSUMERS
    Consumer c;
    c.ring = ring; c.ns = (unsigned)a.stages; c.xs = xs; c.red = red; c.xown = xown; c.warp = warp; c.lane = lane;
    unsigned seq = 0;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    const __half* emb_row = a.embed + (size_t)token * a.H;
    int own_r0, own_r1, hs_k0 = 0;
    cta_rows(a.H, 1, own_r0, own_r1);
    if (tid < own_r1 - own_r0) xown[tid] = __half2float(emb_row[own_r0 + tid]);   // the rows of the residual stream this CTA owns start as the embedding row
    if (KS) { int g0, g1; cta_rows(2 * a.I, 2, g0, g1); hs_k0 = g0 >> 1; }
    int ce = 0;
#define CSTAMP() do { if (tid == 0) AHA_STAMP(a, 0, ce); } while (0)
    CSTAMP();
    FusedLayer nxtc = a.layers[0];
    for (int l = 0; l < a.L; ++l) {
        const FusedLayer Ly = nxtc;
        if (l + 1 < a.L) nxtc = a.layers[l + 1];
        const bool first = (l == 0);
        // variant KS: layer l accumulates its down projection into acc2[l & 1]; layer l + 1 adds it while loading x and
        // re-zeroes its own rows of it one barrier later, two barriers before layer l + 2 accumulates into it again
        float* const acc_prev = a.acc2 + (size_t)((l + 1) & 1) * a.H;
        float* const acc_cur = a.acc2 + (size_t)(l & 1) * a.H;
        // variant KO: x after o_proj lives in xo[l & 1] (double-buffered so that the owners' writes for layer l never meet a
        // reader of layer l - 1)
        float* const xo_cur = KO ? a.xo + (size_t)(l & 1) * a.H : nullptr;
        const float* const xo_prev = KO ? a.xo + (size_t)((l + 1) & 1) * a.H : nullptr;
        // P1: qkv = Wqkv . rmsnorm(x)      (layer 0 reads the embedding row directly: Embedding::forward)
        if (KS && !first) c.template load_x<true>(a.H, KO ? xo_prev : a.x, nullptr, Ly.ln1, a.eps, acc_prev, own_r0, own_r1);
        else c.load_x(a.H, a.x, first ? emb_row : nullptr, Ly.ln1, a.eps);
        if (KO && tid < own_r1 - own_r0) xo_cur[own_r0 + tid] = xown[tid];   // x entering the layer; the o_proj partial sums are added to it after the next barrier
        CSTAMP();
        c.template gemv<FE_QKV>(a, a.qkv_dim, a.H, Ly.qkv_b, a.qkv1, best, bi); CSTAMP();
        grid_barrier(&a.sync[0], seq, a.dbg, a.trace); CSTAMP();
        // P2: attention over the paged cache (+ q/k norm, RoPE, KV append)
        CSTAMP();
        fused_attention<G>(a, c, *as, cs, spages, l, Ly, t_new); CSTAMP();
        if (KO) {
            // P3 (variant KO): only the CTAs of this kv group synchronise; each merges its group's heads and adds its rows
            // of Wo[:, group columns] . attn_group into xo
            int kvh, split, hp0, hp1;
            const bool has_item = attn_item(a, ctx, kvh, split, hp0, hp1);
            if (has_item) {
                group_barrier(&a.gsync[kvh], (unsigned)(l + 1) * (unsigned)a.nsplit); CSTAMP();
                c.load_attn_group(a, kvh, G); CSTAMP();
            } else { CSTAMP(); CSTAMP(); }
            if (KS && !first && tid < own_r1 - own_r0) acc_prev[own_r0 + tid] = 0.f;
            if (has_item) c.o_ksplit(a, split, G, xo_cur);
            CSTAMP();
            grid_barrier(&a.sync[0], seq, a.dbg, a.trace); CSTAMP();
        } else {
            grid_barrier(&a.sync[0], seq, a.dbg, a.trace); CSTAMP();
            // P3: x = resid + Wo . attn
            c.load_attn(a); CSTAMP();
            if (KS && !first && tid < own_r1 - own_r0) acc_prev[own_r0 + tid] = 0.f;   // read last in P1 (before its barrier), accumulated next by layer l + 1
            c.template gemv<FE_RESID>(a, a.H, a.nh * a.hd, Ly.o_b, a.x, best, bi); CSTAMP();
            grid_barrier(&a.sync[0], seq, a.dbg, a.trace); CSTAMP();
        }
        // P4: h = silu(gate) * up, gate/up rows interleaved, input rmsnorm(x)
        float own_mid = 0.f;
        if (KO && tid < own_r1 - own_r0) own_mid = __ldcg(xo_cur + own_r0 + tid);   // this CTA's rows of x after o_proj, in flight with the load below
        c.load_x(a.H, KO ? xo_cur : a.x, nullptr, Ly.ln2, a.eps);
        if (KO && tid < own_r1 - own_r0) xown[tid] = own_mid;   // residual of the down projection (read in its epilogue, CTA barriers away)
        CSTAMP();
        if (KS) {
            // h stays in shared memory; P5 follows without a barrier: acc_cur += Wdown[:, slice] . h[slice]
            c.template gemv<FE_SWIGLU>(a, 2 * a.I, a.H, nullptr, hs - hs_k0, best, bi); CSTAMP();
            CSTAMP(); CSTAMP();
            switch ((a.H / 8 + kFusedConsumers * 32 - 1) / (kFusedConsumers * 32)) {
                case 1: c.template down_ksplit<1>(a, hs, slot_done, acc_cur); break;
                case 2: c.template down_ksplit<2>(a, hs, slot_done, acc_cur); break;
                default: c.template down_ksplit<3>(a, hs, slot_done, acc_cur); break;
            }
            CSTAMP();
            grid_barrier(&a.sync[0], seq, a.dbg, a.trace); CSTAMP();
        } else {
            c.template gemv<FE_SWIGLU>(a, 2 * a.I, a.H, nullptr, a.h1, best, bi); CSTAMP();
            grid_barrier(&a.sync[0], seq, a.dbg, a.trace); CSTAMP();
            // P5: x = x + Wdown . h
            c.load_x(a.I, a.h1, nullptr, nullptr, 0.f); CSTAMP();
            c.template gemv<FE_RESID>(a, a.H, a.I, nullptr, a.x, best, bi); CSTAMP();
            grid_barrier(&a.sync[0], seq, a.dbg, a.trace); CSTAMP();
        }
    }
    // final: logits = lm_head . rmsnorm(x), per-CTA argmax candidate
    if (KS) c.template load_x<true>(a.H, KO ? a.xo + (size_t)((a.L - 1) & 1) * a.H : a.x, nullptr, a.final_norm, a.eps, a.acc2 + (size_t)((a.L - 1) & 1) * a.H, own_r0, own_r1);
    else c.load_x(a.H, a.x, nullptr, a.final_norm, a.eps);
    best = -INFINITY; bi = 0x7fffffff;
    c.template gemv<FE_LOGITS>(a, a.V, a.H, nullptr, a.logits, best, bi); CSTAMP();
    // CTA-level argmax (first maximal index), then the last CTA to arrive reduces across CTAs
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { red[warp] = best; reinterpret_cast<int*>(red)[16 + warp] = bi; }
    consumer_bar_sync();
    if (tid == 0) {
        for (int w = 1; w < kFusedConsumers; ++w) {
            const float ov = red[w];
            const int oi = reinterpret_cast<int*>(red)[16 + w];
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        a.pmax[blockIdx.x] = best;
        a.pidx[blockIdx.x] = bi;
        const unsigned ticket = atom_acq_rel_add(&a.sync[1], 1u);
        reinterpret_cast<int*>(red)[31] = (ticket == gridDim.x - 1) ? 1 : 0;
    }
    consumer_bar_sync();
    if (warp == 0 && reinterpret_cast<int*>(red)[31]) {   // last CTA to arrive: warp 0 reduces the per-CTA candidates
        float gb = -INFINITY;
        int gi = 0x7fffffff;
        for (unsigned i = lane; i < gridDim.x; i += 32) {
            const float v = __ldcg(a.pmax + i);
            const int id = __ldcg(a.pidx + i);
            if (v > gb || (v == gb && id < gi)) { gb = v; gi = id; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, gb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, gi, o);
            if (ov > gb || (ov == gb && oi < gi)) { gb = ov; gi = oi; }
        }
        if (lane == 0) {
            *a.argmax_out = (uint32_t)gi;
            DecodeState* st = a.st;
            st->token = (uint32_t)gi;
            st->pos = t_new + 1;
            if (st->n_hist < a.hist_cap) a.history[st->n_hist] = (uint32_t)gi;
            st->n_hist += 1;
        }
    }
}

template <int G>
inline size_t fused_smem_bytes() {
    return (size_t)kFusedStages * kFusedStageBytes + 2 * kFusedStages * sizeof(uint64_t) + (32 + kFusedMaxOwnRows + 128) * sizeof(float) + (size_t)kFusedMaxK * sizeof(float) +
           (size_t)kFusedMaxPages * sizeof(int) + (size_t)kFusedMaxHs * sizeof(float) + (size_t)kFusedStages * sizeof(int) + 64;
}

}  // namespace aha
