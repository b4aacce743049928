// vision_model.cuh -- Qwen3-VL vision tower: patch-embed GEMM, bilinear pos-embed, 2-D RoPE, 24 pre-LN
// blocks with per-segment full attention, deepstack + final patch mergers.
// Reference: /root/reference/src/models/qwen3vl/model.rs:32-104 (patch embed), :106-185 (merger),
// :187-279 (attention), :281-371 (block), :512-639 (fast_pos_embed_interpolate), :641-690 (rot_pos_emb),
// :692-740 (forward); /root/reference/src/position_embed/rope.rs:75-94,424-441.
#pragma once
#include "text_model.cuh"

namespace aha {

struct VisionCfg {
    int depth = 0, H = 0, I = 0, heads = 0, npos = 0, out_hidden = 0, patch = 16, merge = 2, tpatch = 2, in_ch = 3;
    int act = ACT_GELU_TANH;
    std::vector<int> deepstack;
    static int act_from(const std::string& s) {
        if (s == "gelu_pytorch_tanh" || s == "gelu_new") return ACT_GELU_TANH;
        if (s == "gelu") return ACT_GELU_ERF;
        if (s == "silu" || s == "swish") return ACT_SILU;
        throw std::runtime_error("unsupported activation '" + s + "'");
    }
    static VisionCfg from_json(const Json& j) {
        VisionCfg c;
        c.depth = j.integer("depth"); c.H = j.integer("hidden_size"); c.I = j.integer("intermediate_size");
        c.heads = j.integer("num_heads"); c.npos = j.integer("num_position_embeddings"); c.out_hidden = j.integer("out_hidden_size");
        c.patch = j.integer_or("patch_size", 16); c.merge = j.integer_or("spatial_merge_size", 2);
        c.tpatch = j.integer_or("temporal_patch_size", 2); c.in_ch = j.integer_or("in_channels", 3);
        c.act = act_from(j.string_or("hidden_act", "gelu_pytorch_tanh"));
        c.deepstack = j.int_array("deepstack_visual_indexes");
        return c;
    }
};

// Per-image: bilinear-interpolated learned position embedding added to x, and the (row, col) of every patch
// in merge-block order.  One block per patch of the image.
// linspace(0, n-1, h)[i] = 0 + i*step with step = (n-1)/(h-1) in f32 (tensor_utils.rs:354-365); floor by
// f32->u32 truncation; ceil = min(floor+1, n-1); weights (1-dh)(1-dw), (1-dh)dw, dh(1-dw), dh*dw.
__global__ void vit_pos_embed_kernel(float* __restrict__ x, const __half* __restrict__ pos_embed, int2* __restrict__ rowcol,
                                     int patch0, int t, int h, int w, int merge, int n_side, float step_h, float step_w, int H) {
    const int p = blockIdx.x;              // patch index within the image, merge-block order
    const int per_t = h * w;
    const int q = p % per_t;               // position inside one temporal slice
    const int mw = w / merge;
    const int blk = q / (merge * merge), in = q % (merge * merge);
    const int row = (blk / mw) * merge + in / merge;
    const int col = (blk % mw) * merge + in % merge;
    if (threadIdx.x == 0) rowcol[patch0 + p] = make_int2(row, col);
    const float hv = (h == 1) ? 0.f : 0.f + (float)row * step_h;
    const float wv = (w == 1) ? 0.f : 0.f + (float)col * step_w;
    const unsigned hf = (unsigned)hv, wf = (unsigned)wv;
    const unsigned hc = min(hf + 1u, (unsigned)(n_side - 1)), wc = min(wf + 1u, (unsigned)(n_side - 1));
    const float dh = hv - (float)hf, dw = wv - (float)wf;
    const float w0 = (1.f - dh) * (1.f - dw), w1 = (1.f - dh) * dw, w2 = dh * (1.f - dw), w3 = dh * dw;
    const __half* e0 = pos_embed + (size_t)(hf * n_side + wf) * H;
    const __half* e1 = pos_embed + (size_t)(hf * n_side + wc) * H;
    const __half* e2 = pos_embed + (size_t)(hc * n_side + wf) * H;
    const __half* e3 = pos_embed + (size_t)(hc * n_side + wc) * H;
    float* xr = x + (size_t)(patch0 + p) * H;
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        float pe = __half2float(e0[i]) * w0;
        pe = pe + __half2float(e1[i]) * w1;
        pe = pe + __half2float(e2[i]) * w2;
        pe = pe + __half2float(e3[i]) * w3;
        xr[i] += pe;
    }
}

// 2-D RoPE on q and k in place.  qkv: [N, 3*Hv] with Hv = heads * HDP (head slots of HDP floats, the first HD are the head);
// freq j < HD/4 from the row, HD/4 <= j < HD/2 from the col (table p * inv_freq(dim=HD/2, theta=1e4), rope.rs:424-441);
// emb = cat(f, f); x*cos + rotate_half(x)*sin.
__global__ void vit_rope_kernel(float* __restrict__ qkv, const int2* __restrict__ rowcol, const float* __restrict__ inv_freq, int HD, int HDP,
                                int heads, int Hv) {
    const int p = blockIdx.x;
    const int2 rc = rowcol[p];
    for (int idx = threadIdx.x; idx < 2 * heads * (HD / 2); idx += blockDim.x) {
        const int which = idx / (heads * (HD / 2));        // 0 = q, 1 = k
        const int rem = idx % (heads * (HD / 2));
        const int hh = rem / (HD / 2), j = rem % (HD / 2);
        const float ang = (j < HD / 4) ? (float)rc.x * inv_freq[j] : (float)rc.y * inv_freq[j - HD / 4];
        const float c = cosf(ang), s = sinf(ang);
        float* base = qkv + (size_t)p * 3 * Hv + (size_t)which * Hv + (size_t)hh * HDP;
        const float x1 = base[j], x2 = base[j + HD / 2];
        base[j] = x1 * c - x2 * s;
        base[j + HD / 2] = x2 * c + x1 * s;
    }
}

struct VisionBlock {
    float *n1w, *n1b, *n2w, *n2b;
    LinearW qkv, proj, fc1, fc2;
};
struct Merger {
    float *nw, *nb;
    LinearW fc1, fc2;
    bool post;
};

struct VisionModel {
    VisionCfg cfg;
    Ctx* ctx = nullptr;
    LinearW patch;         // [Hv, 1536] + bias
    __half* pos_embed = nullptr;
    int n_side = 0;
    float* inv_freq = nullptr;  // hd / 4 entries
    int hd = 64, hdp = 64;      // head_dim and the width of a head slot in qkv / attn (64, or 128 when head_dim is not 64: zero padded)
    int Hp = 0, Ip = 0;         // heads * hdp; intermediate size rounded up to a multiple of 64 (zero rows / columns)
    std::vector<VisionBlock> blocks;
    Merger merger;
    std::vector<Merger> ds_mergers;
    int max_patches = 0, patch_dim = 0;
    float *pix = nullptr, *x = nullptr, *xn = nullptr, *qkv = nullptr, *attn = nullptr, *h = nullptr, *mtmp = nullptr;
    float* image_embeds = nullptr;
    std::vector<float*> ds_out;
    int2* rowcol = nullptr;
    bool trace = false;
    float* trace_buf = nullptr;
    int last_N = 0;

    static LinearW lin(Ctx& c, const WeightTable& wt, const std::string& p, int N, int K) {
        return upload_linear(c, wt, {{p + ".weight", N, 0, N}}, K, 0, K, false, {p + ".bias"});
    }
    Merger load_merger(Ctx& c, const WeightTable& wt, const std::string& p, bool post) {
        Merger m;
        const int Hm = cfg.H * cfg.merge * cfg.merge;
        m.post = post;
        m.nw = upload_vec(c, wt, p + "norm.weight", post ? Hm : cfg.H);
        m.nb = upload_vec(c, wt, p + "norm.bias", post ? Hm : cfg.H);
        m.fc1 = lin(c, wt, p + "linear_fc1", Hm, Hm);
        m.fc2 = lin(c, wt, p + "linear_fc2", cfg.out_hidden, Hm);
        return m;
    }
    void load(Ctx& c, const VisionCfg& cf, const WeightTable& wt, const std::string& p, int max_patches_) {
        ctx = &c; cfg = cf; max_patches = max_patches_;
        // head_dim 64 (Qwen3-VL-2B / 4B towers) runs as is; any other head_dim <= 128 (72 in the 8B / 32B towers) is laid out in
        // 128-wide head slots whose tail is zero: the qkv rows and the proj columns are padded at upload, so every kernel downstream
        // (RoPE, the tcgen05 attention, the GEMMs) sees a shape it tiles and the result is unchanged.  The same for an
        // intermediate size that is not a multiple of 64 (4304 -> 4352).
        AHA_REQUIRE(cfg.H % cfg.heads == 0, "vision hidden_size must be a multiple of num_heads");
        hd = cfg.H / cfg.heads;
        AHA_REQUIRE(hd % 4 == 0 && hd <= 128, "vision head_dim must be a multiple of 4 and at most 128");
        hdp = hd == 64 ? 64 : 128;
        Hp = cfg.heads * hdp;
        Ip = (cfg.I + 63) / 64 * 64;
        patch_dim = cfg.in_ch * cfg.tpatch * cfg.patch * cfg.patch;
        AHA_REQUIRE(patch_dim % 16 == 0 && cfg.H % 16 == 0, "vision patch and hidden sizes must be multiples of 16");   // the intermediate size is padded to 64
        patch = upload_linear(c, wt, {{p + "patch_embed.proj.weight", cfg.H, 0, cfg.H}}, patch_dim, 0, patch_dim, false, {p + "patch_embed.proj.bias"});
        {
            std::vector<__half> st((size_t)cfg.npos * cfg.H);
            wt.rows_to_half(p + "pos_embed.weight", cfg.npos, cfg.H, 0, cfg.npos, 0, cfg.H, st.data(), cfg.H);
            pos_embed = upload(c, st);
        }
        n_side = (int)sqrtf((float)cfg.npos);  // (num_position_embeddings as f32).sqrt() as u32, model.rs:395
        std::vector<float> inv(hd / 4);
        for (int j = 0; j < hd / 4; ++j) inv[j] = 1.0f / powf(10000.0f, (float)(2 * j) / (float)(hd / 2));
        inv_freq = upload(c, inv);
        std::vector<int> id_H(cfg.H), id_I(cfg.I), slot_rows(3 * (size_t)Hp, -1), slot_cols((size_t)Hp, -1), pad_I((size_t)Ip, -1);
        for (int i = 0; i < cfg.H; ++i) id_H[i] = i;
        for (int i = 0; i < cfg.I; ++i) { id_I[i] = i; pad_I[i] = i; }
        for (int w3 = 0; w3 < 3; ++w3)
            for (int hh = 0; hh < cfg.heads; ++hh)
                for (int d = 0; d < hd; ++d) slot_rows[(size_t)w3 * Hp + (size_t)hh * hdp + d] = w3 * cfg.H + hh * hd + d;
        for (int hh = 0; hh < cfg.heads; ++hh)
            for (int d = 0; d < hd; ++d) slot_cols[(size_t)hh * hdp + d] = hh * hd + d;
        const bool padded = hdp != hd || Ip != cfg.I;
        blocks.resize(cfg.depth);
        for (int i = 0; i < cfg.depth; ++i) {
            const std::string bp = p + "blocks." + std::to_string(i) + ".";
            VisionBlock& b = blocks[i];
            b.n1w = upload_vec(c, wt, bp + "norm1.weight", cfg.H); b.n1b = upload_vec(c, wt, bp + "norm1.bias", cfg.H);
            b.n2w = upload_vec(c, wt, bp + "norm2.weight", cfg.H); b.n2b = upload_vec(c, wt, bp + "norm2.bias", cfg.H);
            if (!padded) {
                b.qkv = lin(c, wt, bp + "attn.qkv", 3 * cfg.H, cfg.H);
                b.proj = lin(c, wt, bp + "attn.proj", cfg.H, cfg.H);
                b.fc1 = lin(c, wt, bp + "mlp.linear_fc1", cfg.I, cfg.H);
                b.fc2 = lin(c, wt, bp + "mlp.linear_fc2", cfg.H, cfg.I);
            } else {
                b.qkv = upload_linear_mapped(c, wt, bp + "attn.qkv.weight", 3 * cfg.H, cfg.H, slot_rows, id_H, bp + "attn.qkv.bias");
                b.proj = upload_linear_mapped(c, wt, bp + "attn.proj.weight", cfg.H, cfg.H, id_H, slot_cols, bp + "attn.proj.bias");
                b.fc1 = upload_linear_mapped(c, wt, bp + "mlp.linear_fc1.weight", cfg.I, cfg.H, pad_I, id_H, bp + "mlp.linear_fc1.bias");
                b.fc2 = upload_linear_mapped(c, wt, bp + "mlp.linear_fc2.weight", cfg.H, cfg.I, id_H, pad_I, bp + "mlp.linear_fc2.bias");
            }
        }
        merger = load_merger(c, wt, p + "merger.", false);
        for (size_t i = 0; i < cfg.deepstack.size(); ++i) ds_mergers.push_back(load_merger(c, wt, p + "deepstack_merger_list." + std::to_string(i) + ".", true));
        const size_t N = max_patches, Hv = cfg.H, m2 = (size_t)cfg.merge * cfg.merge;
        pix = c.alloc<float>(N * patch_dim); x = c.alloc<float>(N * Hv); xn = c.alloc<float>(N * Hv);
        qkv = c.alloc<float>(N * 3 * Hp); attn = c.alloc<float>(N * Hp); h = c.alloc<float>(N * Ip); mtmp = c.alloc<float>(N * Hv);
        image_embeds = c.alloc<float>(N / m2 * cfg.out_hidden);
        for (size_t i = 0; i < cfg.deepstack.size(); ++i) ds_out.push_back(c.alloc<float>(N / m2 * cfg.out_hidden));
        rowcol = c.alloc<int2>(N);
    }
    void set_trace(bool on) {
        if (on && !trace_buf) trace_buf = ctx->alloc<float>((size_t)(cfg.depth + 1) * max_patches * cfg.H);
        trace = on;
    }
    void gemm(int epi, const float* A, int lda, LinearW& W, const float* resid, int ldr, float* C, int ldc, int M, int act = ACT_NONE) {
        linear_gemm(*ctx, epi, A, lda, W, resid, ldr, C, ldc, M, act);
    }
    void run_merger(Merger& m, const float* in, int N, float* out) {
        Ctx& c = *ctx;
        cudaStream_t st = c.stream;
        const int m2 = cfg.merge * cfg.merge, Hm = cfg.H * m2, R = N / m2;
        if (c.presplit && gemm_on_tc(c, R, m.fc1) && gemm_on_tc(c, R, m.fc2)) {   // norm -> [hi|lo] -> fc1 (GELU epilogue writes [hi|lo]) -> fc2
            ActSplit n = act_split(c, 0, (size_t)R * Hm), hsp = act_split(c, 1, (size_t)R * Hm);
            if (m.post) layernorm_split_kernel<<<R, 256, 0, st>>>(in, m.nw, m.nb, 1e-6f, n.hi, n.lo, Hm);
            else layernorm_split_kernel<<<N, 256, 0, st>>>(in, m.nw, m.nb, 1e-6f, n.hi, n.lo, cfg.H);
            c.cnt.kernels++;
            linear_gemm_split(c, EPI_ACT, n.hi, n.lo, m.fc1, nullptr, 0, nullptr, 0, R, ACT_GELU_ERF, hsp, Hm);
            linear_gemm_split(c, EPI_STORE, hsp.hi, hsp.lo, m.fc2, nullptr, 0, out, cfg.out_hidden, R);
            return;
        }
        if (m.post) layernorm_kernel<<<N / m2, 256, 0, st>>>(in, m.nw, m.nb, 1e-6f, xn, Hm);
        else layernorm_kernel<<<N, 256, 0, st>>>(in, m.nw, m.nb, 1e-6f, xn, cfg.H);
        ctx->cnt.kernels++;
        gemm(EPI_ACT, xn, Hm, m.fc1, nullptr, 0, mtmp, Hm, N / m2, ACT_GELU_ERF);  // Activation::Gelu = erf (model.rs:131)
        gemm(EPI_STORE, mtmp, Hm, m.fc2, nullptr, 0, out, cfg.out_hidden, N / m2);
    }
    // Tensor parallelism (new design): every image (every cu_seqlens segment group) is independent through all blocks and
    // mergers, so a request with several images is sharded BY IMAGE over the ranks -- image i runs on rank (i * world) / n_images --
    // and the (tokens, out_hidden) embeddings + deepstack tensors are exchanged with one broadcast per owner.  A single image stays
    // replicated (its rows would have to be split by heads: not built).
    int tp_rank = 0, tp_world = 1;
    NcclApi::comm_t comm = nullptr;
    void set_tp(int rank, int world, NcclApi::comm_t cm) { tp_rank = rank; tp_world = world; comm = cm; }

    // pixel_values already in `pix` (N rows); grid: host (n_img x 3)
    void forward(int N, const std::vector<std::array<int, 3>>& grid) {
        const int n_img = (int)grid.size();
        if (tp_world <= 1 || n_img < 2 || comm == nullptr || trace) { forward_range(0, N, grid, N); return; }
        const int m2 = cfg.merge * cfg.merge;
        std::vector<int> first_patch(n_img + 1, 0);
        for (int i = 0; i < n_img; ++i) first_patch[i + 1] = first_patch[i] + grid[i][0] * grid[i][1] * grid[i][2];
        AHA_REQUIRE(first_patch[n_img] == N, "pixel_values rows do not match image_grid_thw");
        std::vector<int> lo(tp_world + 1, 0);   // rank r owns images [lo[r], lo[r + 1]): image i belongs to rank (i * world) / n_images (non-decreasing in i)
        for (int r = 0; r <= tp_world; ++r) {
            int i = 0;
            while (i < n_img && (int)(((long long)i * tp_world) / n_img) < r) ++i;
            lo[r] = i;
        }
        {
            const int a = lo[tp_rank], b = lo[tp_rank + 1];
            if (b > a) forward_range(first_patch[a], first_patch[b] - first_patch[a], std::vector<std::array<int, 3>>(grid.begin() + a, grid.begin() + b), N);
        }
        NcclApi& api = NcclApi::get();
        for (int r = 0; r < tp_world; ++r) {
            const int a = lo[r], b = lo[r + 1];
            if (b <= a) continue;
            const size_t off = (size_t)(first_patch[a] / m2) * cfg.out_hidden, cnt = (size_t)((first_patch[b] - first_patch[a]) / m2) * cfg.out_hidden;
            api.check(api.Broadcast(image_embeds + off, image_embeds + off, cnt, NcclApi::kFloat32, r, comm, ctx->stream), "ncclBroadcast(image embeds)");
            for (float* d : ds_out) api.check(api.Broadcast(d + off, d + off, cnt, NcclApi::kFloat32, r, comm, ctx->stream), "ncclBroadcast(deepstack)");
        }
        last_N = N;
    }
    // the tower on patches [p_start, p_start + N) of `pix` (whole images `grid`); outputs land at their global token offsets
    void forward_range(int p_start, int N, const std::vector<std::array<int, 3>>& grid, int N_total) {
        Ctx& c = *ctx;
        cudaStream_t st = c.stream;
        const int Hv = cfg.H;
        const int m2_ = cfg.merge * cfg.merge;
        AHA_REQUIRE(N_total <= max_patches, "image needs " + std::to_string(N_total) + " patches, max_patches is " + std::to_string(max_patches));
        float* const emb_out = image_embeds + (size_t)(p_start / m2_) * cfg.out_hidden;
        gemm(EPI_STORE, pix + (size_t)p_start * patch_dim, patch_dim, patch, nullptr, 0, x, Hv, N);
        int p0 = 0;
        std::vector<std::pair<int, int>> segs;  // (start, len) -- cu_seqlens = cumsum(h*w repeated t), model.rs:709-720
        for (auto& g : grid) {
            const int t = g[0], h_ = g[1], w_ = g[2];
            AHA_REQUIRE(h_ % cfg.merge == 0 && w_ % cfg.merge == 0, "grid h/w must be multiples of spatial_merge_size");
            const float sh = h_ > 1 ? (float)(n_side - 1) / (float)(h_ - 1) : 0.f;
            const float sw = w_ > 1 ? (float)(n_side - 1) / (float)(w_ - 1) : 0.f;
            vit_pos_embed_kernel<<<t * h_ * w_, 256, 0, st>>>(x, pos_embed, rowcol, p0, t, h_, w_, cfg.merge, n_side, sh, sw, Hv);
            c.cnt.kernels++;
            for (int k = 0; k < t; ++k) segs.push_back({p0 + k * h_ * w_, h_ * w_});
            p0 += t * h_ * w_;
        }
        AHA_REQUIRE(p0 == N, "pixel_values rows do not match image_grid_thw");
        if (trace) AHA_CUDA_CHECK(cudaMemcpyAsync(trace_buf, x, (size_t)N * Hv * sizeof(float), cudaMemcpyDeviceToDevice, st));
        const float scaling = (float)(1.0 / std::sqrt((double)hd));
        for (int i = 0; i < cfg.depth; ++i) {
            VisionBlock& b = blocks[i];
            // pre-split data flow (Ctx::presplit): each producer writes the fp16 hi + lo halves its GEMM reads (see text_model.cuh prefill)
            const bool ps = c.presplit && c.attn_impl == 0 && gemm_on_tc(c, N, b.qkv) && gemm_on_tc(c, N, b.proj) && gemm_on_tc(c, N, b.fc1) && gemm_on_tc(c, N, b.fc2);
            ActSplit ao;
            if (ps) {
                ActSplit n1 = act_split(c, 0, (size_t)N * Hv);
                layernorm_split_kernel<<<N, 256, 0, st>>>(x, b.n1w, b.n1b, 1e-6f, n1.hi, n1.lo, Hv); c.cnt.kernels++;
                linear_gemm_split(c, EPI_STORE, n1.hi, n1.lo, b.qkv, nullptr, 0, qkv, 3 * Hp, N);
                ao = act_split(c, 1, (size_t)N * std::max(Hp, Ip));
            } else {
                layernorm_kernel<<<N, 256, 0, st>>>(x, b.n1w, b.n1b, 1e-6f, xn, Hv); c.cnt.kernels++;
                gemm(EPI_STORE, xn, Hv, b.qkv, nullptr, 0, qkv, 3 * Hp, N);
            }
            vit_rope_kernel<<<N, 256, 0, st>>>(qkv, rowcol, inv_freq, hd, hdp, cfg.heads, Hp); c.cnt.kernels++;
            for (auto& sg : segs) {
                FlashArgs fa;
                fa.q = qkv; fa.q_tok_stride = 3 * Hp; fa.q_head_stride = hdp;
                fa.kv.k = qkv + Hp; fa.kv.v = qkv + 2 * Hp; fa.kv.page_table = nullptr; fa.kv.page_shift = 0; fa.kv.page_stride = 0;
                fa.kv.tok_stride = 3 * Hp; fa.kv.head_stride = hdp;
                fa.out = attn; fa.o_tok_stride = Hp; fa.o_head_stride = hdp;
                fa.out_hi = ao.hi; fa.out_lo = ao.lo;
                fa.Sq = sg.second; fa.Skv = sg.second; fa.q0 = sg.first; fa.kv0 = sg.first; fa.groups = 1; fa.scaling = scaling;
                if (hdp == 64) flash_dispatch<64>(c, fa, cfg.heads, false);
                else flash_dispatch<128>(c, fa, cfg.heads, false);
            }
            if (ps) {
                linear_gemm_split(c, EPI_RESID, ao.hi, ao.lo, b.proj, x, Hv, x, Hv, N);
                ActSplit n2 = act_split(c, 0, (size_t)N * Hv);
                layernorm_split_kernel<<<N, 256, 0, st>>>(x, b.n2w, b.n2b, 1e-6f, n2.hi, n2.lo, Hv); c.cnt.kernels++;
                linear_gemm_split(c, EPI_ACT, n2.hi, n2.lo, b.fc1, nullptr, 0, nullptr, 0, N, cfg.act, ao, Ip);
                linear_gemm_split(c, EPI_RESID, ao.hi, ao.lo, b.fc2, x, Hv, x, Hv, N);
            } else {
                gemm(EPI_RESID, attn, Hp, b.proj, x, Hv, x, Hv, N);
                layernorm_kernel<<<N, 256, 0, st>>>(x, b.n2w, b.n2b, 1e-6f, xn, Hv); c.cnt.kernels++;
                gemm(EPI_ACT, xn, Hv, b.fc1, nullptr, 0, h, Ip, N, cfg.act);   // padded rows: act(0 + 0) = 0 for gelu / silu, and fc2's padded columns are zero anyway
                gemm(EPI_RESID, h, Ip, b.fc2, x, Hv, x, Hv, N);
            }
            if (trace) AHA_CUDA_CHECK(cudaMemcpyAsync(trace_buf + (size_t)(i + 1) * max_patches * Hv, x, (size_t)N * Hv * sizeof(float), cudaMemcpyDeviceToDevice, st));
            for (size_t k = 0; k < cfg.deepstack.size(); ++k)
                if (cfg.deepstack[k] == i) run_merger(ds_mergers[k], x, N, ds_out[k] + (size_t)(p_start / m2_) * cfg.out_hidden);
        }
        run_merger(merger, x, N, emb_out);
        last_N = N_total;
        AHA_CUDA_CHECK(cudaGetLastError());
    }
};

}  // namespace aha
