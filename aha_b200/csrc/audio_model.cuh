// audio_model.cuh -- Qwen3-ASR audio path: whisper-style log-mel frontend and the conv + transformer audio tower.
// Reference: /root/reference/src/models/feature_extractor/feature_extraction_whisper.rs:65-115,
// /root/reference/src/utils/audio_utils.rs:1064-1083,1218-1347,1483-1503, src/utils/tensor_utils.rs:525-549,
// /root/reference/src/models/qwen3_asr/model.rs:32-227, src/position_embed/sinusoidal_pe.rs:6-58,
// /root/reference/src/models/common/modules.rs:127-242 (NaiveAttention).
#pragma once
#include "text_model.cuh"
#include "vision_model.cuh"

namespace aha {

struct AudioCfg {
    int d_model = 0, heads = 0, ffn = 0, layers = 0, n_window = 50, mel = 128, out_dim = 0, down = 0, conv_chunksize = 500;
    int act = ACT_GELU_ERF;
    static AudioCfg from_json(const Json& j) {
        AudioCfg c;
        c.d_model = j.integer("d_model"); c.heads = j.integer("encoder_attention_heads"); c.ffn = j.integer("encoder_ffn_dim");
        c.layers = j.integer("encoder_layers"); c.n_window = j.integer_or("n_window", 50); c.mel = j.integer_or("num_mel_bins", 128);
        c.out_dim = j.integer("output_dim"); c.down = j.integer("downsample_hidden_size"); c.conv_chunksize = j.integer_or("conv_chunksize", 500);
        c.act = VisionCfg::act_from(j.string_or("activation_function", "gelu"));
        return c;
    }
};

// ------------------------------------------------------------------------------------------ mel frontend
constexpr int kNfft = 400, kHop = 160, kBins = 201;

// index into the reference's pad_reflect_last_dim((200,200)) output (tensor_utils.rs:525-549): the right pad is
// cut from the already left-padded tensor at start = n - pad_r, reproduced as written.
__device__ __forceinline__ float padded_sample(const float* __restrict__ w, int n, int i) {
    constexpr int P = kNfft / 2;
    auto left_padded = [&](int j) -> float { return j < P ? w[P - j] : w[j - P]; };  // length n + P
    if (i < n + P) return left_padded(i);
    const int r = i - (n + P);           // 0..P-1
    return left_padded(n - 1 - r);
}

// One block per frame: window, 400-point real DFT power (re^2 + im^2 = realfft norm_sqr), mel projection,
// clamp(1e-10), log10 via ln * (1/ln 10) (modules.rs:1256-1258).  Writes log10 mel (n_mels, n_frames).
__global__ void __launch_bounds__(256) mel_frames_kernel(const float* __restrict__ wave, int n, const float* __restrict__ window,
                                                        const float* __restrict__ twc, const float* __restrict__ tws,
                                                        const float* __restrict__ melw /*[n_mels][201]*/, int n_mels,
                                                        float* __restrict__ out, int n_frames) {
    __shared__ float fr[kNfft];
    __shared__ float c[kNfft], s[kNfft];
    __shared__ float pw[kBins];
    const int f = blockIdx.x;
    for (int i = threadIdx.x; i < kNfft; i += 256) {
        fr[i] = padded_sample(wave, n, f * kHop + i) * window[i];
        c[i] = twc[i]; s[i] = tws[i];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kBins; k += 256) {
        float re = 0.f, im = 0.f;
        int m = 0;
        for (int t = 0; t < kNfft; ++t) {
            re = fmaf(fr[t], c[m], re);
            im = fmaf(-fr[t], s[m], im);
            m += k; if (m >= kNfft) m -= kNfft;
        }
        pw[k] = re * re + im * im;
    }
    __syncthreads();
    for (int mm = threadIdx.x; mm < n_mels; mm += 256) {
        const float* wr = melw + (size_t)mm * kBins;
        float acc = 0.f;
        for (int k = 0; k < kBins; ++k) acc = fmaf(wr[k], pw[k], acc);
        acc = fmaxf(acc, 1e-10f);
        out[(size_t)mm * n_frames + f] = logf(acc) * 0.43429448190325176f;
    }
}
__global__ void max_reduce_kernel(const float* __restrict__ x, size_t n, float* __restrict__ out) {
    __shared__ float red[32];
    float m = -INFINITY;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, x[i]);
    m = block_max(m, red);
    if (threadIdx.x == 0) *out = m;
}
// maximum(x, max-8) then (x + 4) * (1/4)   (feature_extraction_whisper.rs:110-113)
__global__ void mel_finish_kernel(float* __restrict__ x, size_t n, const float* __restrict__ mx) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = (fmaxf(x[i], *mx - 8.0f) + 4.0f) * 0.25f;
}

// ------------------------------------------------------------------------------------------ conv stack helpers
// mel (n_mels, T) -> chunked NHWC input [B][n_mels][cw][1] with zero padding of the last chunk (model.rs:171-198)
__global__ void audio_chunk_kernel(const float* __restrict__ mel, int n_mels, int T, int cw, float* __restrict__ out) {
    const int b = blockIdx.x, f = blockIdx.y;
    for (int t = threadIdx.x; t < cw; t += blockDim.x) {
        const int tt = b * cw + t;
        out[((size_t)b * n_mels + f) * cw + t] = tt < T ? mel[(size_t)f * T + tt] : 0.f;
    }
}
// im2col for Conv2d(k=3, stride=2, pad=1) over NHWC input [B][IH][IW][C] -> A[M = B*OH*OW][Kp], column = c*9 + kh*3 + kw
// (matches the (oc, ic, kh, kw) weight flattening); columns >= C*9 are zero padding up to Kp.
__global__ void im2col_k3s2_kernel(const float* __restrict__ in, int B, int IH, int IW, int C, int OH, int OW, int Kp, float* __restrict__ A) {
    const size_t m = blockIdx.x;
    const int ow = (int)(m % OW), oh = (int)((m / OW) % OH), b = (int)(m / ((size_t)OW * OH));
    for (int col = threadIdx.x; col < Kp; col += blockDim.x) {
        float v = 0.f;
        if (col < C * 9) {
            const int c = col / 9, kh = (col % 9) / 3, kw = col % 3;
            const int ih = oh * 2 - 1 + kh, iw = ow * 2 - 1 + kw;
            if (ih >= 0 && ih < IH && iw >= 0 && iw < IW) v = in[(((size_t)b * IH + ih) * IW + iw) * C + c];
        }
        A[m * Kp + col] = v;
    }
}
// NHWC conv output [B][F][T][C] -> [B*T][C*F] with column c*F + f  (permute (0,3,1,2) + reshape, model.rs:206-211)
__global__ void audio_permute_kernel(const float* __restrict__ in, int B, int F, int T, int C, float* __restrict__ out) {
    const int bt = blockIdx.x, b = bt / T, t = bt % T;
    for (int col = threadIdx.x; col < C * F; col += blockDim.x) {
        const int c = col / F, f = col % F;
        out[(size_t)bt * C * F + col] = in[(((size_t)b * F + f) * T + t) * C + c];
    }
}
// x[b*T + t, :] += [sin | cos](t * inv_freq), positions restart at 0 per chunk (sinusoidal_pe.rs:22-58)
__global__ void sinusoid_add_kernel(float* __restrict__ x, int T, int D, const float* __restrict__ inv_freq) {
    const int row = blockIdx.x, t = row % T;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const int j = i < D / 2 ? i : i - D / 2;
        const float ang = (float)t * inv_freq[j];
        x[(size_t)row * D + i] += (i < D / 2) ? sinf(ang) : cosf(ang);
    }
}

struct AudioLayer {
    float *ln1w, *ln1b, *ln2w, *ln2b;
    LinearW qkv, out, fc1, fc2;
};

struct AudioModel {
    AudioCfg cfg;
    Ctx* ctx = nullptr;
    // frontend
    float *window = nullptr, *twc = nullptr, *tws = nullptr, *melw = nullptr, *d_wave = nullptr, *d_mel = nullptr, *d_max = nullptr;
    int max_samples = 0, max_frames = 0;
    // tower
    LinearW conv[3], conv_out, proj1, proj2;
    int convK[3] = {0, 0, 0};
    float *lnpw = nullptr, *lnpb = nullptr, *pe_inv = nullptr;
    std::vector<AudioLayer> layers;
    float *chunks = nullptr, *col = nullptr, *act0 = nullptr, *act1 = nullptr, *x = nullptr, *xn = nullptr, *qkv = nullptr, *attn = nullptr, *h = nullptr;
    float* audio_embeds = nullptr;
    int max_tokens = 0, last_tokens = 0;
    bool trace = false;
    float* trace_buf = nullptr;

    static int out_len(int n) { return (n + 2 - 3) / 2 + 1; }
    // get_feat_extract_output_lengths, /root/reference/src/models/qwen3_asr/processor.rs:187-195
    static int feat_len(int audio_len) {
        const int leave = audio_len % 100;
        if (leave > 0) { const int feat = (leave - 1) / 2 + 1; return ((feat - 1) / 2 + 1 - 1) / 2 + 1 + (audio_len / 100) * 13; }
        return (audio_len / 100) * 13;
    }

    void load(Ctx& c, const AudioCfg& cf, const WeightTable& wt, const std::string& p, int max_frames_) {
        ctx = &c; cfg = cf; max_frames = max_frames_;
        AHA_REQUIRE(cfg.d_model / cfg.heads == 64, "audio encoder head_dim must be 64");
        AHA_REQUIRE(cfg.n_window == 50, "n_window must be 50 (100-frame chunks -> 13 tokens, processor.rs:187-195)");
        // ---- frontend tables
        {
            std::vector<float> w(kNfft), tc(kNfft), ts(kNfft);
            const double n = kNfft - 1.0;
            for (int k = 0; k < kNfft; ++k) {
                const double i = (double)(1 - kNfft + 2 * k);   // audio_utils.rs:1071-1082, symmetric Hann
                w[k] = (float)(0.5 + 0.5 * std::cos(M_PI * i / n));
                tc[k] = (float)std::cos(2.0 * M_PI * k / kNfft);
                ts[k] = (float)std::sin(2.0 * M_PI * k / kNfft);
            }
            window = upload(c, w); twc = upload(c, tc); tws = upload(c, ts);
            melw = upload(c, mel_filter_bank(cfg.mel));
            max_samples = max_frames * kHop;
            d_wave = c.alloc<float>(max_samples);
            d_mel = c.alloc<float>((size_t)cfg.mel * (max_frames + 1));
            d_max = c.alloc<float>(1);
        }
        // ---- tower
        const int C = cfg.down, D = cfg.d_model;
        for (int i = 0; i < 3; ++i) {
            const int cin = i == 0 ? 1 : C;
            const int K = cin * 9, Kp = ceil_div(K, 16) * 16;
            convK[i] = Kp;
            const std::string n = p + "conv2d" + std::to_string(i + 1);
            const aha_tensor_desc& d = wt.get(n + ".weight");
            AHA_REQUIRE((int64_t)WeightTable::numel(d) == (int64_t)C * K, "conv weight size mismatch: " + n);
            std::vector<__half> st((size_t)C * Kp, __float2half_rn(0.f));
            for (int oc = 0; oc < C; ++oc)
                for (int k = 0; k < K; ++k) st[(size_t)oc * Kp + k] = __float2half_rn(WeightTable::at(d, (size_t)oc * K + k));
            conv[i].w = upload(c, st); conv[i].N = C; conv[i].K = Kp;
            conv[i].b = upload_vec(c, wt, n + ".bias", C);
        }
        const int F3 = out_len(out_len(out_len(cfg.mel)));
        conv_out = upload_linear(c, wt, {{p + "conv_out.weight", D, 0, D}}, (int64_t)C * F3, 0, (int64_t)C * F3, false, {});
        layers.resize(cfg.layers);
        for (int l = 0; l < cfg.layers; ++l) {
            const std::string lp = p + "layers." + std::to_string(l) + ".";
            AudioLayer& A = layers[l];
            A.ln1w = upload_vec(c, wt, lp + "self_attn_layer_norm.weight", D); A.ln1b = upload_vec(c, wt, lp + "self_attn_layer_norm.bias", D);
            A.ln2w = upload_vec(c, wt, lp + "final_layer_norm.weight", D); A.ln2b = upload_vec(c, wt, lp + "final_layer_norm.bias", D);
            A.qkv = upload_linear(c, wt, {{lp + "self_attn.q_proj.weight", D, 0, D}, {lp + "self_attn.k_proj.weight", D, 0, D}, {lp + "self_attn.v_proj.weight", D, 0, D}},
                                  D, 0, D, false, {lp + "self_attn.q_proj.bias", lp + "self_attn.k_proj.bias", lp + "self_attn.v_proj.bias"});
            A.out = VisionModel::lin(c, wt, lp + "self_attn.out_proj", D, D);
            A.fc1 = VisionModel::lin(c, wt, lp + "fc1", cfg.ffn, D);
            A.fc2 = VisionModel::lin(c, wt, lp + "fc2", D, cfg.ffn);
        }
        lnpw = upload_vec(c, wt, p + "ln_post.weight", D); lnpb = upload_vec(c, wt, p + "ln_post.bias", D);
        proj1 = VisionModel::lin(c, wt, p + "proj1", D, D);
        proj2 = VisionModel::lin(c, wt, p + "proj2", cfg.out_dim, D);
        std::vector<float> inv(D / 2);
        for (int j = 0; j < D / 2; ++j) inv[j] = 1.0f / powf(10000.0f, (float)(2 * j) / (float)D);
        pe_inv = upload(c, inv);
        // workspaces
        const int cw = cfg.n_window * 2;
        const size_t B = ceil_div(max_frames, cw);
        const int H1 = out_len(cfg.mel), W1 = out_len(cw), H2 = out_len(H1), W2 = out_len(W1), H3 = out_len(H2), W3 = out_len(W2);
        max_tokens = (int)B * W3;
        chunks = c.alloc<float>(B * cfg.mel * cw);
        const size_t m1 = B * H1 * W1, m2 = B * H2 * W2, m3 = B * H3 * W3;
        col = c.alloc<float>(std::max(std::max(m1 * (size_t)convK[0], m2 * (size_t)convK[1]), std::max(m3 * (size_t)convK[2], (size_t)max_tokens * C * H3)));
        act0 = c.alloc<float>(std::max(m1, m3) * C); act1 = c.alloc<float>(m2 * C);
        x = c.alloc<float>((size_t)max_tokens * D); xn = c.alloc<float>((size_t)max_tokens * D);
        qkv = c.alloc<float>((size_t)max_tokens * 3 * D); attn = c.alloc<float>((size_t)max_tokens * D); h = c.alloc<float>((size_t)max_tokens * cfg.ffn);
        audio_embeds = c.alloc<float>((size_t)max_tokens * cfg.out_dim);
    }
    void set_trace(bool on) {
        if (on && !trace_buf) trace_buf = ctx->alloc<float>((size_t)(cfg.layers + 1) * max_tokens * cfg.d_model);
        trace = on;
    }

    // slaney mel filter bank, f32 arithmetic as audio_utils.rs:1158-1301; returned transposed: [n_mels][201]
    static std::vector<float> mel_filter_bank(int n_mels) {
        auto hz2mel = [](float f) -> float {
            const float logstep = 27.0f / logf(6.4f);
            return f >= 1000.0f ? 15.0f + logf(f / 1000.0f) * logstep : 3.0f * f / 200.0f;
        };
        auto mel2hz = [](float m) -> float {
            const float logstep = logf(6.4f) / 27.0f;
            return m >= 15.0f ? 1000.0f * expf(logstep * (m - 15.0f)) : 200.0f * m / 3.0f;
        };
        const float mmin = hz2mel(0.0f), mmax = hz2mel(8000.0f);
        const int nf = n_mels + 2;
        std::vector<float> ff(nf), fft(kBins);
        const float mstep = (mmax - mmin) / (float)(nf - 1);
        for (int i = 0; i < nf; ++i) ff[i] = mel2hz(mmin + (float)i * mstep);
        const float fstep = (16000.0f / 2.0f - 0.0f) / (float)(kBins - 1);
        for (int i = 0; i < kBins; ++i) fft[i] = 0.0f + (float)i * fstep;
        std::vector<float> out((size_t)n_mels * kBins);
        for (int m = 0; m < n_mels; ++m) {
            const float enorm = 2.0f / (ff[m + 2] - ff[m]);
            for (int k = 0; k < kBins; ++k) {
                const float down = (-(ff[m] - fft[k])) / (ff[m + 1] - ff[m]);
                const float up = (ff[m + 2] - fft[k]) / (ff[m + 2] - ff[m + 1]);
                out[(size_t)m * kBins + k] = fmaxf(fminf(down, up), 0.0f) * enorm;
            }
        }
        return out;
    }

    // wave (host) -> d_mel (n_mels, frames); returns frames
    int mel_from_host(const float* wave, size_t n) {
        Ctx& c = *ctx;
        AHA_REQUIRE(n > (size_t)kNfft / 2, "audio too short for reflect padding");
        AHA_REQUIRE(n <= (size_t)max_samples, "audio longer than max_frames allows");
        AHA_CUDA_CHECK(cudaMemcpyAsync(d_wave, wave, n * sizeof(float), cudaMemcpyHostToDevice, c.stream));
        const int frames = (int)(1 + n / kHop) - 1;  // n_frames - 1 (last frame dropped, :104-105)
        AHA_REQUIRE(frames >= 1, "audio too short");
        mel_frames_kernel<<<frames, 256, 0, c.stream>>>(d_wave, (int)n, window, twc, tws, melw, cfg.mel, d_mel, frames);
        const size_t tot = (size_t)cfg.mel * frames;
        max_reduce_kernel<<<1, 1024, 0, c.stream>>>(d_mel, tot, d_max);
        mel_finish_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c.stream>>>(d_mel, tot, d_max);
        c.cnt.kernels += 3;
        AHA_CUDA_CHECK(cudaGetLastError());
        return frames;
    }

    void gemm(int epi, const float* A, int lda, LinearW& W, const float* resid, int ldr, float* C, int ldc, int M, int act = ACT_NONE) {
        linear_gemm(*ctx, epi, A, lda, W, resid, ldr, C, ldc, M, act);
    }

    // mel: device (n_mels, T).  Returns token count; result in audio_embeds.
    int forward(const float* mel, int T) {
        Ctx& c = *ctx;
        cudaStream_t st = c.stream;
        AHA_REQUIRE(T >= 1 && T <= max_frames, "mel length out of range");
        const int cw = cfg.n_window * 2, C = cfg.down, D = cfg.d_model;
        const int B = ceil_div(T, cw);
        std::vector<int> lens(B, cw);
        if (T % cw) lens[B - 1] = T % cw;
        int total = 0;
        for (int L : lens) total += feat_len(L);
        audio_chunk_kernel<<<dim3(B, cfg.mel), 128, 0, st>>>(mel, cfg.mel, T, cw, chunks); c.cnt.kernels++;
        // three Conv2d(k3,s2,p1) + bias + gelu (tanh approx: Tensor::gelu(), model.rs:200-202), NHWC activations.
        // conv_chunksize only bounds the reference's batch per call; the arithmetic is per-sample, so one pass is identical.
        int IH = cfg.mel, IW = cw, Cin = 1;
        const float* in = chunks;
        float* outs[3] = {act0, act1, act0};
        for (int i = 0; i < 3; ++i) {
            const int OH = out_len(IH), OW = out_len(IW);
            const size_t M = (size_t)B * OH * OW;
            im2col_k3s2_kernel<<<(unsigned)M, 128, 0, st>>>(in, B, IH, IW, Cin, OH, OW, convK[i], col); c.cnt.kernels++;
            gemm(EPI_ACT, col, convK[i], conv[i], nullptr, 0, outs[i], C, (int)M, ACT_GELU_TANH);
            in = outs[i]; IH = OH; IW = OW; Cin = C;
        }
        const int F3 = IH, T3 = IW;  // (16, 13)
        audio_permute_kernel<<<B * T3, 256, 0, st>>>(in, B, F3, T3, C, col); c.cnt.kernels++;
        gemm(EPI_STORE, col, C * F3, conv_out, nullptr, 0, x, D, B * T3);
        sinusoid_add_kernel<<<B * T3, 256, 0, st>>>(x, T3, D, pe_inv); c.cnt.kernels++;
        // keep the first `total` rows (model.rs:214-216); encoder over the whole sequence, no mask (:218-220)
        const int N = total;
        if (trace) AHA_CUDA_CHECK(cudaMemcpyAsync(trace_buf, x, (size_t)N * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
        const float scaling = (float)(1.0 / std::sqrt(64.0));
        for (int l = 0; l < cfg.layers; ++l) {
            AudioLayer& A = layers[l];
            layernorm_kernel<<<N, 256, 0, st>>>(x, A.ln1w, A.ln1b, 1e-5f, xn, D); c.cnt.kernels++;
            gemm(EPI_STORE, xn, D, A.qkv, nullptr, 0, qkv, 3 * D, N);
            FlashArgs fa;
            fa.q = qkv; fa.q_tok_stride = 3 * D; fa.q_head_stride = 64;
            fa.kv.k = qkv + D; fa.kv.v = qkv + 2 * D; fa.kv.page_table = nullptr; fa.kv.page_shift = 0; fa.kv.page_stride = 0;
            fa.kv.tok_stride = 3 * D; fa.kv.head_stride = 64;
            fa.out = attn; fa.o_tok_stride = D; fa.o_head_stride = 64;
            fa.Sq = N; fa.Skv = N; fa.q0 = 0; fa.kv0 = 0; fa.groups = 1; fa.scaling = scaling;
            flash_dispatch<64>(c, fa, cfg.heads, false);
            gemm(EPI_RESID, attn, D, A.out, x, D, x, D, N);
            layernorm_kernel<<<N, 256, 0, st>>>(x, A.ln2w, A.ln2b, 1e-5f, xn, D); c.cnt.kernels++;
            gemm(EPI_ACT, xn, D, A.fc1, nullptr, 0, h, cfg.ffn, N, cfg.act);
            gemm(EPI_RESID, h, cfg.ffn, A.fc2, x, D, x, D, N);
            if (trace) AHA_CUDA_CHECK(cudaMemcpyAsync(trace_buf + (size_t)(l + 1) * max_tokens * D, x, (size_t)N * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
        layernorm_kernel<<<N, 256, 0, st>>>(x, lnpw, lnpb, 1e-5f, xn, D); c.cnt.kernels++;
        gemm(EPI_ACT, xn, D, proj1, nullptr, 0, h, D, N, cfg.act);
        gemm(EPI_STORE, h, D, proj2, nullptr, 0, audio_embeds, cfg.out_dim, N);
        last_tokens = N;
        AHA_CUDA_CHECK(cudaGetLastError());
        return N;
    }
};

}  // namespace aha
