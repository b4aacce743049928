// preprocess.cuh -- the host-facing processor pieces of the path: image resize on the GPU, and the integer / scalar host
// routines of Qwen3VLProcessor and Qwen3AsrProcessor.
//
// Replaces (reference):
//   img_smart_resize                       /root/reference/src/utils/img_utils.rs:297-331 (+ round/floor/ceil_by_factor, utils/mod.rs:392-405)
//   img.resize_exact(w, h, CatmullRom)     src/models/qwen3vl/processor.rs:167 -> `image` crate 0.25 imageops::resize:
//                                          vertical_sample into f32, then horizontal_sample back to u8 (third-party, not under
//                                          /root/reference; restated from its published algorithm, see oracle/qwen3vl.py)
//   <|image_pad|> / <|audio_pad|> expansion   qwen3vl/processor.rs:386-399, qwen3_asr/processor.rs:93-97 (on token ids here:
//                                          the tokenizer is outside the path)
//   float_range_normalize                  src/models/common/modules.rs:1353-1368
//   split_audio_into_chunks                src/utils/audio_utils.rs:1743-1760
//   get_feat_extract_output_lengths        src/models/qwen3_asr/processor.rs:187-195
//   resample_simple (sinc, Hann window)    src/utils/audio_utils.rs:66-255 (what load_audio_with_resample applies to bring any input to 16 kHz)
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "common.cuh"

namespace aha {

// ---------------------------------------------------------------------------------------------- img_smart_resize (host, f32 like the reference)
inline uint32_t round_by_factor(float num, uint32_t factor) { return (uint32_t)std::round(num / (float)factor) * factor; }   // f32::round: half away from zero
inline uint32_t floor_by_factor(float num, uint32_t factor) { return (uint32_t)std::floor(num / (float)factor) * factor; }
inline uint32_t ceil_by_factor(float num, uint32_t factor) { return (uint32_t)std::ceil(num / (float)factor) * factor; }

inline void img_smart_resize(uint32_t img_h, uint32_t img_w, uint32_t factor, uint32_t min_pixels, uint32_t max_pixels, uint32_t& h_bar, uint32_t& w_bar) {
    AHA_REQUIRE(img_h > 0 && img_w > 0 && factor > 0, "img_smart_resize: empty image");
    if (std::max(img_h, img_w) / std::min(img_h, img_w) > 200) throw std::runtime_error("absolute aspect ratio mush be smaller than 200");
    h_bar = std::max(factor, round_by_factor((float)img_h, factor));
    w_bar = std::max(factor, round_by_factor((float)img_w, factor));
    if ((uint64_t)h_bar * w_bar > max_pixels) {
        const float beta = std::sqrt((float)(img_h * img_w) / (float)max_pixels);
        h_bar = std::max(factor, floor_by_factor((float)img_h / beta, factor));
        w_bar = std::max(factor, floor_by_factor((float)img_w / beta, factor));
    } else if ((uint64_t)h_bar * w_bar < min_pixels) {
        const float beta = std::sqrt((float)min_pixels / (float)(img_h * img_w));
        h_bar = ceil_by_factor((float)img_h * beta, factor);
        w_bar = ceil_by_factor((float)img_w * beta, factor);
    }
}

// ---------------------------------------------------------------------------------------------- CatmullRom resize (device)
// bc_cubic_spline(x, b = 0, c = 0.5) / 6: the coefficients 9, -15, 6 and -3, 15, -24, 12 are what (12 - 9b - 6c) ... evaluate
// to exactly in f32.  Every product and sum is rounded separately (__fmul_rn / __fadd_rn): the reference is Rust, which never
// contracts a * b + c into an FMA, and the last bit decides a rounding to u8 now and then.
__device__ __forceinline__ float catmullrom_kernel(float x) {
    const float a = fabsf(x);
    float k;
    if (a < 1.0f) {
        const float a2 = __fmul_rn(a, a), a3 = __fmul_rn(a2, a);
        k = __fadd_rn(__fadd_rn(__fmul_rn(9.0f, a3), __fmul_rn(-15.0f, a2)), 6.0f);
    } else if (a < 2.0f) {
        const float a2 = __fmul_rn(a, a), a3 = __fmul_rn(a2, a);
        k = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(-3.0f, a3), __fmul_rn(15.0f, a2)), __fmul_rn(-24.0f, a)), 12.0f);
    } else k = 0.0f;
    return __fdiv_rn(k, 6.0f);
}

constexpr int kResizeMaxTaps = 64;   // 2 * support * max(1, downscale ratio) + 2: covers a 15x downscale

// One output coordinate of imageops::{vertical,horizontal}_sample: the input window [left, right) and its normalised weights.
__device__ __forceinline__ void resize_taps(int out_i, int in_n, int out_n, int& left, int& right, float* ws) {
    const float ratio = __fdiv_rn((float)in_n, (float)out_n);
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float src_support = __fmul_rn(2.0f, sratio);            // CatmullRom support = 2
    float input = __fmul_rn(__fadd_rn((float)out_i, 0.5f), ratio);
    long long l = (long long)floorf(__fsub_rn(input, src_support));
    l = l < 0 ? 0 : (l > in_n - 1 ? in_n - 1 : l);
    long long r = (long long)ceilf(__fadd_rn(input, src_support));
    r = r < l + 1 ? l + 1 : (r > in_n ? in_n : r);
    left = (int)l; right = (int)r;
    input = __fsub_rn(input, 0.5f);
    float sum = 0.0f;
    const int n = min(right - left, kResizeMaxTaps);
    for (int i = 0; i < n; ++i) {
        const float w = catmullrom_kernel(__fdiv_rn(__fsub_rn((float)(left + i), input), sratio));
        ws[i] = w;
        sum = __fadd_rn(sum, w);
    }
    for (int i = 0; i < n; ++i) ws[i] = __fdiv_rn(ws[i], sum);
    right = left + n;
}

// vertical_sample: u8 HWC (h, w, 3) -> f32 (new_h, w, 3); one block per output row
__global__ void resize_vertical_kernel(const uint8_t* __restrict__ img, int h, int w, int new_h, float* __restrict__ tmp) {
    __shared__ float ws[kResizeMaxTaps];
    __shared__ int lr[2];
    const int outy = blockIdx.x;
    if (threadIdx.x == 0) { int l, r; resize_taps(outy, h, new_h, l, r, ws); lr[0] = l; lr[1] = r; }
    __syncthreads();
    const int left = lr[0], n = lr[1] - lr[0];
    for (int xc = threadIdx.x; xc < w * 3; xc += blockDim.x) {
        float t = 0.0f;
        for (int i = 0; i < n; ++i) t = __fadd_rn(t, __fmul_rn((float)img[(size_t)(left + i) * w * 3 + xc], ws[i]));
        tmp[(size_t)outy * w * 3 + xc] = t;
    }
}
// horizontal_sample: f32 (h, w, 3) -> u8 (h, new_w, 3), clamp to [0, 255] and round half away from zero; one block per output column
__global__ void resize_horizontal_kernel(const float* __restrict__ tmp, int h, int w, int new_w, uint8_t* __restrict__ out) {
    __shared__ float ws[kResizeMaxTaps];
    __shared__ int lr[2];
    const int outx = blockIdx.x;
    if (threadIdx.x == 0) { int l, r; resize_taps(outx, w, new_w, l, r, ws); lr[0] = l; lr[1] = r; }
    __syncthreads();
    const int left = lr[0], n = lr[1] - lr[0];
    for (int yc = threadIdx.x; yc < h * 3; yc += blockDim.x) {
        const int y = yc / 3, c = yc % 3;
        float t = 0.0f;
        for (int i = 0; i < n; ++i) t = __fadd_rn(t, __fmul_rn(tmp[((size_t)y * w + left + i) * 3 + c], ws[i]));
        t = fminf(fmaxf(t, 0.0f), 255.0f);
        out[((size_t)y * new_w + outx) * 3 + c] = (uint8_t)roundf(t);
    }
}

// ---------------------------------------------------------------------------------------------- placeholder expansion / audio helpers (host)
// the i-th occurrence of `token_id` becomes counts[i] copies of it; occurrences beyond n_counts stay single
inline std::vector<uint32_t> expand_placeholders(const uint32_t* ids, size_t n, uint32_t token_id, const uint32_t* counts, size_t n_counts) {
    std::vector<uint32_t> out;
    out.reserve(n);
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) {
        if (ids[i] == token_id && k < n_counts) out.insert(out.end(), counts[k++], token_id);
        else out.push_back(ids[i]);
    }
    return out;
}

// ---------------------------------------------------------------------------------------------- video processor (host)
// video_smart_resize (/root/reference/src/utils/video_utils.rs:9-59).  The pixel products are u32 like the reference's (a release build wraps).
// video_ratio 0 = None; the reference passes Some(16): ffmpeg's scaler needs multiples of 16, so the factor becomes lcm(factor, 16).
inline void video_smart_resize(uint32_t num_frames, uint32_t height, uint32_t width, uint32_t temporal_factor, uint32_t factor, uint32_t min_pixels,
                               uint32_t max_pixels, uint32_t video_ratio, uint32_t& h_bar, uint32_t& w_bar) {
    AHA_REQUIRE(temporal_factor > 0 && factor > 0, "video_smart_resize: zero factor");
    if (num_frames < temporal_factor) throw std::runtime_error(std::to_string(num_frames) + " must be larger than temporal_factor " + std::to_string(temporal_factor));
    if (height < factor || width < factor)
        throw std::runtime_error("height:" + std::to_string(height) + " or width:" + std::to_string(width) + " must be larger than factor:" + std::to_string(factor));
    const uint32_t ratio = std::max(height, width) / std::min(height, width);
    if (ratio > 200) throw std::runtime_error("absolute aspect ratio mush be smaller than 200, got " + std::to_string(ratio));
    uint32_t image_factor = factor;
    if (video_ratio) {
        uint32_t a = image_factor, b = video_ratio;
        while (b) { const uint32_t t = a % b; a = b; b = t; }
        image_factor = image_factor / a * video_ratio;   // lcm
    }
    h_bar = round_by_factor((float)height, image_factor);
    w_bar = round_by_factor((float)width, image_factor);
    const uint32_t t_bar = round_by_factor((float)num_frames, temporal_factor);
    if (t_bar * h_bar * w_bar > max_pixels) {
        const float beta = std::sqrt((float)(num_frames * height * width) / (float)max_pixels);
        h_bar = std::max(image_factor, floor_by_factor((float)height / beta, image_factor));
        w_bar = std::max(image_factor, floor_by_factor((float)width / beta, image_factor));
    } else if (t_bar * h_bar * w_bar < min_pixels) {
        const float beta = std::sqrt((float)min_pixels / (float)(num_frames * height * width));
        h_bar = ceil_by_factor((float)height * beta, image_factor);
        w_bar = ceil_by_factor((float)width * beta, image_factor);
    }
}

// Frame sampling of get_video_data (qwen3vl/processor.rs:481-491, 526-527): `fps` frames per second of video, clamped to
// [min_frames, max_frames] and to the frame count; every sample_interval-th decoded frame is kept.  nframes is what video_smart_resize
// is called with -- the number of kept frames can differ from it (the reference has the same gap).
inline std::vector<uint32_t> video_sample_frames(uint32_t total_frames, uint32_t rate_num, uint32_t rate_den, uint32_t fps, uint32_t min_frames, uint32_t max_frames,
                                                 uint32_t& nframes) {
    AHA_REQUIRE(total_frames > 0 && rate_num > 0 && rate_den > 0, "No frames extracted from video");
    const float rate = (float)rate_num / (float)rate_den;
    nframes = (uint32_t)std::round((float)total_frames / rate * (float)fps);
    nframes = std::min(std::min(std::max(nframes, min_frames), max_frames), total_frames);
    AHA_REQUIRE(nframes > 0, "video frame sampling: max_frames is 0");
    const uint32_t interval = (uint32_t)std::round((float)total_frames / (float)nframes);
    std::vector<uint32_t> idx;
    for (uint32_t f = 0; f < total_frames; ++f) if (f % interval == 0) idx.push_back(f);
    return idx;
}

// calculate_timestamps (qwen3vl/processor.rs:282-307): frame indices padded with the last one to a multiple of t_merge_size, seconds = index / fps
// in f32, one stamp per group = the mean of its first and last frame's time
inline std::vector<float> video_timestamps(const uint32_t* frame_indices, size_t n, float fps, uint32_t t_merge_size) {
    AHA_REQUIRE(n > 0 && t_merge_size > 0, "calculate_timestamps: no frames");
    std::vector<uint32_t> idx(frame_indices, frame_indices + n);
    if (n % t_merge_size != 0) idx.insert(idx.end(), t_merge_size - n % t_merge_size, idx.back());
    std::vector<float> stamps;
    for (size_t i = 0; i < idx.size(); i += t_merge_size) stamps.push_back(((float)idx[i] / fps + (float)idx[i + t_merge_size - 1] / fps) / 2.0f);
    return stamps;
}

// The video half of process_info (qwen3vl/processor.rs:404-437) on token ids.  For every <|video_pad|> left in the prompt, in order of
// video index: t frame groups of [stamp text ids of "<{:.1} seconds>", <|vision_start|>, h*w/merge^2 x <|video_pad|>, <|vision_end|>] replace the
// FIRST <|vision_start|><|video_pad|><|vision_end|> triple if the prompt holds one anywhere, else the first lone <|video_pad|> (text.replacen(.., 1)).
// stamp_ids / stamp_lens: the tokenised timestamp strings of all frame groups of all videos, back to back (the tokenizer stays with the caller).
inline std::vector<uint32_t> expand_video_placeholders(const uint32_t* ids, size_t n, uint32_t video_tok, uint32_t vstart_tok, uint32_t vend_tok,
                                                       const uint32_t* grid_thw, size_t n_videos, uint32_t merge, const uint32_t* stamp_ids,
                                                       const uint32_t* stamp_lens, size_t n_stamps) {
    AHA_REQUIRE(merge > 0, "expand_video_placeholders: merge size 0");
    // expanded runs are kept apart from unexpanded tokens (the reference writes "<|placeholder|>" and renames it at the end)
    struct Tok { uint32_t id; bool done; };
    std::vector<Tok> text(n);
    for (size_t i = 0; i < n; ++i) text[i] = {ids[i], false};
    size_t index = 0, stamp = 0, stamp_off = 0;
    for (;;) {
        size_t lone = text.size(), triple = text.size();
        for (size_t i = 0; i < text.size(); ++i) {
            if (text[i].done || text[i].id != video_tok) continue;
            if (lone == text.size()) lone = i;
            if (i > 0 && i + 1 < text.size() && !text[i - 1].done && text[i - 1].id == vstart_tok && !text[i + 1].done && text[i + 1].id == vend_tok) { triple = i; break; }
        }
        if (lone == text.size()) break;
        if (index >= n_videos) throw std::runtime_error("more <|video_pad|> placeholders than video_grid_thw rows");
        const uint32_t t = grid_thw[3 * index], h = grid_thw[3 * index + 1], w = grid_thw[3 * index + 2];
        const uint32_t frame_seqlen = h * w / (merge * merge);
        std::vector<Tok> rep;
        for (uint32_t f = 0; f < t; ++f) {
            if (stamp >= n_stamps) throw std::runtime_error("fewer timestamp token runs than video frame groups");
            for (uint32_t k = 0; k < stamp_lens[stamp]; ++k) rep.push_back({stamp_ids[stamp_off + k], true});
            stamp_off += stamp_lens[stamp++];
            rep.push_back({vstart_tok, true});
            rep.insert(rep.end(), frame_seqlen, Tok{video_tok, true});
            rep.push_back({vend_tok, true});
        }
        size_t a = lone, b = lone + 1;
        if (triple != text.size()) { a = triple - 1; b = triple + 2; }
        text.erase(text.begin() + a, text.begin() + b);
        text.insert(text.begin() + a, rep.begin(), rep.end());
        ++index;
    }
    std::vector<uint32_t> out(text.size());
    for (size_t i = 0; i < text.size(); ++i) out[i] = text[i].id;
    return out;
}

inline size_t feat_extract_output_length(size_t audio_len) {
    const size_t leave = audio_len % 100;
    if (leave > 0) {
        const size_t feat = (leave - 1) / 2 + 1;
        return ((feat - 1) / 2 + 1 - 1) / 2 + 1 + (audio_len / 100) * 13;
    }
    return (audio_len / 100) * 13;
}

inline void float_range_normalize(float* t, size_t n) {
    float peak = 0.f;
    for (size_t i = 0; i < n; ++i) peak = std::max(peak, std::fabs(t[i]));
    if (peak == 0.0f) return;
    if (peak > 1.0f) { const float mul = (float)(1.0 / (double)peak); for (size_t i = 0; i < n; ++i) t[i] = t[i] * mul; }   // affine(1 / peak as f64, 0)
    for (size_t i = 0; i < n; ++i) t[i] = std::min(std::max(t[i], -1.0f), 1.0f);
}

// lengths of the chunks split_audio_into_chunks cuts (the last one is the remainder and may be 0, exactly like the reference)
inline std::vector<size_t> split_audio_into_chunks(size_t total_len, size_t sr, float max_chunk_sec) {
    const float total_sec = (float)total_len / (float)sr;
    if (total_sec <= max_chunk_sec) return {total_len};
    const size_t max_len = (size_t)std::round(max_chunk_sec * (float)sr);
    std::vector<size_t> splits(total_len / max_len, max_len);
    splits.push_back(total_len % max_len);
    return splits;
}

// ---------------------------------------------------------------------------------------------- sinc resampling
// get_sinc_resample_kernel, SincInterpHann branch: filter bank [new][2 * width + orig] in the reference's f32 arithmetic (affine = x * (f32)mul).
struct SincBank { std::vector<float> taps; int orig = 0, fresh = 0, width = 0, K = 0; };
inline SincBank sinc_resample_bank(int64_t orig_freq, int64_t new_freq, int64_t lowpass_filter_width = 6, double rolloff = 0.99) {
    if (orig_freq <= 0 || new_freq <= 0) throw std::runtime_error("Frequencies must be positive");
    if (lowpass_filter_width <= 0) throw std::runtime_error("Low pass filter width should be positive");
    int64_t a = orig_freq, b = new_freq;
    while (b) { const int64_t r = a % b; a = b; b = r; }
    SincBank B;
    B.orig = (int)(orig_freq / a); B.fresh = (int)(new_freq / a);
    const double base_freq = (double)std::min(B.orig, B.fresh) * rolloff;
    B.width = (int)std::ceil((double)lowpass_filter_width * (double)B.orig / base_freq);
    B.K = 2 * B.width + B.orig;
    B.taps.resize((size_t)B.fresh * B.K);
    const float inv_orig = (float)(1.0 / (double)B.orig), inv_new = (float)(1.0 / (double)B.fresh), bf = (float)base_freq;
    const float lw = (float)lowpass_filter_width, warg = (float)(M_PI / (double)lowpass_filter_width / 2.0), pi = (float)M_PI;
    const float scale = (float)(base_freq / (double)B.orig);
    for (int j = 0; j < B.fresh; ++j) {
        const float tj = (float)(-j) * inv_new + 0.0f;
        for (int k = 0; k < B.K; ++k) {
            const float idx = (float)(k - B.width) * inv_orig + 0.0f;
            float t = (tj + idx) * bf + 0.0f;
            t = std::min(std::max(t, -lw), lw);
            const float c = std::cos(t * warg);
            const float window = c * c;
            const float ts = t * pi;
            const float sinc = ts == 0.0f ? 1.0f : std::sin(ts) / ts;
            B.taps[(size_t)j * B.K + k] = sinc * window * scale;
        }
    }
    return B;
}
inline size_t sinc_resample_out_len(const SincBank& B, size_t length) {   // apply_sinc_resample_kernel: min(ceil(new * len / orig), conv frames * new)
    const size_t frames = length / B.orig + 1;   // (len + 2w + orig - K) / orig + 1
    const size_t target = (size_t)std::ceil((double)B.fresh * (double)length / (double)B.orig);
    return std::min(target, frames * (size_t)B.fresh);
}
// out[i * new + j] = sum_k taps[j][k] * padded[i * orig + k],  padded = [width zeros | wave | width + orig zeros]; taps in order, f32.
// SMEM: the whole bank [new][K] is staged in shared memory (48 -> 16 kHz: 15 KB).  Otherwise (44.1 -> 16 kHz: 160 x 477 taps = 305 KB; 22.05 -> 16 kHz:
// 600 KB) the taps are read through L1 / L2 from a TRANSPOSED copy [K][new], so that the threads of a warp (consecutive j) read consecutive words.
template <bool SMEM>
__global__ void sinc_resample_kernel(const float* __restrict__ wave, long long length, const float* __restrict__ taps, int orig, int fresh, int width, int K,
                                     float* __restrict__ out, long long n_out) {
    extern __shared__ float s_taps[];
    if (SMEM) {
        for (int i = threadIdx.x; i < fresh * K; i += blockDim.x) s_taps[i] = taps[i];
        __syncthreads();
    }
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    const long long i = o / fresh;
    const int j = (int)(o - i * fresh);
    const long long p0 = i * orig - width;   // wave index of tap 0
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
        const long long p = p0 + k;
        const float x = (p >= 0 && p < length) ? wave[p] : 0.f;
        const float t = SMEM ? s_taps[(size_t)j * K + k] : __ldg(taps + (size_t)k * fresh + j);
        acc = __fadd_rn(acc, __fmul_rn(x, t));   // separate multiply and add like a scalar f32 convolution (no fma contraction)
    }
    out[o] = acc;
}

}  // namespace aha
