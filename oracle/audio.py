"""Whisper-style log-mel frontend -- restates
/root/reference/src/models/feature_extractor/feature_extraction_whisper.rs:12-115,
src/utils/audio_utils.rs:1064-1083 (hann), :1158-1301 (mel bank), :1303-1347 (stft),
:1483-1503 (frames), src/utils/tensor_utils.rs:525-549 (reflect pad),
src/models/common/modules.rs:1256-1258 (log10), :1353-1368 (float_range_normalize)."""
import math
import numpy as np

F32 = np.float32


def create_hann_window(n):
    """audio_utils.rs:1064-1083: symmetric Hann, 0.5 + 0.5*cos(pi*i/(N-1)), i = 1-N, 3-N, ..., N-1 (f64 -> f32)."""
    if n == 1:
        return np.ones(1, F32)
    i = np.arange(1 - n, n, 2, dtype=np.float64)
    return (0.5 + 0.5 * np.cos(math.pi * i / (n - 1.0))).astype(F32)


def hertz_to_mel_slaney(f):
    f = F32(f)
    logstep = F32(27.0) / F32(np.log(F32(6.4)))
    if f >= F32(1000.0):
        return F32(15.0) + F32(np.log(f / F32(1000.0))) * logstep
    return F32(3.0) * f / F32(200.0)


def mel_to_hertz_slaney(m):
    m = F32(m)
    logstep = F32(np.log(F32(6.4))) / F32(27.0)
    if m >= F32(15.0):
        return F32(1000.0) * F32(np.exp(logstep * (m - F32(15.0))))
    return F32(200.0) * m / F32(3.0)


def _linspace(start, end, steps):
    step = (F32(end) - F32(start)) / F32(steps - 1)
    return (F32(start) + np.arange(steps, dtype=F32) * step).astype(F32)


def mel_filter_bank(num_frequency_bins, num_mel_filters, min_frequency, max_frequency, sampling_rate):
    """audio_utils.rs:1218-1301 with norm="slaney", MelScale::Slaney, triangles in Hz.
    Returns (num_frequency_bins, num_mel_filters)."""
    mel_min = hertz_to_mel_slaney(min_frequency)
    mel_max = hertz_to_mel_slaney(max_frequency)
    mel_freqs = _linspace(mel_min, mel_max, num_mel_filters + 2)
    filter_freqs = np.array([mel_to_hertz_slaney(m) for m in mel_freqs], dtype=F32)
    fft_freqs = _linspace(0.0, F32(sampling_rate) / F32(2.0), num_frequency_bins)
    # create_triangular_filter_bank :1195-1216
    diff = filter_freqs[1:] - filter_freqs[:-1]
    slopes = filter_freqs[None, :] - fft_freqs[:, None]
    down = (-slopes[:, :-2]) / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(np.minimum(down, up), F32(0)).astype(F32)
    enorm = (F32(2.0) / (filter_freqs[2:num_mel_filters + 2] - filter_freqs[:num_mel_filters])).astype(F32)
    return (fb * enorm[None, :]).astype(F32)


def pad_reflect_last_dim(t, pad_l, pad_r):
    """tensor_utils.rs:525-549.  NOTE the right pad is sliced from the ALREADY left-padded tensor
    using the ORIGINAL length (start = last_dim - pad_r), i.e. it mirrors
    orig[L-pad_r-pad_l : L-pad_l], not the true reflection -- reproduced as written."""
    last = t.shape[-1]
    if pad_l >= last or pad_r >= last:
        raise ValueError("pad must be less than last dim")
    out = t
    if pad_l > 0:
        out = np.concatenate([out[..., 1:1 + pad_l][..., ::-1], out], axis=-1)
    if pad_r > 0:
        start = last - pad_r
        out = np.concatenate([out, out[..., start:start + pad_r][..., ::-1]], axis=-1)
    return out


def float_range_normalize(t):
    """modules.rs:1353-1368."""
    peak = F32(np.max(np.abs(t.astype(F32))))
    if peak == 0:
        return t
    if peak > 1.0:
        t = (t.astype(F32) * F32(1.0 / np.float64(peak))).astype(F32)      # Tensor::affine(1 / peak as f64, 0): mul is converted to f32, then v * mul + add in f32
    return np.clip(t, -1.0, 1.0).astype(F32)


class WhisperFeatureExtractor:
    def __init__(self, feature_size=128, hop_length=160, n_fft=400, sampling_rate=16000):
        self.hop, self.n_fft, self.sr = hop_length, n_fft, sampling_rate
        self.window = create_hann_window(n_fft)
        self.mel_filters = mel_filter_bank(1 + n_fft // 2, feature_size, 0.0, 8000.0, sampling_rate).T.copy()

    def extract_fbank_features(self, waveform):
        """waveform (1, n) f32 -> (1, n_mels, n_frames-1).  feature_extraction_whisper.rs:93-115."""
        w = pad_reflect_last_dim(waveform.astype(F32), self.n_fft // 2, self.n_fft // 2)
        samples = w.shape[1]
        n_frames = 1 + (samples - self.n_fft) // self.hop
        idx = np.arange(n_frames)[:, None] * self.hop + np.arange(self.n_fft)[None, :]
        frames = (w[0][idx] * self.window[None, :]).astype(F32)
        spec = np.fft.rfft(frames.astype(np.float64), axis=-1)
        power = (spec.real.astype(F32) ** 2 + spec.imag.astype(F32) ** 2).astype(F32)  # norm_sqr
        mag = power.T[:, : n_frames - 1]  # drop last frame
        mel = np.matmul(self.mel_filters, mag).astype(F32)
        mel = np.maximum(mel, F32(1e-10))
        log10 = (np.log(mel) * F32(1.0 / math.log(10.0))).astype(F32)
        log10 = np.maximum(log10, log10.max() - F32(8.0))
        return (((log10 + F32(4.0)) * F32(0.25)).astype(F32))[None]

    def call(self, raw_speech, sampling_rate):
        if sampling_rate != self.sr:
            raise ValueError("sampling rate mismatch")
        return self.extract_fbank_features(raw_speech)


def split_audio_into_chunks(total_len, sr, max_chunk_sec):
    """audio_utils.rs:1743-1760 -> chunk lengths in samples.  The remainder is pushed even when it is 0 (as in the reference)."""
    total_sec = F32(total_len) / F32(sr)
    if total_sec <= F32(max_chunk_sec):
        return [total_len]
    q = F32(max_chunk_sec) * F32(sr)
    max_len = int(np.floor(q + F32(0.5)))           # f32::round on a positive value
    return [max_len] * (total_len // max_len) + [total_len % max_len]


def get_feat_extract_output_lengths(audio_len):
    """qwen3_asr/processor.rs:187-195."""
    leave = audio_len % 100
    if leave > 0:
        feat = (leave - 1) // 2 + 1
        return ((feat - 1) // 2 + 1 - 1) // 2 + 1 + (audio_len // 100) * 13
    return (audio_len // 100) * 13


# ----------------------------------------------------------------------------- sinc resampling
def get_sinc_resample_kernel(orig_freq, new_freq, gcd_val, lowpass_filter_width=6, rolloff=0.99):
    """audio_utils.rs:66-151 (SincInterpHann branch): the (new_freq/gcd, 2*width + orig_freq/gcd) filter bank and `width`.
    Candle's `affine(mul, add)` on an f32 tensor computes `x * (mul as f32) + (add as f32)`; every step below is f32."""
    if orig_freq <= 0 or new_freq <= 0:
        raise ValueError("Frequencies must be positive")
    if lowpass_filter_width <= 0:
        raise ValueError("Low pass filter width should be positive")
    orig, new = orig_freq // gcd_val, new_freq // gcd_val
    base_freq = float(min(orig, new)) * rolloff                                   # f64
    width = int(np.ceil(float(lowpass_filter_width) * float(orig) / base_freq))
    idx = np.arange(-width, width + orig, dtype=F32) * F32(1.0 / orig) + F32(0.0)
    t0 = np.arange(0, -new, -1, dtype=F32) * F32(1.0 / new) + F32(0.0)            # arange_step(0, -new, -1)
    t = ((t0[:, None] + idx[None, :]).astype(F32) * F32(base_freq) + F32(0.0)).astype(F32)
    t = np.clip(t, F32(-lowpass_filter_width), F32(lowpass_filter_width))
    window = np.cos((t * F32(np.pi / lowpass_filter_width / 2.0)).astype(F32)).astype(F32)
    window = (window * window).astype(F32)
    scale = base_freq / float(orig)
    ts = (t * F32(np.pi)).astype(F32)
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(ts == 0, F32(1.0), (np.sin(ts).astype(F32) / ts).astype(F32)).astype(F32)
    kernels = ((sinc * window).astype(F32) * F32(scale)).astype(F32)
    return kernels, width


def resample(wave, orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """audio_utils.rs:154-243: zero-pad (width, width + orig), strided conv1d with the filter bank, interleave the `new` phases, cut to
    ceil(new * len / orig).  wave: (channels, len) f32.  The f32 accumulation order of candle's conv1d is an assumption (taps in order)."""
    wave = np.asarray(wave, F32)
    if orig_freq <= 0 or new_freq <= 0:
        raise ValueError("Frequencies must be positive")
    if orig_freq == new_freq:
        return wave.copy()
    g = int(np.gcd(orig_freq, new_freq))
    kernels, width = get_sinc_resample_kernel(orig_freq, new_freq, g, lowpass_filter_width, rolloff)
    orig, new = orig_freq // g, new_freq // g
    ch, length = wave.shape
    padded = np.concatenate([np.zeros((ch, width), F32), wave, np.zeros((ch, width + orig), F32)], axis=1)
    K = kernels.shape[1]
    n_out = (padded.shape[1] - K) // orig + 1
    out = np.zeros((ch, n_out, new), F32)
    starts = np.arange(n_out) * orig
    for k in range(K):
        out = (out + padded[:, starts + k][:, :, None] * kernels[None, None, :, k]).astype(F32)
    flat = out.reshape(ch, n_out * new)
    target = int(np.ceil(float(new) * float(length) / float(orig)))
    return flat[:, :min(target, flat.shape[1])].copy()


def resample_simple(wave, orig_freq, new_freq):
    """audio_utils.rs:245-255."""
    return resample(wave, orig_freq, new_freq, 6, 0.99)
