"""Host-side processor routines of the path, bound to the C ABI (no GPU needed for these): the Python mirror of
Qwen3VLProcessor / Qwen3AsrProcessor helpers -- img_smart_resize (/root/reference/src/utils/img_utils.rs:297-331),
placeholder expansion (qwen3vl/processor.rs:386-399, qwen3_asr/processor.rs:93-97), get_feat_extract_output_lengths
(qwen3_asr/processor.rs:187-195), float_range_normalize (common/modules.rs:1353-1368), split_audio_into_chunks
(utils/audio_utils.rs:1743-1760).  The arithmetic lives in libaha_b200.so (csrc/preprocess.cuh)."""
import ctypes as C

import numpy as np

from . import _lib as L


class ProcessorError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise ProcessorError(L.load().aha_b200_last_error(None).decode())


def img_smart_resize(img_h, img_w, factor=32, min_pixels=65536, max_pixels=16777216):
    h, w = C.c_uint32(0), C.c_uint32(0)
    _check(L.load().aha_b200_img_smart_resize(int(img_h), int(img_w), int(factor), int(min_pixels), int(max_pixels), C.byref(h), C.byref(w)))
    return int(h.value), int(w.value)


def expand_placeholders(ids, token_id, counts):
    ids = np.ascontiguousarray(np.asarray(ids, np.uint32).reshape(-1))
    counts = np.ascontiguousarray(np.asarray(list(counts), np.uint32))
    n = C.c_size_t(0)
    lib = L.load()
    u32 = C.POINTER(C.c_uint32)
    _check(lib.aha_b200_expand_placeholders(ids.ctypes.data_as(u32), ids.size, int(token_id), counts.ctypes.data_as(u32), counts.size, None, 0, C.byref(n)))
    out = np.empty(n.value, np.uint32)
    _check(lib.aha_b200_expand_placeholders(ids.ctypes.data_as(u32), ids.size, int(token_id), counts.ctypes.data_as(u32), counts.size,
                                            out.ctypes.data_as(u32), out.size, C.byref(n)))
    return out


def feat_extract_output_length(n_frames):
    return int(L.load().aha_b200_feat_extract_output_length(int(n_frames)))


def float_range_normalize(wave):
    w = np.array(wave, np.float32, copy=True).reshape(-1)
    _check(L.load().aha_b200_float_range_normalize(w.ctypes.data_as(C.POINTER(C.c_float)), w.size))
    return w.reshape(np.shape(wave))


def sinc_resample_bank(orig_freq, new_freq):
    """get_sinc_resample_kernel (audio_utils.rs:66-151, Hann, width 6, rolloff 0.99) -> (taps [new/g, K], width)."""
    lib = L.load()
    dims = (C.c_int32 * 4)()
    _check(lib.aha_b200_sinc_resample_bank(int(orig_freq), int(new_freq), None, 0, dims))
    taps = np.empty((dims[0], dims[1]), np.float32)
    _check(lib.aha_b200_sinc_resample_bank(int(orig_freq), int(new_freq), taps.ctypes.data_as(C.POINTER(C.c_float)), taps.size, dims))
    return taps, int(dims[2])


def split_audio_into_chunks(total_len, sample_rate, max_chunk_sec):
    n = C.c_size_t(0)
    lib = L.load()
    _check(lib.aha_b200_split_audio_into_chunks(int(total_len), int(sample_rate), float(max_chunk_sec), None, 0, C.byref(n)))
    out = (C.c_size_t * max(n.value, 1))()
    _check(lib.aha_b200_split_audio_into_chunks(int(total_len), int(sample_rate), float(max_chunk_sec), out, n.value, C.byref(n)))
    return [int(out[i]) for i in range(n.value)]


# ---- video half of Qwen3VLProcessor (qwen3vl/processor.rs:253-307, 404-437, 447-571; utils/video_utils.rs:9-59) ----
def video_smart_resize(num_frames, height, width, temporal_factor=2, factor=32, min_pixels=4096, max_pixels=25165824, video_ratio=16):
    """-> (height, width) of the frames get_video_data asks the scaler for (video_ratio 16: lcm with the factor; 0 = None)."""
    h, w = C.c_uint32(0), C.c_uint32(0)
    _check(L.load().aha_b200_video_smart_resize(int(num_frames), int(height), int(width), int(temporal_factor), int(factor), int(min_pixels),
                                                int(max_pixels), int(video_ratio), C.byref(h), C.byref(w)))
    return int(h.value), int(w.value)


def video_sample_frames(total_frames, rate_num, rate_den=1, fps=2, min_frames=4, max_frames=768):
    """-> (nframes handed to video_smart_resize, kept frame indices) of get_video_data's sampling."""
    lib = L.load()
    nf, n = C.c_uint32(0), C.c_size_t(0)
    args = (int(total_frames), int(rate_num), int(rate_den), int(fps), int(min_frames), int(max_frames))
    _check(lib.aha_b200_video_sample_frames(*args, C.byref(nf), None, 0, C.byref(n)))
    out = np.empty(n.value, np.uint32)
    _check(lib.aha_b200_video_sample_frames(*args, C.byref(nf), out.ctypes.data_as(C.POINTER(C.c_uint32)), out.size, C.byref(n)))
    return int(nf.value), out


def video_timestamps(frame_indices, fps, t_merge_size=2):
    """calculate_timestamps: one f32 stamp (seconds) per group of t_merge_size sampled frames."""
    idx = np.ascontiguousarray(np.asarray(frame_indices, np.uint32).reshape(-1))
    lib = L.load()
    n = C.c_size_t(0)
    u32 = idx.ctypes.data_as(C.POINTER(C.c_uint32))
    _check(lib.aha_b200_video_timestamps(u32, idx.size, float(fps), int(t_merge_size), None, 0, C.byref(n)))
    out = np.empty(n.value, np.float32)
    _check(lib.aha_b200_video_timestamps(u32, idx.size, float(fps), int(t_merge_size), out.ctypes.data_as(C.POINTER(C.c_float)), out.size, C.byref(n)))
    return out


def format_timestamp(seconds):
    """format!("<{:.1} seconds>", t) -- the text in front of every frame group."""
    buf = C.create_string_buffer(64)
    _check(L.load().aha_b200_format_timestamp(float(seconds), buf, 64))
    return buf.value.decode()


def expand_video_placeholders(ids, video_grid_thw, stamp_token_runs, video_token_id, vision_start_token_id, vision_end_token_id, merge_size=2):
    """The <|video_pad|> expansion of process_info on token ids; stamp_token_runs: one list of token ids per frame group, videos in order."""
    ids = np.ascontiguousarray(np.asarray(ids, np.uint32).reshape(-1))
    grid = np.ascontiguousarray(np.asarray(video_grid_thw, np.uint32).reshape(-1, 3))
    lens = np.ascontiguousarray(np.asarray([len(r) for r in stamp_token_runs], np.uint32))
    flat = np.ascontiguousarray(np.asarray([t for r in stamp_token_runs for t in r], np.uint32))
    lib = L.load()
    u32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
    n = C.c_size_t(0)
    args = (u32(ids), ids.size, int(video_token_id), int(vision_start_token_id), int(vision_end_token_id), u32(grid), grid.shape[0], int(merge_size),
            u32(flat), u32(lens), lens.size)
    _check(lib.aha_b200_expand_video_placeholders(*args, None, 0, C.byref(n)))
    out = np.empty(n.value, np.uint32)
    _check(lib.aha_b200_expand_video_placeholders(*args, u32(out), out.size, C.byref(n)))
    return out
