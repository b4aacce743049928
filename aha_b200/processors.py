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
