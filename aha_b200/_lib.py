"""ctypes bindings of include/aha_b200.h.  No torch types cross this boundary."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libaha_b200.so")

AHA_F32, AHA_F16, AHA_BF16, AHA_U32, AHA_I64, AHA_U8 = range(6)
_NP2AHA = {np.dtype(np.float32): AHA_F32, np.dtype(np.float16): AHA_F16, np.dtype(np.uint32): AHA_U32,
           np.dtype(np.int64): AHA_I64, np.dtype(np.uint8): AHA_U8}


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_int32), ("rank", C.c_int32), ("shape", C.c_int64 * 8),
                ("data", C.c_void_p)]


class MM(C.Structure):
    _fields_ = [("data_vec", C.POINTER(TensorDesc)), ("n", C.c_size_t)]


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("tp_rank", C.c_int32), ("tp_world", C.c_int32), ("max_ctx", C.c_int32),
                ("max_prefill", C.c_int32), ("max_patches", C.c_int32), ("max_frames", C.c_int32),
                ("use_graph", C.c_int32), ("decode_impl", C.c_int32), ("gemm_impl", C.c_int32),
                ("tp_comm", C.c_void_p), ("reserved", C.c_int32 * 8)]


class GenParams(C.Structure):
    _fields_ = [("temperature", C.c_float), ("repeat_penalty", C.c_float), ("repeat_last_n", C.c_int32),
                ("max_tokens", C.c_uint32), ("seed", C.c_uint64), ("top_p", C.c_float), ("top_k", C.c_int32),
                ("flags", C.c_uint32), ("reserved", C.c_uint32)]


GEN_EOS_ON_FIRST, GEN_CONTINUE_RNG, GEN_REUSE_PREFIX = 1, 2, 4
TOKEN_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32)


class AsrChunk(C.Structure):
    _fields_ = [("ids", C.POINTER(C.c_uint32)), ("seq_len", C.c_size_t), ("input_features", TensorDesc)]


class BatchRequest(C.Structure):
    _fields_ = [("ids", C.POINTER(C.c_uint32)), ("seq_len", C.c_size_t), ("mm", C.POINTER(MM)), ("params", GenParams)]


class Usage(C.Structure):
    _fields_ = [("prompt_tokens", C.c_uint32), ("completion_tokens", C.c_uint32), ("prompt_secs", C.c_double),
                ("completion_secs", C.c_double), ("vision_secs", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("graph_launches", C.c_uint64),
                ("kernels_per_decode_step", C.c_uint64), ("weight_bytes", C.c_uint64),
                ("kv_bytes_per_token", C.c_uint64), ("decode_bytes_per_step_fixed", C.c_uint64)]


# every symbol include/aha_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_U32P = C.POINTER(C.c_uint32)
_F32P = C.POINTER(C.c_float)
SYMBOLS = {
    "aha_b200_abi_version": (C.c_int, []),
    "aha_b200_create": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(TensorDesc), C.c_size_t, _U32P, C.c_size_t,
                                  C.POINTER(Options), C.POINTER(_P)]),
    "aha_b200_forward_initial": (C.c_int, [_P, _U32P, C.c_size_t, C.c_size_t, C.POINTER(MM), _F32P, _U32P]),
    "aha_b200_forward_step": (C.c_int, [_P, _U32P, C.c_size_t, C.c_size_t, _F32P, _U32P]),
    "aha_b200_forward_extend": (C.c_int, [_P, _U32P, C.c_size_t, C.c_size_t, _F32P, _U32P]),
    "aha_b200_last_prefix_hit": (C.c_size_t, [_P]),
    "aha_b200_prefix_match": (C.c_size_t, [_U32P, C.c_size_t, _U32P, C.c_size_t, _U32P, C.c_size_t, C.c_int]),
    "aha_b200_mm_fingerprint": (C.c_uint64, [C.POINTER(MM)]),
    "aha_b200_clear_cache": (C.c_int, [_P]),
    "aha_b200_stop_token_ids": (C.c_size_t, [_P, _U32P, C.c_size_t]),
    "aha_b200_generate": (C.c_int, [_P, _U32P, C.c_size_t, C.POINTER(MM), C.POINTER(GenParams), _U32P, C.c_size_t,
                                    C.POINTER(C.c_size_t), C.POINTER(Usage)]),
    "aha_b200_generate_batch": (C.c_int, [_P, C.POINTER(BatchRequest), C.c_size_t, _U32P, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(Usage)]),
    "aha_b200_batch_open": (C.c_int, [_P]),
    "aha_b200_batch_add": (C.c_int, [_P, C.POINTER(BatchRequest), C.POINTER(C.c_int32), _U32P, C.POINTER(C.c_int32), C.POINTER(Usage)]),
    "aha_b200_batch_step": (C.c_int, [_P, _U32P, C.POINTER(C.c_int32), C.POINTER(C.c_size_t)]),
    "aha_b200_batch_close": (C.c_int, [_P]),
    "aha_b200_generate_stream": (C.c_int, [_P, _U32P, C.c_size_t, C.POINTER(MM), C.POINTER(GenParams), TOKEN_CALLBACK, C.c_void_p,
                                           C.POINTER(Usage)]),
    "aha_b200_asr_generate": (C.c_int, [_P, C.POINTER(AsrChunk), C.c_size_t, C.POINTER(GenParams), _U32P, C.c_size_t,
                                        C.POINTER(C.c_size_t), TOKEN_CALLBACK, C.c_void_p, C.POINTER(Usage)]),
    "aha_b200_debug_sample": (C.c_int, [_P, _F32P, C.POINTER(GenParams), _U32P, C.c_size_t, C.c_uint32, _U32P]),
    "aha_b200_mel_spectrogram": (C.c_int, [_P, _F32P, C.c_size_t, _F32P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "aha_b200_image_patchify": (C.c_int, [_P, C.POINTER(C.c_uint8), C.c_size_t, C.c_size_t, _F32P, C.c_size_t, _U32P]),
    "aha_b200_img_smart_resize": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _U32P, _U32P]),
    "aha_b200_image_resize": (C.c_int, [_P, C.POINTER(C.c_uint8), C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint8)]),
    "aha_b200_image_preprocess": (C.c_int, [_P, C.POINTER(C.c_uint8), C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32, _F32P, C.c_size_t, _U32P]),
    "aha_b200_expand_placeholders": (C.c_int, [_U32P, C.c_size_t, C.c_uint32, _U32P, C.c_size_t, _U32P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "aha_b200_video_smart_resize": (C.c_int, [C.c_uint32] * 8 + [_U32P, _U32P]),
    "aha_b200_video_sample_frames": (C.c_int, [C.c_uint32] * 6 + [_U32P, _U32P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "aha_b200_video_timestamps": (C.c_int, [_U32P, C.c_size_t, C.c_float, C.c_uint32, _F32P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "aha_b200_format_timestamp": (C.c_int, [C.c_float, C.c_char_p, C.c_size_t]),
    "aha_b200_video_preprocess": (C.c_int, [_P, C.POINTER(C.c_uint8), C.c_size_t, C.c_size_t, C.c_size_t, _F32P, C.c_size_t, _U32P]),
    "aha_b200_expand_video_placeholders": (C.c_int, [_U32P, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, _U32P, C.c_size_t, C.c_uint32, _U32P, _U32P,
                                                     C.c_size_t, _U32P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "aha_b200_feat_extract_output_length": (C.c_size_t, [C.c_size_t]),
    "aha_b200_float_range_normalize": (C.c_int, [_F32P, C.c_size_t]),
    "aha_b200_resample": (C.c_int, [_P, _F32P, C.c_size_t, C.c_int64, C.c_int64, _F32P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "aha_b200_sinc_resample_bank": (C.c_int, [C.c_int64, C.c_int64, _F32P, C.c_size_t, C.POINTER(C.c_int32)]),
    "aha_b200_split_audio_into_chunks": (C.c_int, [C.c_size_t, C.c_uint32, C.c_float, C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(C.c_size_t)]),
    "aha_b200_embed": (C.c_int, [_P, _U32P, C.c_size_t, _F32P]),
    "aha_b200_rerank": (C.c_int, [_P, _U32P, C.c_size_t, _U32P, C.POINTER(C.c_size_t), C.c_size_t, _F32P]),
    "aha_b200_nccl_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "aha_b200_rope_index": (C.c_int, [_U32P, C.c_size_t, _U32P, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "aha_b200_rope_index_mm": (C.c_int, [_U32P, C.c_size_t, _U32P, C.c_size_t, _U32P, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "aha_b200_destroy": (None, [_P]),
    "aha_b200_last_error": (C.c_char_p, [_P]),
    "aha_b200_stream": (_P, [_P]),
    "aha_b200_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "aha_b200_reset_stats": (C.c_int, [_P]),
    "aha_b200_set_trace": (C.c_int, [_P, C.c_int]),
    "aha_b200_debug_read": (C.c_int, [_P, C.c_char_p, C.c_int, _F32P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "aha_b200_decode_steps": (C.c_int, [_P, C.c_uint32, C.c_size_t, C.c_size_t, _U32P, C.POINTER(C.c_double)]),
    "aha_b200_debug_gemm": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _F32P, C.POINTER(C.c_uint16), _F32P, _F32P, _F32P,
                                      C.c_int, C.POINTER(C.c_double)]),
    "aha_b200_bench_kernel": (C.c_int, [_P, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
}

_lib = None


def load():
    """Load libaha_b200.so (built in-tree by aha_b200.build).  Fails loudly: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m aha_b200.build` (nvcc, sm_100a). "
                          "aha_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class BF16(np.ndarray):
    """uint16 array whose elements are bfloat16 bit patterns (numpy has no bf16 dtype): `arr.view(BF16)` marks it for make_desc."""


def to_bf16(x):
    """float32 -> bfloat16 bits (round to nearest even), tagged as BF16."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16).view(BF16)


def make_desc(arr, name=None):
    """TensorDesc for a C-contiguous numpy array (the caller keeps `arr` alive)."""
    if isinstance(arr, BF16):
        d = make_desc(arr.view(np.ndarray).view(np.float16), name)     # same bytes; only the dtype tag differs
        d.dtype = AHA_BF16
        return d
    if arr.dtype not in _NP2AHA:
        raise TypeError(f"unsupported dtype {arr.dtype}")
    if not arr.flags["C_CONTIGUOUS"]:
        raise ValueError("tensor must be C-contiguous")
    d = TensorDesc()
    d.name = name.encode() if name is not None else None
    d.dtype = _NP2AHA[arr.dtype]
    d.rank = arr.ndim
    for i, s in enumerate(arr.shape):
        d.shape[i] = s
    d.data = arr.ctypes.data
    return d
