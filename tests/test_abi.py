"""The C-ABI library loads and exports every symbol include/aha_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "aha_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aha_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_has_the_inference_model_seam():
    syms = header_symbols()
    for s in ("aha_b200_create", "aha_b200_forward_initial", "aha_b200_forward_step", "aha_b200_clear_cache",
              "aha_b200_stop_token_ids", "aha_b200_destroy", "aha_b200_last_error", "aha_b200_generate"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in include/aha_b200.h but not exported"


def test_python_bindings_cover_the_header(lib_built):
    from aha_b200 import _lib
    assert sorted(_lib.SYMBOLS) == header_symbols()
    lib = _lib.load()
    assert lib.aha_b200_abi_version() == 3


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "aha_b200.h")).read()
    assert "torch" not in text.lower() and "at::" not in text and "#include <cuda" not in text


def test_product_never_imports_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "aha_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(d, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_create_fails_loudly_without_gpu(lib_built):
    import pytest
    from conftest import has_gpu
    if has_gpu():
        pytest.skip("GPU present")
    from aha_b200 import B200Error, B200Model, synth
    cfg = synth.get_config("qwen3", "tiny")
    with pytest.raises(B200Error, match="no CPU fallback"):
        B200Model("qwen3", cfg, synth.make_weights("qwen3", cfg))
