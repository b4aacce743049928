import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` under gpurun)")


def has_gpu():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        cudart = None
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


GOLDEN = os.path.join(ROOT, "tests", "golden")
TOL = 1e-3  # north_star: logits within 1e-3 of the fp32 reference path


@pytest.fixture(scope="session")
def lib_built():
    from aha_b200 import build
    return build.build_lib()


def make_model(kind, preset, seed=0, **kw):
    """(config, weights, B200Model) -- GPU only."""
    from aha_b200 import B200Model, synth
    cfg = synth.get_config(kind, preset)
    w = synth.make_weights(kind, cfg, seed)
    tc = cfg if kind == "qwen3" else (cfg["text_config"] if kind == "qwen3vl" else cfg["thinker_config"]["text_config"])
    m = B200Model(kind, cfg, w, eos_ids=[tc["eos_token_id"]], **kw)
    return cfg, w, m


def make_oracle(kind, cfg, w):
    from oracle.qwen3 import Qwen3Model
    from oracle.qwen3vl import Qwen3VLModel
    from oracle.qwen3_asr import Qwen3ASRModel
    tc = cfg if kind == "qwen3" else (cfg["text_config"] if kind == "qwen3vl" else cfg["thinker_config"]["text_config"])
    cls = {"qwen3": Qwen3Model, "qwen3vl": Qwen3VLModel, "qwen3_asr": Qwen3ASRModel}[kind]
    return cls(cfg, w, [tc["eos_token_id"]])


def top2_gap(logits):
    l = np.asarray(logits).reshape(-1)
    i = np.argpartition(l, -2)[-2:]
    a, b = sorted(l[i])
    return float(b - a)
