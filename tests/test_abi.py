"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/serenc.h declares, and its
pure-host helpers agree with the reference's integer arithmetic. No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from interspeech_ser_b200 import _lib, configs
from interspeech_ser_b200.build import build_library

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build_library()
    return _lib.load_library(build_if_missing=False)


def header_symbols():
    with open(os.path.join(REPO, "include", "serenc.h")) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(serenc_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for sym in header_symbols():
        assert hasattr(lib, sym), sym


def test_config_struct_layout():
    # serenc_config: 14 int32 + 1 float + 8 reserved int32
    assert C.sizeof(_lib.SerencConfig) == 4 * (14 + 1 + 8)


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.serenc_version()
    assert isinstance(lib.serenc_last_error(), bytes)


@pytest.mark.parametrize("n", [0, 1, 399, 400, 401, 719, 720, 4001, 17777, 32000, 64000, 192000, 320000])
def test_num_frames_matches_formula(lib, n):
    # HF _get_feat_extract_output_lengths (modeling_wavlm.py:640-659); SURVEY: 64000 -> 199, equals floor((L-400)/320)+1
    expect = configs.w2v_num_frames(n)
    assert lib.serenc_w2v_num_frames(n) == expect
    if n >= 400:
        assert expect == (n - 400) // 320 + 1
    else:
        assert expect == 0


def test_wavlm_bucket_matches_hf_golden(lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "wavlm_buckets.npz"))
    mine = np.array([lib.serenc_wavlm_bucket(int(d), 320, 800) for d in g["delta"]], dtype=np.int32)
    assert np.array_equal(mine, g["bucket"])
    kat = {-1: 1, 1: 161, 79: 239, 80: 240, 81: 240, 100: 247, 200: 271, 400: 295, -400: 135, 778: 319, -778: 159, 5000: 319}
    for d, b in kat.items():
        assert lib.serenc_wavlm_bucket(d, 320, 800) == b


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = _lib.SerencConfig()
    cfg.arch, cfg.hidden, cfg.layers, cfg.heads, cfg.ffn = 0, 128, 2, 2, 256
    cfg.conv_dim, cfg.pos_conv_kernel, cfg.pos_conv_groups = 512, 16, 4
    h = C.c_void_p()
    st = lib.serenc_create(C.byref(cfg), 0, C.byref(h))
    assert st == -5  # SERENC_ERR_NO_DEVICE: no CPU fallback exists
    assert b"no CPU fallback" in lib.serenc_last_error() or b"sm_" in lib.serenc_last_error()


def test_engine_refuses_cpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from interspeech_ser_b200.engine import Engine
    from interspeech_ser_b200.weights import random_init
    cfg = configs.get_config("tiny/wavlm")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(cfg, random_init(cfg, 0), 0)
