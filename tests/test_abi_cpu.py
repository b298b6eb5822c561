"""The C-ABI boundary without a GPU: the shared library builds / loads, exports every entry point that
include/endodav_b200.h declares, the ctypes binding knows all of them, and context creation fails
loudly (never falls back) when no sm_100 device is present."""
import ctypes
import os
import re

import pytest
import torch

from endodav_b200 import build as edv_build
from endodav_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "endodav_b200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(edv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib_path = edv_build.build()           # no-op when the in-tree .so is up to date
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n
    assert sorted(engine.EXPORTS) == names, "ctypes binding and header disagree"


def test_create_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = engine.load_library()
    cfg = engine.EdvConfig()
    cfg.dim, cfg.depth, cfg.heads, cfg.features, cfg.num_frames = 384, 12, 6, 64, 32
    for i, v in enumerate((2, 5, 8, 11)):
        cfg.taps[i] = v
    for i, v in enumerate((48, 96, 192, 384)):
        cfg.out_channels[i] = v
    cfg.dtype = engine.EDV_F16
    ctx = ctypes.c_void_p(0)
    rc = lib.edv_create(ctypes.byref(cfg), ctypes.byref(ctx))
    assert rc == -5 and not ctx.value                       # EDV_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.edv_last_error(ctypes.c_void_p(0))


def test_bad_config_is_rejected_before_touching_the_device():
    lib = engine.load_library()
    cfg = engine.EdvConfig()
    cfg.dim, cfg.depth, cfg.heads = 100, 12, 6              # dim != heads * 64
    ctx = ctypes.c_void_p(0)
    assert lib.edv_create(ctypes.byref(cfg), ctypes.byref(ctx)) == -1   # EDV_ERR_ARG
