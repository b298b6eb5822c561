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


def _plan_sum(plan, a):
    """Evaluate a float32 sum exactly the way the stitch kernels do from the plan table (eight strided
    accumulators per leaf, butterfly ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), trailing elements, then the tree)."""
    import numpy as np

    L, I, levels, root = (int(v) for v in plan[:4])
    lv = plan[4:4 + levels + 1]
    leaf_off = plan[4 + levels + 1:4 + levels + 1 + L + 1]
    pairs = plan[4 + levels + 1 + L + 1:].reshape(I, 2)
    val = np.zeros(L + I, dtype=np.float32)
    for i in range(L):
        blk = a[leaf_off[i]:leaf_off[i + 1]]
        n = blk.size
        m = n - n % 8
        r = blk[:m].reshape(-1, 8)
        acc = r[0].copy()
        for row in r[1:]:
            acc = (acc + row).astype(np.float32)
        for o in (1, 2, 4):
            acc = (acc + acc[np.arange(8) ^ o]).astype(np.float32)
        res = acc[0]
        for x in blk[m:]:
            res = np.float32(res + x)
        val[i] = res
    for h in range(1, levels + 1):
        for i in range(int(lv[h - 1]), int(lv[h])):
            val[L + i] = np.float32(val[pairs[i, 0]] + val[pairs[i, 1]])
    return val[root]


@pytest.mark.parametrize("H,W", [(1, 1), (2, 8), (4, 4), (3, 11), (7, 9), (30, 44), (40, 56), (64, 80), (256, 320)])
def test_stitch_plan_reproduces_numpy_sum_bit_exactly(H, W):
    """edv_op_stitch_plan (host-only) encodes numpy's pairwise-summation tree for the 8*H*W overlap elements:
    evaluating it the way the kernels do gives np.sum's float32 result bit for bit (utils/util.py:46-51)."""
    import numpy as np

    plan = engine.stitch_plan(H, W)
    n = 8 * H * W
    L, I, levels = int(plan[0]), int(plan[1]), int(plan[2])
    assert plan.size == 4 + levels + 1 + L + 1 + 2 * I and I == L - 1
    leaf_off = plan[4 + levels + 1:4 + levels + 1 + L + 1]
    sizes = np.diff(leaf_off)
    assert leaf_off[0] == 0 and leaf_off[-1] == n and sizes.min() >= 8 and sizes.max() <= 128
    rng = np.random.default_rng(H * 1000 + W)
    for trial in range(3):
        p = (rng.random((8 * H, W), dtype=np.float32) * 3 + 0.1).astype(np.float32)
        t = (rng.random((8 * H, W), dtype=np.float32) * 2).astype(np.float32)
        ones = np.ones_like(t)
        for arr in (ones * p * p, ones * p, ones * p * t, ones * t):
            assert _plan_sum(plan, arr.ravel()) == np.sum(arr)


def test_stitch_plan_rejects_oversized_overlap():
    lib = engine.load_library()
    assert lib.edv_op_stitch_plan(2048, 1024, None, 0) < 0      # 8*H*W = 2^24: np.sum(ones) no longer exact in float32
