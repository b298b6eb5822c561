"""Helpers shared by the golden-fixture tests (CPU and GPU)."""
import json
import os

import numpy as np

from oracle import weights

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ORACLE_KEYS = ("encoder", "features", "out_channels", "num_frames", "pe", "r", "lora_type",
               "residual_block_indexes", "temporal_lora", "disable_conv_head", "include_cls_token", "use_bn", "use_clstoken")


def manifest():
    with open(os.path.join(GOLDEN_DIR, "manifest.json")) as f:
        return json.load(f)


def load_case(name):
    m = manifest()[name]
    arrays = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    return m, arrays


def oracle_cfg(ctor):
    return weights.full_cfg({k: ctor[k] for k in ORACLE_KEYS if k in ctor})


def subsample_like_golden(name, s, arr):
    """make_golden.py stores a strided view of the two largest maps of one case."""
    if name in ("fwd_vits_lora_res_convhead", "fwd_vits_nocls_res"):
        if s == 0:
            return arr[:, :, ::4, ::4]
        if s == 1:
            return arr[:, :, ::2, ::2]
    return arr
