"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by the ``-m "not gpu"``
tests to pin ``oracle/endodav_oracle.py`` against the real reference when it is present.

The reference imports two packages that are absent from this image and are used only
for trivia:
  * ``easydict.EasyDict``        -- kwargs dict with attribute access
                                    (models/endodav/dpt_temporal.py:19,35-40)
  * ``fvcore.nn.weight_init``    -- ``c2_msra_fill`` in ResBottleneckBlock.__init__
                                    (models/backbones/layers/utils.py:8,134-135)
Both are replaced by in-memory stubs; nothing of the reference is copied.
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("ENDODAV_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "endodav", "endodav.py"))


def _install_stubs():
    import torch

    if "easydict" not in sys.modules:
        ed = types.ModuleType("easydict")

        class EasyDict(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError as e:  # pragma: no cover
                    raise AttributeError(k) from e

            def __setattr__(self, k, v):
                self[k] = v

        ed.EasyDict = EasyDict
        sys.modules["easydict"] = ed
    if "fvcore" not in sys.modules:
        fv = types.ModuleType("fvcore")
        fvnn = types.ModuleType("fvcore.nn")
        wi = types.ModuleType("fvcore.nn.weight_init")

        def c2_msra_fill(m):
            torch.nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                torch.nn.init.constant_(m.bias, 0)

        wi.c2_msra_fill = c2_msra_fill
        fvnn.weight_init = wi
        fv.nn = fvnn
        sys.modules["fvcore"] = fv
        sys.modules["fvcore.nn"] = fvnn
        sys.modules["fvcore.nn.weight_init"] = wi


def import_reference():
    """Return the reference ``models.endodav.endodav`` module (unmodified)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import importlib

        importlib.import_module("models.endodav.endodav")
    return sys.modules["models.endodav.endodav"]


def build_reference_model(cfg: dict):
    """Instantiate the reference ``endodav`` with constructor kwargs ``cfg``."""
    ref_mod = import_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ref_mod.endodav(**cfg)
    return m.eval()


def build_reference_endodac(kw: dict):
    """Instantiate the reference image model ``endodac`` (models/endodac/endodac.py:144) unmodified."""
    import importlib

    import_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("models.endodac.endodac")
        m = mod.endodac(**kw)
    return m.eval()
