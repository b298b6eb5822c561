"""endodav_b200 -- B200-native (sm_100a) implementation of the EndoDAV video-depth forward path.

Public surface mirrors the reference: ``from endodav_b200 import endodav`` replaces
``from models.endodav.endodav import endodav`` and ``from endodav_b200 import endodac`` replaces
``from models.endodac.endodac import endodac``."""
from .model import endodav, endodac, parameter_layout  # noqa: F401
from .engine import EndoDAVError  # noqa: F401
from . import video  # noqa: F401
from . import metrics  # noqa: F401

__all__ = ["endodav", "endodac", "parameter_layout", "EndoDAVError", "video", "metrics"]
