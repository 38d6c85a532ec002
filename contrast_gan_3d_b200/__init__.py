"""B200-native hot path of contrast-gan-3D: drop-in ResnetGenerator / PatchGANDiscriminator /
Trainer / CCTAContrastCorrector whose numerical work runs in libcgan3d.so (hand-written sm_100a
CUDA behind a C ABI, see include/cgan3d.h)."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
