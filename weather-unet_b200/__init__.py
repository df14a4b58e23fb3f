"""weather-unet_b200 — B200-native (sm_100a) drop-in for the cUNet generator hot path of
Sota0726/weather-Unet (cunet.py / nets.py / utils.py / ops.py surface).

The directory name carries a hyphen, so import it as ``weather_unet_b200`` (alias package at the
repo root), or put this directory on ``sys.path`` and use the reference's own module names:
``from cunet import Conditional_UNet``.
"""
from . import _lib  # noqa: F401
from .cunet import Conditional_UNet  # noqa: F401
from .utils import AdaIN  # noqa: F401
from .nets import r_double_conv  # noqa: F401

__all__ = ["Conditional_UNet", "AdaIN", "r_double_conv"]
