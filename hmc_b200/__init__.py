"""hmc_b200 -- B200-native HaploBuilder / EM E-step path of HMC."""
from .device import DeviceBuilder, HmcError, FLAG_SCALED, FLAG_KEEP_TRELLIS, FLAG_PANEL_ROWS  # noqa: F401

from .host import HaploModel  # noqa: F401,E402

__all__ = ["HaploModel", "DeviceBuilder", "HmcError", "FLAG_SCALED", "FLAG_KEEP_TRELLIS", "FLAG_PANEL_ROWS"]
