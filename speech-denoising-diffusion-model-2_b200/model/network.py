"""Denoiser registry (reference model/network.py:1-12): ``config.init_obj('network', module_network, ...)`` finds
the class by name.  Only the denoiser on the hot path of config_unet.json is provided."""
from .unet_modified2 import UNetModified2  # noqa: F401

__all__ = ["UNetModified2"]
