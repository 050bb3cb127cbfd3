"""Denoiser registry (reference model/network.py:1-12): ``config.init_obj('network', module_network, ...)`` finds
the class by name.  The denoisers of config_unet.json (UNetModified2), config_diffwave.json (DiffWave) and config_wavegrad.json (WaveGrad) are provided."""
from .diffwave import DiffWave  # noqa: F401
from .unet_modified2 import UNetModified2  # noqa: F401
from .wavegrad import WaveGrad  # noqa: F401

__all__ = ["UNetModified2", "DiffWave", "WaveGrad"]
