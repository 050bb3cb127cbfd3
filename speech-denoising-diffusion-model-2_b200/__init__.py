"""sddm_b200 — B200-native (sm_100a) reverse-diffusion speech enhancement.

Drop-in for ONE hot path of yangye1098/Speech-Denoising-Diffusion-Model-2: ``SDDM.infer`` with the
``GaussianDiffusion`` schedule and the ``UNetModified2`` denoiser (config_unet.json), behind the reference's own
config / registry surface (``ConfigParser.init_obj``; type names ``GaussianDiffusion``, ``UNetModified2``, ``SDDM``;
reference ``state_dict`` key layout).  The device work is hand-written CUDA reached through a C ABI
(``include/sddm_b200.h``, ``csrc/libsddm_b200.so``); PyTorch only owns tensors and streams.  There is no CPU
fallback: every compute entry point raises if the extension or a CUDA device is missing.
"""
from . import _lib  # noqa: F401
from ._lib import PREC_BF16, PREC_BF16_ACT, PREC_BF16X3, PREC_FP32, SddmError, library_path  # noqa: F401

__all__ = ["_lib", "PREC_FP32", "PREC_BF16", "PREC_BF16_ACT", "PREC_BF16X3", "SddmError", "library_path"]
__version__ = "0.1.0"
