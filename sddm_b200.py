"""Import alias: ``import sddm_b200`` loads the package that lives in ``speech-denoising-diffusion-model-2_b200/``
(a directory name Python cannot import directly because of the hyphens)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "speech-denoising-diffusion-model-2_b200")
_spec = _ilu.spec_from_file_location("sddm_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["sddm_b200"] = _mod
_spec.loader.exec_module(_mod)
