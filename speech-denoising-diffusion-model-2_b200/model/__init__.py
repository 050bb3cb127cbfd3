"""Registries mirroring the reference's ``model`` package: ``model.diffusion`` / ``model.network`` / ``model.model``
are looked up by ``ConfigParser.init_obj`` with plain ``getattr`` (reference parse_config.py:82-95)."""
from . import diffusion, metric, model, network  # noqa: F401
