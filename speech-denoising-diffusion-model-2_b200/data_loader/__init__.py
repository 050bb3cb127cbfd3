from . import data_loaders  # noqa: F401
