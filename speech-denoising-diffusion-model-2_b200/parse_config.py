"""``ConfigParser`` — host mirror of reference parse_config.py:12-159 (JSON config -> objects by type name).

Kept surface: ``ConfigParser(config, resume, modification, run_id)``, ``from_args`` (-c/-r/-d + custom options),
``init_obj`` / ``init_ftn`` (registry lookup with ``getattr(module, cfg[name]['type'])``), ``__getitem__``,
``get_logger``, and the ``config`` / ``save_dir`` / ``log_dir`` / ``resume`` attributes.  The reference JSON files
(config_unet.json, ...) load unchanged.
"""
from __future__ import annotations

import json
import logging
import logging.handlers
import os
from argparse import ArgumentParser
from collections import OrderedDict
from datetime import datetime
from functools import partial
from pathlib import Path


def read_json(fname):
    with Path(fname).open("rt") as handle:
        return json.load(handle, object_hook=OrderedDict)


def write_json(content, fname):
    with Path(fname).open("wt") as handle:
        json.dump(content, handle, indent=4, sort_keys=False)


def setup_logging(save_dir, level=logging.INFO):
    """Console + rotating ``info.log`` under the run directory (the reference configures the same pair from
    logger/logger_config.json)."""
    root = logging.getLogger()
    root.setLevel(level)
    fmt = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    have = {type(h) for h in root.handlers}
    if logging.StreamHandler not in have:
        sh = logging.StreamHandler()
        sh.setFormatter(logging.Formatter("%(message)s"))
        root.addHandler(sh)
    fh = logging.handlers.RotatingFileHandler(str(Path(save_dir) / "info.log"), maxBytes=10485760, backupCount=20, encoding="utf8")
    fh.setFormatter(fmt)
    root.addHandler(fh)


class ConfigParser:
    _LEVELS = {0: logging.WARNING, 1: logging.INFO, 2: logging.DEBUG}

    def __init__(self, config, resume=None, modification=None, run_id=None):
        self._config = _apply_overrides(config, modification)
        self.resume = resume
        run_id = datetime.now().strftime(r"%m%d_%H%M%S") if run_id is None else run_id
        self._save_dir = Path(self.config["trainer"]["save_dir"]) / self.config["name"] / run_id
        self._save_dir.mkdir(parents=True, exist_ok=(run_id == ""))
        self._log_dir = self._save_dir
        write_json(self.config, self.save_dir / "config.json")
        setup_logging(self.log_dir)
        self.log_levels = dict(self._LEVELS)

    @classmethod
    def from_args(cls, args, options=""):
        for opt in options:
            args.add_argument(*opt.flags, default=None, type=opt.type)
        if isinstance(args, ArgumentParser):
            args = args.parse_args()
        if args.device is not None:
            os.environ["CUDA_VISIBLE_DEVICES"] = args.device
        if args.resume is not None:
            resume = Path(args.resume)
            cfg_fname = resume.parent / "config.json"
        else:
            assert args.config is not None, "Configuration file need to be specified. Add '-c config.json', for example."
            resume = None
            cfg_fname = Path(args.config)
        config = read_json(cfg_fname)
        if args.config and resume:
            config.update(read_json(args.config))      # fine-tuning: -c overrides the checkpoint's config
        modification = {opt.target: getattr(args, _opt_name(opt.flags)) for opt in options}
        return cls(config, resume, modification)

    def _type_and_args(self, name, kwargs):
        entry = self[name]
        module_args = dict(entry["args"])
        assert all(k not in module_args for k in kwargs), "Overwriting kwargs given in config file is not allowed"
        module_args.update(kwargs)
        return entry["type"], module_args

    def init_obj(self, name, module, *args, **kwargs):
        """``config.init_obj('name', module, a, b=1)`` == ``module.<config['name']['type']>(a, b=1, **config args)``."""
        type_name, module_args = self._type_and_args(name, kwargs)
        return getattr(module, type_name)(*args, **module_args)

    def init_ftn(self, name, module, *args, **kwargs):
        type_name, module_args = self._type_and_args(name, kwargs)
        return partial(getattr(module, type_name), *args, **module_args)

    def __getitem__(self, name):
        return self.config[name]

    def get_logger(self, name, verbosity=2):
        assert verbosity in self.log_levels, "verbosity option {} is invalid. Valid options are {}.".format(
            verbosity, self.log_levels.keys())
        logger = logging.getLogger(name)
        logger.setLevel(self.log_levels[verbosity])
        return logger

    @property
    def config(self):
        return self._config

    @property
    def save_dir(self):
        return self._save_dir

    @property
    def log_dir(self):
        return self._log_dir


def _apply_overrides(config, modification):
    """CLI overrides address nested keys as 'a;b;c' (reference parse_config.py:137-159)."""
    if modification is None:
        return config
    for path, value in modification.items():
        if value is None:
            continue
        *parents, leaf = path.split(";")
        node = config
        for key in parents:
            node = node[key]
        node[leaf] = value
    return config


def _opt_name(flags):
    for flag in flags:
        if flag.startswith("--"):
            return flag.replace("--", "")
    return flags[0].replace("--", "")
