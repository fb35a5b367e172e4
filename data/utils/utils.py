"""Boundary shim for `from data.utils.utils import *` in the reference's main_scene_generation.py:6, which relies on
that star-import for the names `torch`, `OmegaConf` and `instantiate_from_config` (data/utils/utils.py:75-81,178-181).
The rest of the reference's module (Lightning callbacks, data modules) is training-only and out of scope."""
import importlib

import torch  # noqa: F401  (re-exported)

from sgam_neurips22_b200.config import OmegaConf  # noqa: F401  (re-exported)


def get_obj_from_str(string, reload=False):
    module, cls = string.rsplit(".", 1)
    if reload:
        importlib.reload(importlib.import_module(module))
    return getattr(importlib.import_module(module, package=None), cls)


def instantiate_from_config(config):
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))
