"""A minimal stand-in for the slice of OmegaConf 2.0 the inference entry point uses (main_scene_generation.py:23-25):
`OmegaConf.load(path)` -> a config whose nodes allow attribute and item access, assignment, `**` unpacking, and
read as None for missing keys (the reference depends on that: model.py:60).  omegaconf itself is not installed in
the target image; if it is importable it is used instead."""
import yaml


class ConfigNode(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            return None

    def __setattr__(self, k, v):
        self[k] = _wrap(v)

    def __missing__(self, k):
        return None

    def get(self, k, default=None):
        return self[k] if k in self else default

    def to_container(self):
        return _unwrap(self)


def _wrap(v):
    if isinstance(v, ConfigNode):
        return v
    if isinstance(v, dict):
        return ConfigNode({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return [_wrap(x) for x in v]
    return v


def _unwrap(v):
    if isinstance(v, dict):
        return {k: _unwrap(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_unwrap(x) for x in v]
    return v


class _OmegaConfShim:
    @staticmethod
    def load(path):
        with open(path) as f:
            return _wrap(yaml.safe_load(f))

    @staticmethod
    def create(obj=None):
        return _wrap(obj or {})

    @staticmethod
    def to_container(cfg, resolve=True):
        return _unwrap(cfg)


try:  # pragma: no cover - omegaconf is absent from the target image
    from omegaconf import OmegaConf  # type: ignore
except Exception:  # noqa: BLE001
    OmegaConf = _OmegaConfShim
