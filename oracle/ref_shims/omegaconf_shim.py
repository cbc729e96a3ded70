"""Minimal omegaconf stand-in (test infrastructure; see oracle/ref_shims/__init__.py).

Only the semantics the reference's hot-path modules use: attribute + item access, `.get`,
OmegaConf.create / to_object / to_yaml, open_dict.
"""
import contextlib

import yaml


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class ListConfig(list):
    pass


def _wrap(o):
    if isinstance(o, dict):
        return DictConfig({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, (list, tuple)):
        return ListConfig([_wrap(v) for v in o])
    return o


def _unwrap(o):
    if isinstance(o, dict):
        return {k: _unwrap(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_unwrap(v) for v in o]
    return o


class OmegaConf:
    @staticmethod
    def create(obj=None):
        if isinstance(obj, str):
            obj = yaml.safe_load(obj)
        return _wrap(obj if obj is not None else {})

    @staticmethod
    def to_object(cfg):
        return _unwrap(cfg)

    @staticmethod
    def to_container(cfg, resolve=True, **_):
        return _unwrap(cfg)

    @staticmethod
    def to_yaml(cfg, **_):
        return yaml.safe_dump(_unwrap(cfg))

    @staticmethod
    def load(path):
        with open(path) as f:
            return _wrap(yaml.safe_load(f))

    @staticmethod
    def set_struct(cfg, flag):
        return None


@contextlib.contextmanager
def open_dict(cfg):
    yield cfg
