"""Config objects for the mirrored nn.Module constructors.

The reference builds its modules from OmegaConf `DictConfig`s (attribute access, `.get`, item
access; `OmegaConf.create` / `to_object`; `open_dict`).  When omegaconf is installed we use it;
otherwise (this image has no omegaconf and no network) a minimal stand-in with the same
semantics is used, so `MaxEntIRL(cfg)` accepts either.
"""
import contextlib

try:  # pragma: no cover - depends on the environment
    from omegaconf import DictConfig, ListConfig, OmegaConf, open_dict  # noqa: F401
    HAVE_OMEGACONF = True
except Exception:  # noqa: BLE001
    HAVE_OMEGACONF = False

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
            return _wrap(obj if obj is not None else {})

        @staticmethod
        def to_object(cfg):
            return _unwrap(cfg)

        @staticmethod
        def to_container(cfg, **_):
            return _unwrap(cfg)

    @contextlib.contextmanager
    def open_dict(cfg):
        yield cfg


def as_cfg(obj):
    """Accept plain dicts as well as DictConfigs."""
    if HAVE_OMEGACONF:
        return obj if isinstance(obj, (DictConfig, ListConfig)) else OmegaConf.create(obj)
    return OmegaConf.create(obj)
