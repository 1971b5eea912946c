import importlib

from .. import error

_REGISTRY = {}


def register(id, entry_point=None, **kwargs):
    _REGISTRY[id] = (entry_point, kwargs.get("kwargs", {}))


def make(id, **kwargs):
    if ":" in id:
        module, id = id.split(":", 1)
        importlib.import_module(module)
    if id not in _REGISTRY:
        raise error.UnregisteredEnv("No registered env with id: {}".format(id))
    entry_point, defaults = _REGISTRY[id]
    if callable(entry_point):
        cls = entry_point
    else:
        mod_name, attr = entry_point.split(":")
        cls = getattr(importlib.import_module(mod_name), attr)
    kw = dict(defaults)
    kw.update(kwargs)
    return cls(**kw)
