"""Parameter-object plumbing shared by models, mesh generators, sources and
simulations.  Replaces the reference's dependency on the ``properties`` package
(casingSimulations/base.py:10-95) with a small declarative layer: a class lists
its parameters in ``_props`` (name -> default), ``setKwargs`` semantics reject
unknown keywords (SimPEG.utils.setKwargs, used at run.py:86, mesh.py:85,
model.py:136, sources.py:54), ``serialize``/``save`` write the same JSON shape
(``__class__`` key + one entry per parameter, base.py:63-89)."""
import copy as _copy
import json
import os
import warnings

import numpy as np

from .info import __version__

_REGISTRY = {}


def memo_property(fn):
    """Property evaluated once and kept under a '_'-prefixed name, i.e. among the lazily cached derived quantities that
    BaseCasing._invalidate drops whenever a public parameter is assigned.  (The reference recomputes its face masks and
    s_e on every access, sources.py:570, 698, 822 test `_srcList` instead of `_s_e`; on a 14.6 M-face mesh one evaluation
    costs seconds of numpy.)"""
    key = "_memo_" + fn.__name__

    def getter(self):
        d = self.__dict__
        if key not in d:
            object.__setattr__(self, key, fn(self))
        return d[key]
    getter.__name__ = fn.__name__
    getter.__doc__ = fn.__doc__
    return property(getter)


def load_properties(filename):
    """utils.py:11-24: rebuild an instance from its JSON file."""
    with open(filename, "r") as f:
        d = json.load(f)
    return BaseCasing.deserialize(d)


class BaseCasing(object):
    _props = {"filename": None, "directory": ".", "version": __version__}
    _array_props = ()

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        _REGISTRY[cls.__name__] = cls

    @classmethod
    def _all_props(cls):
        out = {}
        for k in reversed(cls.__mro__):
            out.update(getattr(k, "_props", {}))
        return out

    @classmethod
    def _all_array_props(cls):
        out = set()
        for k in cls.__mro__:
            out.update(getattr(k, "_array_props", ()))
        return out

    def _init_props(self, **kwargs):
        for name, default in self._all_props().items():
            if name not in self.__dict__:
                object.__setattr__(self, name, _copy.deepcopy(default))
        for key, val in kwargs.items():
            if key not in self._all_props() and not hasattr(type(self), key):
                raise Exception("{} attr is not recognized".format(key))
            setattr(self, key, val)

    def __init__(self, **kwargs):
        self._init_props(**kwargs)

    def __setattr__(self, name, value):
        if name in type(self)._all_array_props() and value is not None:
            value = np.atleast_1d(np.asarray(value, dtype=float))
        if isinstance(value, str) and name in getattr(type(self), "_instance_props", ()):
            value = load_properties(value)                      # LoadableInstance, base.py:10-18
        check = getattr(self, "_check_" + name, None)
        if check is not None:
            value = check(value)
        object.__setattr__(self, name, value)
        if not name.startswith("_"):
            self._invalidate(name)
            obs = getattr(self, "_observe_" + name, None)
            if obs is not None:
                obs(value)

    def _invalidate(self, name):
        """drop lazily cached derived quantities when a parameter changes"""
        for k in [k for k in self.__dict__ if k.startswith("_") and not k.startswith("__")]:
            if k not in ("_prob", "_survey", "_fields", "_physprops"):
                object.__delattr__(self, k)

    def _check_version(self, value):
        if value != __version__:
            warnings.warn("Instance was created on version {}, but the current version is {}".format(value, __version__))
        return value

    def validate(self):
        for k in type(self).__mro__:
            for name, fn in vars(k).items():
                if name.startswith("_validate_") and callable(fn):
                    fn(self)
        return True

    # ---- serialization ---------------------------------------------------
    def serialize(self):
        out = {"__class__": type(self).__name__}
        for name in self._all_props():
            val = getattr(self, name, None)
            if val is None:
                continue
            if isinstance(val, BaseCasing):
                val = val.serialize()
            elif isinstance(val, np.ndarray):
                val = val.tolist()
            elif isinstance(val, (list, tuple)) and val and isinstance(val[0], BaseCasing):
                val = [v.serialize() for v in val]
            out[name] = val
        return out

    @classmethod
    def deserialize(cls, d, trusted=True):
        d = dict(d)
        klass = _REGISTRY.get(d.pop("__class__", cls.__name__), cls)
        kw = {}
        for k, v in d.items():
            if isinstance(v, dict) and "__class__" in v:
                v = BaseCasing.deserialize(v)
            elif isinstance(v, list) and v and isinstance(v[0], dict) and "__class__" in v[0]:
                v = [BaseCasing.deserialize(x) for x in v]
            kw[k] = v
        with warnings.catch_warnings():
            warnings.simplefilter("default")
            return klass(**kw)

    def save(self, filename=None, directory=None):
        self.validate()
        filename = self.filename if filename is None else filename
        directory = self.directory if directory is None else directory
        if not os.path.isdir(directory):
            os.mkdir(directory)
        f = "/".join([directory, filename])
        with open(f, "w") as outfile:
            json.dump(self.serialize(), outfile)
        print("Saved {}".format(f))

    def copy(self):
        return type(self).deserialize(self.serialize())
