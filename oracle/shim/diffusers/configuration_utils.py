"""register_to_config / ConfigMixin / FrozenDict: records the merged constructor kwargs BEFORE running __init__
(the reference reads self.config inside __init__, cogvideox_control.py:629; SURVEY probe P9-ii)."""
import functools
import inspect
import json
import os
from collections import OrderedDict


class FrozenDict(OrderedDict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for k, v in self.items():
            object.__setattr__(self, k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        kwargs.pop("kwargs", None)
        cur = dict(getattr(self, "_internal_dict", {}))
        cur.update(kwargs)
        object.__setattr__(self, "_internal_dict", FrozenDict(cur))

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        cfg = {k: v for k, v in dict(config).items() if not k.startswith("_")}
        cfg.update(kwargs)
        sig = inspect.signature(cls.__init__).parameters
        if not any(p.kind == inspect.Parameter.VAR_KEYWORD for p in sig.values()):
            cfg = {k: v for k, v in cfg.items() if k in sig}
        return cls(**cfg)

    @classmethod
    def load_config(cls, path, subfolder=None, **kwargs):
        d = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(d, cls.config_name)) as f:
            return json.load(f)


def register_to_config(init):
    @functools.wraps(init)
    def inner_init(self, *args, **kwargs):
        init_kwargs = {k: v for k, v in kwargs.items() if not k.startswith("_")}
        sig = inspect.signature(init)
        params = {n: p.default for i, (n, p) in enumerate(sig.parameters.items())
                  if i > 0 and p.kind not in (inspect.Parameter.VAR_KEYWORD, inspect.Parameter.VAR_POSITIONAL)}
        new = {}
        for arg, name in zip(args, params.keys()):
            new[name] = arg
        new.update({k: init_kwargs.get(k, default) for k, default in params.items() if k not in new})
        new.update({k: v for k, v in init_kwargs.items() if k not in new})
        self.register_to_config(**new)
        init(self, *args, **init_kwargs)

    return inner_init
