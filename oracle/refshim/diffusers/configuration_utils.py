"""ConfigMixin / register_to_config / FrozenDict — plumbing only (no arithmetic)."""
import functools
import inspect
from collections import OrderedDict


class FrozenDict(OrderedDict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for k, v in self.items():
            object.__setattr__(self, k, v)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)


class ConfigMixin:
    config_name = None

    def register_to_config(self, **kwargs):
        kwargs.pop("kwargs", None)
        prev = dict(getattr(self, "_internal_dict", {}))
        prev.update(kwargs)
        self._internal_dict = FrozenDict(prev)

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        cfg = dict(config)
        cfg.update(kwargs)
        sig = inspect.signature(cls.__init__).parameters
        return cls(**{k: v for k, v in cfg.items() if k in sig})


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = [p for n, p in sig.parameters.items() if n != "self"]
        new = {p.name: p.default for p in params if p.default is not inspect.Parameter.empty}
        for a, p in zip(args, params):
            new[p.name] = a
        new.update({k: v for k, v in kwargs.items() if not k.startswith("_")})
        self.register_to_config(**new)
        init(self, *args, **kwargs)

    return inner
