"""ConfigMixin / register_to_config / FrozenDict restated (diffusers 0.11.1 configuration_utils)."""
import functools
import inspect
from collections import OrderedDict


class FrozenDict(OrderedDict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for k, v in self.items():
            object.__setattr__(self, k, v)
        object.__setattr__(self, "_frozen", True)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)


class ConfigMixin:
    config_name = None

    def register_to_config(self, **kwargs):
        kwargs.pop("kwargs", None)
        if not hasattr(self, "_internal_dict"):
            internal = dict(kwargs)
        else:
            internal = {**self._internal_dict, **kwargs}
        self._internal_dict = FrozenDict(internal)

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def inner_init(self, *args, **kwargs):
        init_kwargs = {k: v for k, v in kwargs.items() if not k.startswith("_")}
        init(self, *args, **init_kwargs)
        sig = inspect.signature(init)
        params = {n: p.default for i, (n, p) in enumerate(sig.parameters.items())
                  if i > 0 and p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)}
        new_kwargs = {}
        for arg, name in zip(args, params.keys()):
            new_kwargs[name] = arg
        for k, default in params.items():
            if k not in new_kwargs:
                new_kwargs[k] = init_kwargs.get(k, default)
        # extra **kwargs of the UNet (lora, SparseCausalAttention_index, least_sc_channel, ...)
        for k, v in init_kwargs.items():
            if k not in new_kwargs:
                new_kwargs[k] = v
        self.register_to_config(**new_kwargs)
    return inner_init
