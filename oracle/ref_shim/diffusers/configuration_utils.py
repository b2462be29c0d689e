import functools
import inspect


class _Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        if not hasattr(self, "_internal_dict"):
            object.__setattr__(self, "_internal_dict", _Config())
        self._internal_dict.update(kwargs)
        # diffusers 0.13.x also exposes every ctor kwarg as an attribute (unet.in_channels works)
        for k, v in kwargs.items():
            if not hasattr(self, k):
                try:
                    object.__setattr__(self, k, v)
                except Exception:
                    pass

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        init(self, *args, **kwargs)
        sig = inspect.signature(init)
        params = [p for p in sig.parameters.values()][1:]
        cfg = {p.name: p.default for p in params}
        for p, a in zip(params, args):
            cfg[p.name] = a
        cfg.update(kwargs)
        self.register_to_config(**cfg)

    return inner
