import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields


class BaseOutput(OrderedDict):
    """dataclass-style output that is also indexable by position and key (diffusers.utils.BaseOutput)."""

    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


class logging:  # noqa: N801 - mirrors `from diffusers.utils import logging`
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


def is_accelerate_available():
    return False
