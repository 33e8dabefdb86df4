import logging as _pylogging
from collections import OrderedDict


class BaseOutput(OrderedDict):
    """Dataclass-style output; attribute access only is what the reference uses (`.sample`)."""

    def __post_init__(self):
        pass


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


logging = _Logging()


def deprecate(*a, **k):
    pass


def is_accelerate_available():
    return False
