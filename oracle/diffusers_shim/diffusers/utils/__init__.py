"""diffusers.utils stand-ins (oracle shim, test infrastructure)."""
import logging as _logging
from collections import OrderedDict

WEIGHTS_NAME = "diffusion_pytorch_model.bin"


class BaseOutput(OrderedDict):
    def __post_init__(self):
        import dataclasses
        for f in dataclasses.fields(self):
            self[f.name] = getattr(self, f.name)


class logging:  # noqa: N801  (mirrors diffusers.utils.logging module API)
    @staticmethod
    def get_logger(name):
        return _logging.getLogger(name)


def is_accelerate_available():
    return False


def deprecate(*a, **k):
    return None
