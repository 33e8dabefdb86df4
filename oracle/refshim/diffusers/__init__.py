"""Minimal stand-in for diffusers==0.16.0 (names used by /root/reference only). Test infrastructure."""
from .schedulers import DDIMScheduler  # noqa: F401


class _Dummy:
    def __init__(self, *a, **k):
        raise NotImplementedError("refshim dummy: not on the guided-denoising path")


def __getattr__(name):  # StableDiffusionPipeline, DDIMInverseScheduler, ...
    if name.startswith("__"):
        raise AttributeError(name)
    return type(name, (_Dummy,), {})
