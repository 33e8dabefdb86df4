"""ModelMixin — nn.Module with .dtype/.device (plumbing)."""
import torch
from torch import nn


class ModelMixin(nn.Module):
    _supports_gradient_checkpointing = False

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def enable_xformers_memory_efficient_attention(self, *a, **k):
        raise RuntimeError("xformers is not available in the shim (math attention path is used)")
