"""FeedForward(geglu) restated from diffusers 0.16.0 models/attention.py (from memory; KAT in tests).

net = [GEGLU(dim, 4*dim), Dropout, Linear(4*dim, dim)];  GEGLU: Linear(dim, 8*dim) -> chunk(2) -> h * gelu_erf(gate).
State-dict keys: net.0.proj.{weight,bias}, net.2.{weight,bias}.
"""
import torch
import torch.nn.functional as F
from torch import nn


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        assert activation_fn == "geglu", "only geglu is used by the reference"
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        self.net = nn.ModuleList([GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out)])

    def forward(self, hidden_states):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class AdaLayerNorm(nn.Module):  # imported by the reference, never instantiated (num_embeds_ada_norm=None)
    def __init__(self, *a, **k):
        raise NotImplementedError
