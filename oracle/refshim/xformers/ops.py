import torch


def memory_efficient_attention(q, k, v, attn_bias=None, op=None):
    """softmax(q k^T * dh^-0.5) v — published semantics of xformers.ops.memory_efficient_attention."""
    s = torch.einsum("bqd,bkd->bqk", q.float(), k.float()) * (q.shape[-1] ** -0.5)
    if attn_bias is not None:
        s = s + attn_bias
    return torch.einsum("bqk,bkd->bqd", s.softmax(-1), v.float()).to(q.dtype)
