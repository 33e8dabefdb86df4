"""Diagnostic (GPU): error of the temporal-attention backward vs fp32 autograd, next to the eager fp16 op sequence."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motionclone_b200 import ops
from oracle import mc_oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_kernels_gpu import _make_qkv, _to_oracle, _from_oracle, _ref_grads

dev = torch.device("cuda:0")
def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max()).item(), ((a - b).norm() / b.norm()).item()

for (L, H, DH, B, P) in [(16, 8, 40, 1, 64), (16, 8, 32, 1, 16), (8, 8, 40, 1, 64), (16, 8, 160, 1, 16)]:
    C = H * DH
    q, k, v = _make_qkv(B, L, P, C, seed=3, fused=True, dev=dev)
    scale = DH ** -0.5
    g = torch.Generator().manual_seed(9)
    d_o = torch.randn(B, L, P, C, generator=g).to(dev, torch.float16)
    idx = torch.randint(0, L, (B * P, H, L, 1), generator=g).to(dev, torch.uint8)
    d_g = (torch.randn(B * P, H, L, 1, generator=g) * 0.5).to(dev, torch.float16)
    for name, (a_do, a_dg) in {"o": (d_o, None), "gather": (None, d_g), "o+gather": (d_o, d_g)}.items():
        dq, dk, dv = ops.temporal_attention_backward(q, k, v, H, scale, a_do, None, idx if a_dg is not None else None, a_dg)
        gq, gk, gv = _ref_grads(q, k, v, H, scale, a_do, None, idx, a_dg)
        # eager fp16 autograd of the reference op sequence
        qh, kh, vh = (_to_oracle(t).contiguous().detach().requires_grad_(True) for t in (q, k, v))
        probs = O.temporal_probs(qh, kh, H, scale)
        probs2 = O.attention_probs(O.heads_to_batch(qh, H), O.heads_to_batch(kh, H), scale)
        out = O.batch_to_heads(torch.bmm(probs2, O.heads_to_batch(vh, H)), H)
        loss = 0
        if a_do is not None: loss = loss + (out * _to_oracle(a_do)).sum()
        if a_dg is not None: loss = loss + (torch.gather(probs, -1, idx.long()) * a_dg).sum()
        eq, ek, ev = torch.autograd.grad(loss, (qh, kh, vh), allow_unused=True)
        f = lambda t: _from_oracle(t, B, P)
        print(f"L={L} DH={DH} {name:9s} dq mine {rel(dq, gq)} eager {rel(f(eq), gq)} | dk mine {rel(dk, gk)} eager {rel(f(ek), gk)}"
              + (f" | dv mine {rel(dv, gv)} eager {rel(f(ev), gv)}" if a_do is not None else ""))
