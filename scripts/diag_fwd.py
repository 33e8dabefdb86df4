"""Diagnostic: spatial attention forward vs fp64 math on a few shapes; prints where the error sits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motionclone_b200 import ops

dev = torch.device("cuda:0")
for (B, H, N, dh, mul) in [(2, 8, 1024, 80, 2.0), (2, 8, 1024, 80, 1.0), (16, 8, 256, 160, 1.0), (1, 1, 256, 160, 1.0), (1, 1, 128, 160, 1.0),
                           (1, 1, 256, 64, 1.0), (1, 1, 256, 80, 1.0), (1, 1, 256, 32, 1.0), (1, 1, 256, 40, 1.0), (1, 1, 192, 160, 1.0)]:
    torch.manual_seed(1)
    C = H * dh
    q = torch.randn(B, N, C, device=dev, dtype=torch.float16) * mul
    k = torch.randn(B, N, C, device=dev, dtype=torch.float16)
    v = torch.randn(B, N, C, device=dev, dtype=torch.float16)
    scale = dh ** -0.5
    for rep in range(2):
        o, lse = ops.spatial_attention_forward(q, k, v, H, scale, want_lse=True)
        q4, k4, v4 = (t.double().view(B, N, H, dh).transpose(1, 2) for t in (q, k, v))
        s = torch.matmul(q4, k4.transpose(-1, -2)) * scale
        want = torch.matmul(torch.softmax(s, -1), v4).transpose(1, 2).reshape(B, N, C)
        wl = torch.logsumexp(s, -1)
        e = (o.double() - want).abs()
        el = (lse.double() - wl).abs()
        bad = (e > 1e-2).nonzero()
        msg = ""
        if len(bad):
            rows = bad[:, 1].unique()
            cols = (bad[:, 2] % dh).unique()
            msg = f" bad rows {rows[:6].tolist()}..{rows[-3:].tolist()} (n={len(rows)}) cols {cols[:8].tolist()}..(n={len(cols)}) frames {bad[:,0].unique().tolist()[:4]}"
        print(f"B={B} H={H} N={N} dh={dh} mul={mul} rep={rep}: o err {e.max().item():.3e} lse err {el.max().item():.3e} nan {torch.isnan(o).sum().item()}{msg}", flush=True)
