"""Known-answer tests for the oracle's scheduler / schedule / encoding pieces and for the pieces restated from
diffusers 0.16.0 (SURVEY.md §8c lists the expected values; they were computed from the reference's formulas)."""
import math

import numpy as np
import pytest
import torch

from oracle import mc_oracle as O


def test_alphas_cumprod_kat():
    acp = O.alphas_cumprod()
    want = {0: 0.999149978, 1: 0.998289526, 400: 0.289899617, 699: 0.035603896, 700: 0.035295468, 999: 0.001578963}
    for t, v in want.items():
        assert abs(acp[t].item() - v) < 2e-7 * max(1.0, v / 1e-3), (t, acp[t].item())
    assert acp.dtype == torch.float32 and acp.shape == (1000,)


@pytest.mark.parametrize("S,G,gs,head,boundary,tail", [
    (100, 50, 0.3, [999, 993, 987, 981], (700, 699), [29, 14, 0]),      # configs/t2v_camera.yaml
    (300, 180, 0.4, [999, 997, 995, 992], (600, 599), [10, 5, 0]),      # configs/t2v_object.yaml
    (100, 40, 0.3, [999, 991, 984, 976], (700, 699), [24, 12, 0]),      # configs/i2v_rgb.yaml
    (200, 120, 0.4, [999, 996, 992, 989], (600, 599), [15, 8, 0]),      # configs/i2v_sketch.yaml
    (10, 5, 0.3, [999, 924, 850, 775], (700, 699), [350, 175, 0]),      # plumbing config (BASELINE configs[0])
    (50, 25, 0.3, [999, 987, 974, 962], (700, 699), [58, 29, 0]),       # bench mapping of t2v_camera
])
def test_uneven_timesteps_kat(S, G, gs, head, boundary, tail):
    ts = list(O.uneven_timesteps(S, G, gs))
    assert len(ts) == S and len(set(ts)) == S
    assert ts[:4] == head and ts[-3:] == tail
    assert (ts[G - 1], ts[G]) == boundary
    assert all(a > b for a, b in zip(ts, ts[1:]))


def test_warm_cool_multipliers():
    # G=50, warm=cool=10: steps 0-9 -> 0.1..1.0, 10-40 -> 1.0, 41-49 -> 0.9..0.1 (strict '>' at motionclone_functions.py:232)
    s = [O.loss_scale(i, 50, 10, 10) for i in range(50)]
    assert np.allclose(s[:10], [(i + 1) / 10 for i in range(10)])
    assert all(v == 1.0 for v in s[10:41])
    assert np.allclose(s[41:], [(50 - i) / 10 for i in range(41, 50)])
    # short schedules: both multipliers apply at once (G=5 < warm_up): step 0 -> 0.1 * 1.0, step 4 -> 0.5 * 0.1
    assert math.isclose(O.loss_scale(4, 5, 10, 10), 0.5 * 0.1)


def test_positional_encoding_formula():
    pe = O.positional_encoding(64, 32)
    assert pe.shape == (1, 32, 64)
    for p, i in [(0, 0), (3, 0), (7, 5), (31, 31)]:
        w = 10000 ** (-(2 * i) / 64)
        assert abs(pe[0, p, 2 * i].item() - math.sin(p * w)) < 1e-5
        assert abs(pe[0, p, 2 * i + 1].item() - math.cos(p * w)) < 1e-5


def test_timestep_embedding_flip_sin_to_cos():
    e = O.timestep_embedding(torch.tensor([0, 500]), 320)
    assert e.shape == (2, 320) and e.dtype == torch.float32
    assert torch.allclose(e[0, :160], torch.ones(160)) and torch.allclose(e[0, 160:], torch.zeros(160))  # [cos | sin]
    f1 = math.exp(-math.log(10000) * 1 / 160)
    assert abs(e[1, 1].item() - math.cos(500 * f1)) < 1e-5 and abs(e[1, 161].item() - math.sin(500 * f1)) < 1e-5


def test_dtype_rules_the_kernels_rely_on():
    x = torch.randn(8).half()
    assert (torch.tensor(0.5) * x).dtype == torch.float16  # fp32 0-dim x fp16 tensor -> fp16
    assert torch.nn.functional.mse_loss(x, x.flip(0)).dtype == torch.float16
    p = torch.tensor([[0.25, 0.5, 0.5, 0.1]])
    assert torch.topk(p, 1).indices.item() == 1 and p.argmax(-1).item() == 1  # ties: lowest index
    assert O.top1_lowest_index(p)[1].item() == 1


def test_cfg_ddim_fp16_sequence_matches_float_math():
    g = torch.Generator().manual_seed(0)
    ec, eu, x, sc = (torch.randn(1, 4, 8, 8, 8, generator=g).half() for _ in range(4))
    acp = O.alphas_cumprod()
    ts = O.uneven_timesteps(50, 25, 0.3)
    for step in (0, 24, 25, 49):
        a_t, a_p = O.ddim_scalars(acp, ts, step)
        got = O.cfg_ddim_step_fp16_sequence(ec, eu, x, sc, 7.5, a_t, a_p).float()
        want = O.ddim_guided_step(O.cfg_combine(ec.float(), eu.float(), 7.5), x.float(), sc.float(), a_t, a_p)
        # eps ~ 8.5 x N(0,1) reaches |e| ~ 30 (fp16 spacing 1.6e-2), then x 1/sqrt(a_t) (5.3 at t=700): ~1 % of max
        assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item()
    a_t, a_p = O.ddim_scalars(acp, ts, 49)
    assert a_p.item() == 1.0  # last step: final_alpha_cumprod -> the direction term vanishes


def test_motion_loss_closed_form_gradient():
    g = torch.Generator().manual_seed(1)
    s = torch.randn(6, 8, 16, 16, generator=g, dtype=torch.float64, requires_grad=True)
    idx = torch.randint(0, 16, (6, 8, 16, 1), generator=g).to(torch.uint8)
    ref = torch.rand(6, 8, 16, 1, generator=g, dtype=torch.float64)
    p = s.softmax(-1)
    loss = 2000.0 * O.motion_loss({"m": p}, {"m": [ref, idx]})
    (ds,) = torch.autograd.grad(loss, s)
    closed = O.motion_loss_dscores_closed_form(p.detach(), idx, ref, 2000.0)
    assert torch.allclose(ds, closed, atol=1e-12)
