"""Parity AT THE BENCH CONFIGURATION (BASELINE.json configs[1]: 16 x 512 x 512, SD1.5 + v3_sd15_mm widths), at the
32-frame long-clip shape (configs[4]) and with SparseCtrl at SD1.5 width (configs[3]): one guided and one plain DDIM
step through this package's `pipe.single_step_video` against `oracle.single_step` = the reference's op sequence
(baddbmm / softmax / bmm attention everywhere, eager elementwise ops) in fp16 ON THE SAME DEVICE, same weights, same
inputs, same motion representation. Full 50-step CPU/fp32 reference runs at this size take hours (2 PFLOP), so the
size-independent statement is per-step: the two implementations start from identical latents and must produce the same
x_{t-1} up to fp16 arithmetic.

Bars (absolute, with the fp16 spacing at the tensor's magnitude next to them):
  * guided step and plain step: max |x_ours - x_oracle| <= 4 ulp(max |x|) and mean |diff| <= 0.5 ulp(max |x|)
    (measured on B200 at 16 x 512 x 512: max 2 ulp, mean 0.27 ulp - two independently rounded fp16 results);
  * guidance gradient: cosine >= 0.995, max-abs error <= 8 % of max |g| (measured: cosine 0.9990, 4.4 %; |g| <= 2.4e-3
    is accumulated in fp16 through the backward of 60 % of the UNet by two different kernel sets, each ~2-3 % from the
    fp32 gradient - tests/test_pipeline_gpu.py holds both against the fp32 reference gradient at the fixture sizes);
  * extraction: top-1 index sets of all six guided modules equal the oracle's except on rows that are near-ties in the
    ORACLE's own probabilities (top-2 gap <= 2 fp16 ulps of the probability), and those are < 1 % of rows.
(The north-star's absolute 1e-3 on final latents is below half an fp16 ulp once |x| >= 2; DESIGN.md §2 has the arithmetic:
the DDIM recursion amplifies x_T by 1/sqrt(alpha_bar_999) = 25.2 with a random-init UNet, so |x| reaches ~1e2.)
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mc_oracle as O  # noqa: E402
from motionclone_b200.synthetic import (SPARSECTRL_LATENT_KWARGS, UNET_SD15_CONFIG, synthetic_condition,  # noqa: E402
                                        synthetic_inputs, synthetic_state_dict)

BASE = dict(cfg_scale=7.5, negative_prompt="", warm_up_steps=10, cool_up_steps=10, motion_guidance_weight=2000,
            motion_guidance_blocks=["up_blocks.1"], add_noise_step=400, inference_steps=50, guidance_steps=30,
            guidance_scale=0.4, height=512, width=512, new_prompt="synthetic")
CASES = {
    "c2_t2v_16x512x512": dict(BASE, video_length=16),
    "c5_longclip_32x512x512": dict(BASE, video_length=32),
    "c4_sparsectrl_latent_16x512x512": dict(BASE, video_length=16, image_index=[0], controlnet_scale=1.0,
                                            sparsectrl="latent", guidance_steps=20, guidance_scale=0.3),
}


def _ulp(x: float) -> float:
    """fp16 spacing at magnitude x."""
    return 2.0 ** (math.floor(math.log2(max(x, 2.0 ** -14))) - 10)


@pytest.fixture(scope="module", params=list(CASES))
def bench_case(request):
    import motionclone_b200 as mc
    dev = torch.device("cuda:0")
    icfg = dict(CASES[request.param])
    L = icfg["video_length"]
    inp = synthetic_inputs(L, 512, 512, 768, 42)
    h = lambda t: t.to(dev, torch.float16)  # noqa: E731
    icfg.update(video_latents=inp["clip_latents"].half(), video_noise=inp["clip_noise"].half())
    cn_kwargs = cn = None
    if icfg.get("sparsectrl"):
        cn_kwargs = SPARSECTRL_LATENT_KWARGS
        cond = synthetic_condition("latent", 1, 512, 512, L, 47)
        images = cond["cond_latents"].permute(1, 0, 2, 3).unsqueeze(0)
        icfg.update(controlnet_images=images.half())
    pipe = mc.build_pipeline(UNET_SD15_CONFIG, icfg, device=dev, weight_seed=42, controlnet_kwargs=cn_kwargs)
    pipe.set_prompt_embeds(h(inp["text_embeddings"]))
    shapes = {k: v.shape for k, v in pipe.unet.state_dict().items()}
    sd = {k: h(v) for k, v in synthetic_state_dict(shapes, 42).items()}
    if cn_kwargs is not None:
        cshapes = {k: v.shape for k, v in pipe.controlnet.state_dict().items()}
        cn = dict(sd={k: h(v) for k, v in synthetic_state_dict(cshapes, 43).items()}, kwargs=cn_kwargs,
                  image_index=icfg["image_index"], scale=icfg["controlnet_scale"], images=h(images))
    use_cn = cn is not None
    rep = pipe.obtain_motion_representation(motion_representation_path=None, use_controlnet=use_cn)
    with torch.no_grad():
        rep_o, probs_o = O.obtain_motion_representation(sd, UNET_SD15_CONFIG, h(inp["clip_latents"]), h(inp["clip_noise"]),
                                                        h(inp["text_embeddings"][[0]]), icfg["add_noise_step"],
                                                        controlnet=cn)
    torch.cuda.empty_cache()
    return dict(name=request.param, pipe=pipe, sd=sd, icfg=icfg, inp=inp, rep=rep, rep_o=rep_o, probs_o=probs_o, cn=cn,
                dev=dev, h=h)


def test_extraction_index_sets_vs_device_oracle(bench_case):
    c = bench_case
    rows = bad_rows = 0
    worst = worst_abs = 0.0
    for n, (val_o, idx_o) in c["rep_o"].items():
        val, idx = c["rep"][n]
        bad = (idx != idx_o).squeeze(-1)
        rows += bad.numel()
        bad_rows += int(bad.sum())
        p = c["probs_o"][n].float()  # the oracle's own probabilities [N, heads, L, L]
        top2 = p.topk(2, dim=-1).values
        gap = top2[..., 0] - top2[..., 1]
        ulp = 2.0 ** (torch.floor(torch.log2(top2[..., 0].clamp_min(2.0 ** -14))) - 10)  # fp16 spacing at the top-1 probability
        gap_ulps = gap / ulp
        if bool(bad.any()):
            worst = max(worst, float(gap_ulps[bad].max()))
            worst_abs = max(worst_abs, float(gap[bad].max()))
        assert (val.float() - val_o.float()).abs().max().item() <= 8e-3
    print(f"{c['name']}: top-1 index mismatches vs same-device oracle {bad_rows}/{rows}; largest oracle top-2 gap on a "
          f"mismatching row: {worst:.1f} fp16 ulps of the probability, {worst_abs:.2e} absolute")
    # the two sides feed the softmax with q, k that differ by the fp16 rounding of ~100 upstream layers computed by
    # different kernels: rows whose two largest probabilities are close can swap. The top-1 VALUES are held to 8e-3 above;
    # an index swap is only legitimate where the oracle's own top-2 gap is well inside that: <= 4e-3 absolute (the bar of
    # the fixture test, test_pipeline_gpu.py, halved) and <= 64 fp16 ulps of the probability (6 % relative). The extreme
    # over ~400 000 rows moves from run to run of the code base (23-34 ulps measured in round 2 as attention tile orders
    # changed); the mismatch RATE (0.3 %) does not.
    assert worst <= 64.0 and worst_abs <= 4e-3 and bad_rows / rows < 0.01


@pytest.mark.parametrize("kind", ["guided", "plain"])
def test_single_step_vs_device_oracle(bench_case, kind):
    c = bench_case
    pipe, icfg, inp, h = c["pipe"], c["icfg"], c["inp"], c["h"]
    L = icfg["video_length"]
    step_index = 0 if kind == "guided" else icfg["guidance_steps"]
    timesteps = O.uneven_timesteps(icfg["inference_steps"], icfg["guidance_steps"], icfg["guidance_scale"])
    acp = O.alphas_cumprod()
    lat = h(inp["noisy_latents"])
    if kind == "plain":  # a latent of the magnitude the loop has at the first plain step (after 30 guided steps)
        lat = (lat * 8.0).half()
    # identical motion representation on both sides (the oracle's)
    rep = {n: [v[0].clone(), v[1].clone()] for n, v in c["rep_o"].items()}
    pipe.motion_representation_dict = rep
    pipe._repr_on_device = None
    pipe.scheduler.customized_set_timesteps(icfg["inference_steps"], icfg["guidance_steps"], icfg["guidance_scale"],
                                            device=c["dev"], timestep_spacing_type="uneven")
    # the per-sample state sample_video sets up before its loop (guidance.py: text embeddings, loss weight, SparseCtrl)
    pipe.text_embeddings = h(inp["text_embeddings"])
    pipe.motion_scale = icfg["motion_guidance_weight"]
    pipe.add_controlnet = c["cn"] is not None
    if c["cn"] is not None:  # what sample_video does at guidance.py:224-230
        pipe.controlnet_images = c["cn"]["images"]
    ours = pipe.single_step_video(lat, step_index, pipe.scheduler.timesteps[step_index], {})
    stats = {}
    want = O.single_step(c["sd"], UNET_SD15_CONFIG, icfg, lat, step_index, timesteps, acp, h(inp["text_embeddings"]), rep,
                         stats=stats, controlnet=c["cn"])
    mag = want.float().abs().max().item()
    ulp = _ulp(mag)
    diff = (ours.float() - want.float()).abs()
    print(f"{c['name']} {kind} step: max|x|={mag:.2f} (fp16 ulp {ulp:.4f}); max abs diff {diff.max().item():.4f} = "
          f"{diff.max().item() / ulp:.2f} ulp; mean abs diff {diff.mean().item():.5f} = {diff.mean().item() / ulp:.3f} ulp; "
          f"frac within 1 ulp {(diff <= ulp).float().mean().item():.5f}")
    assert torch.isfinite(ours).all()
    assert diff.max().item() <= 4 * ulp and diff.mean().item() <= 0.5 * ulp
    if kind == "guided":
        g_o = stats["grad"][step_index].to(c["dev"])
        g = pipe.last_gradient.float()
        cos = torch.nn.functional.cosine_similarity(g.flatten(), g_o.flatten(), dim=0).item()
        rel = (g - g_o).abs().max().item() / g_o.abs().max().item()
        print(f"{c['name']} guidance gradient: cosine {cos:.6f}, max-abs rel err {rel:.4f}, max|g|={g_o.abs().max().item():.4f}")
        assert cos >= 0.995 and rel <= 8e-2
    torch.cuda.empty_cache()
