"""Runs the UNMODIFIED reference (/root/reference/motionclone/**) on synthetic inputs. TEST INFRASTRUCTURE ONLY.

Works only in the build container (where /root/reference is mounted); nothing under tests/ -m gpu, smoke() or
bench.py imports this. It exists to (1) generate the golden fixtures under tests/golden/ (oracle/gen_golden.py) and
(2) validate oracle/mc_oracle.py, the CPU restatement that travels to the GPU box.

The reference's own wiring is followed step by step (t2v_video_sample.py:42-73): the nine functions of
motionclone/utils/motionclone_functions.py are bound with __get__ onto the pipeline / scheduler / unet instances,
prep_unet_attention + prep_unet_conv are applied, customized_set_timesteps is called. Only the off-path pieces are
stubbed: VAE, tokenizer/CLIP (synthetic latents / embeddings, BASELINE.json configs) and the video decoder.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MOTIONCLONE_REFERENCE", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "motionclone"))


def _import_reference():
    if not reference_available():
        raise RuntimeError("reference tree not mounted (expected in the build container only)")
    for p in (REFERENCE_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import motionclone.utils.motionclone_functions as mf  # noqa
    import motionclone.models.unet as unet_mod  # noqa
    import motionclone.pipelines.pipeline_animation as pipe_mod  # noqa
    import motionclone.utils.xformer_attention as xa  # noqa
    import motionclone.utils.conv_layer as cl  # noqa
    from diffusers import DDIMScheduler
    from omegaconf import OmegaConf
    return mf, unet_mod, pipe_mod, xa, cl, DDIMScheduler, OmegaConf


class _LatentDist:
    def __init__(self, z):
        self.z = z

    def sample(self, generator=None):
        return self.z


class _StubVAE:
    """vae.encode(video).latent_dist.sample() -> preset clip latents [(f), 4, h, w]; scaling_factor 1 (off path).
    A call whose batch is not the clip length is the SparseCtrl condition-image encode
    (motionclone_functions.py:125): it returns the preset condition latents [(n_img), 4, h, w]."""

    def __init__(self, clip_latents_fchw, dtype, device, cond_latents=None):
        self._z = clip_latents_fchw
        self._cond = cond_latents
        self.dtype, self.device = dtype, device
        self.config = types.SimpleNamespace(scaling_factor=1.0)

    def encode(self, x):
        if self._cond is not None and x.shape[0] != self._z.shape[0]:
            return types.SimpleNamespace(latent_dist=_LatentDist(self._cond))
        return types.SimpleNamespace(latent_dist=_LatentDist(self._z))

    def decode(self, z):
        return types.SimpleNamespace(sample=torch.zeros(z.shape[0], 3, 8, 8))


def build_reference_pipeline(unet_config: dict, infer_cfg: dict, inputs: dict, weight_seed: int = 42,
                             dtype=torch.float32, device="cpu", controlnet_kwargs: dict | None = None):
    """Returns (pipeline, mf). `infer_cfg` carries the YAML keys of configs/t2v_*.yaml plus video_length/height/width."""
    from motionclone_b200.synthetic import NOISE_SCHEDULER_KWARGS, load_synthetic_weights

    mf, unet_mod, pipe_mod, xa, cl, DDIMScheduler, OmegaConf = _import_reference()
    torch.manual_seed(42)  # t2v_video_sample.py:21 (set_all_seed)
    unet = unet_mod.UNet3DConditionModel(**unet_config)
    load_synthetic_weights(unet, weight_seed)
    unet = unet.to(device=device, dtype=dtype).eval()

    controlnet = None
    if controlnet_kwargs is not None:  # i2v_video_sample.py:41-59 (random-init SparseCtrl; zero-convs drawn non-zero)
        import motionclone.models.sparse_controlnet as scn
        unet.config.num_attention_heads = 8
        unet.config.projection_class_embeddings_input_dim = None
        controlnet = scn.SparseControlNetModel.from_unet(unet, controlnet_additional_kwargs=dict(controlnet_kwargs))
        load_synthetic_weights(controlnet, weight_seed + 1)
        controlnet = controlnet.to(device=device).eval()  # cond embedding stays fp16 as the reference builds it
        for p in controlnet.parameters():
            p.requires_grad = False
    pipeline = object.__new__(pipe_mod.AnimationPipeline)  # skip DiffusionPipeline.register_modules plumbing
    pipeline.unet = unet
    pipeline.controlnet = controlnet
    pipeline.scheduler = DDIMScheduler(**NOISE_SCHEDULER_KWARGS)
    pipeline.vae_scale_factor = 8
    clip = inputs["clip_latents"].to(device=device, dtype=dtype)  # [1,4,f,h,w]
    cond_lat = inputs.get("cond_latents")  # [n_img, 4, h, w] for the simplified (latent) condition embedding
    pipeline.vae = _StubVAE(clip[0].permute(1, 0, 2, 3).contiguous(), dtype, torch.device(device),
                            None if cond_lat is None else cond_lat.to(device=device, dtype=dtype))
    text = inputs["text_embeddings"].to(device=device, dtype=dtype)
    pipeline.tokenizer = lambda *a, **k: types.SimpleNamespace(input_ids=torch.zeros(1, 77, dtype=torch.long))
    pipeline.tokenizer.model_max_length = 77
    pipeline.text_encoder = lambda ids: (text[[0]],)
    pipeline._encode_prompt = lambda *a, **k: text
    type(pipeline).device = property(lambda self: torch.device(device))
    type(pipeline).progress_bar = lambda self, total=None: _NullBar()

    # t2v_video_sample.py:57-65
    pipeline.scheduler.customized_step = mf.schedule_customized_step.__get__(pipeline.scheduler)
    pipeline.scheduler.customized_set_timesteps = mf.schedule_set_timesteps.__get__(pipeline.scheduler)
    pipeline.unet.forward = mf.unet_customized_forward.__get__(pipeline.unet)
    pipeline.sample_video = mf.sample_video.__get__(pipeline)
    pipeline.single_step_video = mf.single_step_video.__get__(pipeline)
    pipeline.get_temp_attn_prob = mf.get_temp_attn_prob.__get__(pipeline)
    pipeline.add_noise = mf.add_noise.__get__(pipeline)
    pipeline.compute_temp_loss = mf.compute_temp_loss.__get__(pipeline)
    pipeline.obtain_motion_representation = mf.obtain_motion_representation.__get__(pipeline)
    for p in pipeline.unet.parameters():  # t2v_video_sample.py:67-68
        p.requires_grad = False
    config = OmegaConf.create(dict(infer_cfg))
    config.video_path = "synthetic.mp4"
    config.new_prompt = "synthetic"
    pipeline.input_config, pipeline.unet.input_config = config, config
    pipeline.unet = xa.prep_unet_attention(pipeline.unet, pipeline.input_config.motion_guidance_blocks)
    pipeline.unet = cl.prep_unet_conv(pipeline.unet)
    pipeline.scheduler.customized_set_timesteps(config.inference_steps, config.guidance_steps, config.guidance_scale,
                                                device=device, timestep_spacing_type="uneven")
    # video_preprocess (decord) is off the path: the stub VAE ignores its output; SparseCtrl's image condition reads the
    # clip's pixels ([f, 3, H, W] in [-1, 1]) at motionclone_functions.py:51
    pix = inputs.get("clip_pixels")
    n_frames = int(infer_cfg["video_length"])
    mf.video_preprocess = (lambda *a, **k: pix.to(device=device, dtype=dtype)) if pix is not None \
        else (lambda *a, **k: torch.zeros(n_frames, 3, 8, 8))  # batch == clip length: the stub VAE's "clip" branch
    return pipeline, mf


class _NullBar:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, *a):
        pass


def run_reference(unet_config: dict, infer_cfg: dict, inputs: dict, repr_path: str, weight_seed: int = 42,
                  dtype=torch.float32, device="cpu", record_grad_steps=(0,), controlnet_kwargs: dict | None = None):
    """obtain_motion_representation (motionclone_functions.py:25-82) then the sample_video loop (:102-171).

    Returns a dict of CPU fp32 tensors: motion representation, per-step latents, per-step losses and the
    guidance gradient at `record_grad_steps`.
    """
    pipeline, mf = build_reference_pipeline(unet_config, infer_cfg, inputs, weight_seed, dtype, device, controlnet_kwargs)
    use_cn = controlnet_kwargs is not None
    out = {}

    # --- extraction: randn_tensor(generator) is replaced by the preset clip noise (CUDA/CPU streams differ) ---
    noise = inputs["clip_noise"].to(device=device, dtype=dtype)
    mf.randn_tensor = lambda shape, generator=None, device=None, dtype=None: noise
    pipeline.obtain_motion_representation(generator=None, motion_representation_path=repr_path, use_controlnet=use_cn)
    rep = torch.load(repr_path)
    probs = pipeline.get_temp_attn_prob()  # processors still hold the extraction pass's q,k
    out["extract_probs_0"] = next(iter(probs.values())).float().cpu()
    for i, pr in enumerate(probs.values()):  # top-1 minus top-2 probability of every guided module's rows: an index
        top2 = pr.float().topk(2, dim=-1).values   # mismatch is only acceptable where the REFERENCE's own row is a near-tie
        out[f"extract_top2gap_{i}"] = (top2[..., 0] - top2[..., 1]).cpu()
    out["repr_names"] = list(rep.keys())
    for i, (k, (val, idx)) in enumerate(rep.items()):
        out[f"repr_val_{i}"] = val.float().cpu()
        out[f"repr_idx_{i}"] = idx.cpu()

    # --- one plain UNet forward (pins the oracle's UNet restatement in isolation) ---
    with torch.no_grad():
        x0 = inputs["noisy_latents"].to(device=device, dtype=dtype)
        t500 = torch.tensor(500, device=device)
        out["unet_fwd_t500_cond"] = pipeline.unet(x0, t500, encoder_hidden_states=pipeline.text_encoder(None)[0] * 0
                                                  + inputs["text_embeddings"].to(device=device, dtype=dtype)[[1]]
                                                  ).sample.float().cpu()

    # --- sampling: record every step's latents; losses and gradients through thin recording wrappers ---
    latents_per_step, losses, grads = [], [], {}
    ref_step = pipeline.single_step_video
    ref_loss = pipeline.compute_temp_loss
    ref_autograd_grad = torch.autograd.grad
    state = {"step": -1}

    def rec_loss(d):
        v = ref_loss(d)
        losses.append(v.detach().float().cpu())
        return v

    def rec_grad(*a, **k):
        g = ref_autograd_grad(*a, **k)
        if state["step"] in record_grad_steps:
            grads[state["step"]] = g[0].detach().float().cpu()
        return g

    def rec_step(noisy_latents, step_index, step_t, extra):
        state["step"] = step_index
        r = ref_step(noisy_latents, step_index, step_t, extra)
        latents_per_step.append(r.detach().float().cpu())
        return r

    pipeline.compute_temp_loss = rec_loss
    pipeline.single_step_video = rec_step
    torch.autograd.grad = rec_grad
    try:
        pipeline.sample_video(generator=None, noisy_latents=inputs["noisy_latents"].to(device=device, dtype=dtype),
                              add_controlnet=use_cn)
    finally:
        torch.autograd.grad = ref_autograd_grad
    out["timesteps"] = pipeline.scheduler.timesteps.cpu()
    out["latents_per_step"] = torch.stack(latents_per_step)
    out["losses"] = torch.stack(losses) if losses else torch.zeros(0)
    for s, g in grads.items():
        out[f"grad_step_{s}"] = g
    return out, pipeline
