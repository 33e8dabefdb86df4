"""CPU restatement of MotionClone's guided denoising path. TEST INFRASTRUCTURE — never imported by motionclone_b200/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file, and
only as the checker / CPU baseline. It is plain functional PyTorch over a state dict (fp32 on CPU is the truth mode;
the same code runs in fp16 on a GPU as the same-device comparator), written from the reference's algorithm, each
function citing the reference lines it follows (paths relative to /root/reference/motionclone/).

PARITY PINNING: the reference ships no tests or golden vectors. This restatement is pinned instead
against outputs of the UNMODIFIED reference run in the build container over a shim of its missing third-party
imports (oracle/ref_runner.py, oracle/gen_golden.py -> tests/golden/ref_*.npz), see tests/test_oracle_golden.py.
Third-party arithmetic restated from the diffusers==0.16.0 release (environment.yaml:13; not vendored): GEGLU
feed-forward, Timesteps/TimestepEmbedding, DDIMScheduler betas -- each has a known-answer test.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------------------------
# scheduler state (diffusers 0.16 DDIMScheduler.__init__, used at utils/motionclone_functions.py:332-333)
# ----------------------------------------------------------------------------------------------------------------
def alphas_cumprod(beta_start=0.00085, beta_end=0.012, num_train_timesteps=1000) -> Tensor:
    betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    return torch.cumprod(1.0 - betas, dim=0)


def uneven_timesteps(num_inference_steps: int, guidance_steps: int, guidance_scale: float,
                     num_train_timesteps: int = 1000) -> np.ndarray:
    """utils/motionclone_functions.py:432-445 ("uneven" spacing)."""
    split = int((1 - guidance_scale) * num_train_timesteps)
    tg = np.linspace(split, num_train_timesteps - 1, guidance_steps).round()[::-1].copy().astype(np.int64)
    tv = np.linspace(0, split - 1, num_inference_steps - guidance_steps).round()[::-1].copy().astype(np.int64)
    return np.concatenate((tg, tv))


def loss_scale(step_index: int, guidance_steps: int, warm_up_steps: int, cool_up_steps: int) -> float:
    """utils/motionclone_functions.py:228-234 — both multipliers can apply; the cool-down test is a strict '>'."""
    s = 1.0
    if step_index < warm_up_steps:
        s *= (step_index + 1) / warm_up_steps
    if step_index > guidance_steps - cool_up_steps:
        s *= (guidance_steps - step_index) / cool_up_steps
    return s


# ----------------------------------------------------------------------------------------------------------------
# elementwise pieces: add_noise, CFG combine, guided DDIM update
# ----------------------------------------------------------------------------------------------------------------
def add_noise(acp: Tensor, timestep: int, x0: Tensor, noise: Tensor) -> Tensor:
    """utils/motionclone_functions.py:19-23. `acp[timestep]` is a 0-dim fp32 CPU tensor; results take x0's dtype."""
    a = acp[timestep]
    return a ** 0.5 * x0 + (1 - a) ** 0.5 * noise


def cfg_combine(eps_cond: Tensor, eps_uncond: Tensor, cfg_scale: float) -> Tensor:
    """utils/motionclone_functions.py:239 / :255 — cond + s*(cond - uncond)."""
    return eps_cond + cfg_scale * (eps_cond - eps_uncond)


def ddim_scalars(acp: Tensor, timesteps: Sequence[int], step_index: int):
    """utils/motionclone_functions.py:326-335: (alpha_t, alpha_prev) as 0-dim fp32 tensors."""
    t = int(timesteps[step_index])
    prev_t = int(timesteps[step_index + 1]) if step_index + 1 < len(timesteps) else -1
    a_t = acp[t]
    a_prev = acp[prev_t] if prev_t >= 0 else torch.tensor(1.0)
    return a_t, a_prev


def ddim_guided_step(eps: Tensor, x: Tensor, score: Optional[Tensor], a_t: Tensor, a_prev: Tensor,
                     guidance_scale: float = 1.0, reciprocal_div: bool = False) -> Tensor:
    """utils/motionclone_functions.py:339-389 with eta=0, prediction_type=epsilon, no clip/threshold.

    x0 uses the un-guided eps (:340); only the direction term uses eps' = eps - gs*sqrt(1-a_t)*score (:382, :386).
    Every op rounds to the tensor dtype, exactly as the eager op sequence does. `reciprocal_div=True` reproduces the
    CUDA TensorIterator behaviour for `tensor / cpu_scalar` (multiply by the fp32 reciprocal); CPU divides.
    """
    a_t = a_t.to(torch.float32).cpu()
    a_prev = a_prev.to(torch.float32).cpu()
    beta_t = 1 - a_t
    if reciprocal_div:
        x0 = (x - beta_t ** 0.5 * eps) * (1.0 / (a_t ** 0.5))
    else:
        x0 = (x - beta_t ** 0.5 * eps) / a_t ** 0.5
    pred_eps = eps
    if score is not None and guidance_scale > 0.0:
        pred_eps = pred_eps - guidance_scale * (1 - a_t) ** 0.5 * score
    direction = (1 - a_prev - 0.0) ** 0.5 * pred_eps
    return a_prev ** 0.5 * x0 + direction


def cfg_ddim_step_fp16_sequence(eps_cond: Tensor, eps_uncond: Tensor, x: Tensor, score: Optional[Tensor],
                                cfg_scale: float, a_t: Tensor, a_prev: Tensor, guidance_scale: float = 1.0) -> Tensor:
    """What the reference's eager CUDA ops compute for fp16 tensors at utils/motionclone_functions.py:239 + :339-389,
    restated on fp32 values with an explicit fp16 rounding after every op. On CUDA a 0-dim
    fp32 CPU operand stays fp32 (opmath) and `tensor / cpu_scalar` multiplies by the fp32 reciprocal; CPU eager half
    ops instead cast the 0-dim operand to fp16 first, so `ddim_guided_step` on CPU half tensors is NOT this sequence.
    """
    h = lambda t: t.to(torch.float16).to(torch.float32)  # noqa: E731
    a_t = a_t.to(torch.float32).cpu()
    a_prev = a_prev.to(torch.float32).cpu()
    sb, inv_sa = (1 - a_t) ** 0.5, 1.0 / (a_t ** 0.5)
    sap, c = a_prev ** 0.5, (1 - a_prev - 0.0) ** 0.5
    ec, eu, xf = eps_cond.float(), eps_uncond.float(), x.float()
    e = h(ec + h(cfg_scale * h(ec - eu)))
    x0 = h(h(xf - h(sb * e)) * inv_sa)
    e2 = e
    if score is not None and guidance_scale > 0.0:
        e2 = h(e - h((guidance_scale * (1 - a_t) ** 0.5) * score.float()))
    return h(h(sap * x0) + h(c * e2)).to(torch.float16)


# ----------------------------------------------------------------------------------------------------------------
# temporal attention, top-1 extraction, motion loss
# ----------------------------------------------------------------------------------------------------------------
def heads_to_batch(t: Tensor, heads: int) -> Tensor:
    """models/attention.py:367-372."""
    b, s, d = t.shape
    return t.reshape(b, s, heads, d // heads).permute(0, 2, 1, 3).reshape(b * heads, s, d // heads)


def batch_to_heads(t: Tensor, heads: int) -> Tensor:
    """models/attention.py:374-379."""
    b, s, d = t.shape
    return t.reshape(b // heads, heads, s, d).permute(0, 2, 1, 3).reshape(b // heads, s, d * heads)


def attention_probs(q: Tensor, k: Tensor, scale: float) -> Tensor:
    """models/attention.py:466-483 and :594-609: baddbmm(alpha=scale) -> softmax(-1) -> input dtype.

    q,k are [B*heads, S, dh]; scores are materialised in q.dtype (fp16 rounding on GPU) before the softmax.
    """
    scores = torch.baddbmm(torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device),
                           q, k.transpose(-1, -2), beta=0, alpha=scale)
    return scores.softmax(dim=-1).to(q.dtype)


def attention_math(q: Tensor, k: Tensor, v: Tensor, heads: int, scale: float) -> Tensor:
    """models/attention.py:461-490 (`_attention`): q,k,v are [B, S, C] pre head split; returns [B, S, C]."""
    qh, kh, vh = (heads_to_batch(t, heads) for t in (q, k, v))
    out = torch.bmm(attention_probs(qh, kh, scale).to(vh.dtype), vh)
    return batch_to_heads(out, heads)


def temporal_probs(q: Tensor, k: Tensor, heads: int, scale: float) -> Tensor:
    """utils/motionclone_functions.py:267-280: recorded q,k [b*d, f, C] -> probs [b*d, heads, f, f]."""
    p = attention_probs(heads_to_batch(q, heads).contiguous(), heads_to_batch(k, heads).contiguous(), scale)
    return p.reshape(-1, heads, p.shape[1], p.shape[2])


def top1(probs: Tensor) -> Tuple[Tensor, Tensor]:
    """utils/motionclone_functions.py:79: topk(k=1, dim=-1) -> (values, indices.uint8). Ties: lowest index."""
    val, idx = torch.topk(probs, k=1, dim=-1)
    return val, idx.to(torch.uint8)


def top1_lowest_index(probs: Tensor) -> Tuple[Tensor, Tensor]:
    """Deterministic statement of the tie rule (lowest index among maxima); what the CUDA kernel implements."""
    val = probs.max(dim=-1, keepdim=True).values
    L = probs.shape[-1]
    ar = torch.arange(L, device=probs.device).expand_as(probs)
    idx = torch.where(probs == val, ar, torch.full_like(ar, L)).min(dim=-1, keepdim=True).values
    return val, idx.to(torch.uint8)


def motion_loss(probs: Dict[str, Tensor], representation: Dict[str, Sequence[Tensor]]) -> Tensor:
    """utils/motionclone_functions.py:85-100: sum over modules of mse(gather(P, idx_ref), val_ref)."""
    losses = []
    for name, p in probs.items():
        val_ref, idx_ref = representation[name]
        cur = torch.gather(p, index=idx_ref.to(torch.int64).to(p.device), dim=-1)
        losses.append(F.mse_loss(cur, val_ref.to(dtype=cur.dtype, device=cur.device).detach()))
    return torch.stack(losses).sum()


def motion_loss_dscores_closed_form(p: Tensor, idx_ref: Tensor, val_ref: Tensor, weight: float) -> Tensor:
    """d(weight * mean((P[idx]-ref)^2)) / d(scores) in closed form:
    dP[r, idx_r] = 2*weight*(P_idx - ref)/N ;  dS_j = P_j * (delta_{j,idx} - P_idx) * dP."""
    idx = idx_ref.to(torch.int64)
    p_idx = torch.gather(p, -1, idx)
    n = p_idx.numel()
    dp = 2.0 * weight * (p_idx - val_ref.to(p.dtype)) / n
    onehot = torch.zeros_like(p).scatter_(-1, idx, 1.0)
    return p * (onehot - p_idx) * dp


# ----------------------------------------------------------------------------------------------------------------
# UNet3D (functional over a state dict).  Layout follows the reference: 5-D [b, c, f, h, w] between blocks.
# ----------------------------------------------------------------------------------------------------------------
def positional_encoding(d_model: int, max_len: int = 32) -> Tensor:
    """models/motion_module.py:237-241 (fp32 build; max_len default 32 at :60)."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def timestep_embedding(timesteps: Tensor, dim: int) -> Tensor:
    """diffusers 0.16 Timesteps(dim, flip_sin_to_cos=True, freq_shift=0) (models/unet.py:101): fp32 [cos | sin]."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class _SD:
    """Prefix view over a flat state dict."""

    def __init__(self, sd: Dict[str, Tensor], prefix: str = ""):
        self.sd, self.prefix = sd, prefix

    def sub(self, name: str) -> "_SD":
        return _SD(self.sd, f"{self.prefix}{name}.")

    def __getitem__(self, name: str) -> Tensor:
        return self.sd[self.prefix + name]

    def has(self, name: str) -> bool:
        return (self.prefix + name) in self.sd

    def linear(self, name: str, x: Tensor) -> Tensor:
        return F.linear(x, self[name + ".weight"], self.sd.get(self.prefix + name + ".bias"))


def _frames_op(fn, x: Tensor) -> Tensor:
    """models/resnet.py:10-29 (Inflated conv / groupnorm): fold f into the batch, apply per frame, unfold."""
    b, c, f, h, w = x.shape
    y = fn(x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w))
    return y.reshape(b, f, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def _conv(sd: _SD, name: str, x: Tensor, stride=1, padding=1) -> Tensor:
    return _frames_op(lambda t: F.conv2d(t, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding), x)


def _gn(sd: _SD, name: str, x: Tensor, groups: int, eps: float) -> Tensor:
    return _frames_op(lambda t: F.group_norm(t, groups, sd[name + ".weight"], sd[name + ".bias"], eps), x)


def _resnet(sd: _SD, x: Tensor, temb: Tensor, groups: int, eps: float) -> Tensor:
    """models/resnet.py:183-213 == utils/conv_layer.py:4-48 numerically (time_embedding_norm='default', scale 1)."""
    h = F.silu(_gn(sd, "norm1", x, groups, eps))
    h = _conv(sd, "conv1", h)
    h = h + sd.linear("time_emb_proj", F.silu(temb))[:, :, None, None, None]
    h = F.silu(_gn(sd, "norm2", h, groups, eps))
    h = _conv(sd, "conv2", h)
    if sd.has("conv_shortcut.weight"):
        x = _conv(sd, "conv_shortcut", x, padding=0)
    return x + h


def _feed_forward(sd: _SD, x: Tensor) -> Tensor:
    """diffusers 0.16 FeedForward(activation_fn='geglu'): Linear(d, 8d) -> h*gelu_erf(gate) -> Linear(4d, d)."""
    h, gate = sd.linear("net.0.proj", x).chunk(2, dim=-1)
    return sd.linear("net.2", h * F.gelu(gate))


# models/attention.py:449-459: with xformers enabled (the reference's GPU configuration, pipelines' enable_xformers...)
# spatial attention goes through xformers.ops.memory_efficient_attention (:535-542); without it through the math path
# (:461-490). "math" is the parity truth (CPU, and fp16 on a GPU); "sdpa" = the library flash kernel standing in for the
# absent xformers wheel, used ONLY by bench.py's same-GPU comparator leg (`gpu_reference`), never by a parity test.
SPATIAL_ATTENTION = "math"


def attention_xformers_seam(q: Tensor, k: Tensor, v: Tensor, heads: int, scale: float) -> Tensor:
    """models/attention.py:535-542 with torch's fused SDPA in the role of xformers.ops.memory_efficient_attention."""
    b, s, c = q.shape
    q4, k4, v4 = (t.reshape(b, t.shape[1], heads, c // heads).transpose(1, 2) for t in (q, k, v))
    return F.scaled_dot_product_attention(q4, k4, v4, scale=scale).transpose(1, 2).reshape(b, s, c)


def _cross_attention(sd: _SD, x: Tensor, ctx: Optional[Tensor], heads: int) -> Tensor:
    """models/attention.py:387-459 via the math path (:461-490) [or the xformers seam, see SPATIAL_ATTENTION]."""
    ctx = x if ctx is None else ctx
    q, k, v = sd.linear("to_q", x), sd.linear("to_k", ctx), sd.linear("to_v", ctx)
    scale = (q.shape[-1] // heads) ** -0.5
    core = attention_xformers_seam if SPATIAL_ATTENTION == "sdpa" else attention_math
    return sd.linear("to_out.0", core(q, k, v, heads, scale))


def _spatial_transformer(sd: _SD, x: Tensor, text: Tensor, heads: int, groups: int) -> Tensor:
    """models/attention.py:95-142 (Transformer3DModel.forward) + :256-300 (BasicTransformerBlock.forward)."""
    b, c, f, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    ctx = text.repeat_interleave(f, dim=0)  # 'b n c -> (b f) n c'  (:100)
    res = xf
    t = F.group_norm(xf, groups, sd["norm.weight"], sd["norm.bias"], 1e-6)
    t = F.conv2d(t, sd["proj_in.weight"], sd["proj_in.bias"])
    t = t.permute(0, 2, 3, 1).reshape(b * f, h * w, c)
    blk = sd.sub("transformer_blocks.0")
    ln = lambda n, z: F.layer_norm(z, (c,), blk[n + ".weight"], blk[n + ".bias"], 1e-5)  # noqa: E731
    t = _cross_attention(blk.sub("attn1"), ln("norm1", t), None, heads) + t
    t = _cross_attention(blk.sub("attn2"), ln("norm2", t), ctx, heads) + t
    t = _feed_forward(blk.sub("ff"), ln("norm3", t)) + t
    t = t.reshape(b * f, h, w, c).permute(0, 3, 1, 2).contiguous()
    t = F.conv2d(t, sd["proj_out.weight"], sd["proj_out.bias"])
    out = t + res
    return out.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def _versatile_attention(sd: _SD, x: Tensor, video_length: int, heads: int, pe: Tensor, record: Optional[dict],
                         name: str) -> Tensor:
    """models/motion_module.py:274-345 (VersatileAttention.forward, Temporal mode, self-attention)."""
    d = x.shape[1]
    bf, _, c = x.shape
    b = bf // video_length
    t = x.reshape(b, video_length, d, c).permute(0, 2, 1, 3).reshape(b * d, video_length, c)  # '(b f) d c -> (b d) f c'
    t = t + pe[:, :video_length].to(t.dtype)  # :281-282
    q, k, v = sd.linear("to_q", t), sd.linear("to_k", t), sd.linear("to_v", t)
    if record is not None:
        record[name] = (q, k)  # processor.record_qkv (utils/xformer_attention.py:31-34)
    scale = (c // heads) ** -0.5
    o = sd.linear("to_out.0", attention_math(q, k, v, heads, scale))
    return o.reshape(b, d, video_length, c).permute(0, 2, 1, 3).reshape(bf, d, c)  # '(b d) f c -> (b f) d c'


def _motion_module(sd: _SD, x: Tensor, heads: int, groups: int, pes: Dict[int, Tensor], record: Optional[dict],
                   name: str, guided: bool) -> Tensor:
    """models/motion_module.py:137-161 (TemporalTransformer3DModel.forward) + :213-225 (TemporalTransformerBlock)."""
    tt = sd.sub("temporal_transformer")
    b, c, f, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    res = xf
    t = F.group_norm(xf, groups, tt["norm.weight"], tt["norm.bias"], 1e-6)
    t = t.permute(0, 2, 3, 1).reshape(b * f, h * w, c)
    t = tt.linear("proj_in", t)
    blk = tt.sub("transformer_blocks.0")
    if c not in pes:
        pes[c] = positional_encoding(c).to(device=x.device)
    n_attn = 0
    while blk.has(f"attention_blocks.{n_attn}.to_q.weight"):
        n_attn += 1
    for i in range(n_attn):  # UNet: (Temporal_Self, Temporal_Self); SparseCtrl: (Temporal_Self,) — configs/sparsectrl/*.yaml:14
        n = F.layer_norm(t, (c,), blk[f"norms.{i}.weight"], blk[f"norms.{i}.bias"], 1e-5)
        aname = f"{name}.temporal_transformer.transformer_blocks.0.attention_blocks.{i}"
        t = _versatile_attention(blk.sub(f"attention_blocks.{i}"), n, f, heads, pes[c],
                                 record if guided else None, aname) + t
    t = _feed_forward(blk.sub("ff"), F.layer_norm(t, (c,), blk["ff_norm.weight"], blk["ff_norm.bias"], 1e-5)) + t
    t = tt.linear("proj_out", t)
    t = t.reshape(b * f, h, w, c).permute(0, 3, 1, 2).contiguous()
    out = t + res
    return out.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def _down_and_mid(sd: _SD, cfg: dict, x: Tensor, emb: Tensor, text: Tensor, mm_heads: int, pes: dict,
                  record: Optional[dict], guided, prefix: str = ""):
    """conv_in output -> (mid-block output, skip list). Shared by the UNet (models/unet_blocks.py:382-421, :493-521,
    :271-278) and SparseCtrl (models/sparse_controlnet.py:529-552), which reuse the same block classes."""
    groups, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    heads = cfg["attention_head_dim"]  # used as the head COUNT (models/unet_blocks.py:343-344)
    chans = cfg["block_out_channels"]
    skips = [x]
    for i, btype in enumerate(cfg["down_block_types"]):
        blk = sd.sub(f"down_blocks.{i}")
        for j in range(cfg["layers_per_block"]):
            x = _resnet(blk.sub(f"resnets.{j}"), x, emb, groups, eps)
            if btype.startswith("CrossAttn"):
                x = _spatial_transformer(blk.sub(f"attentions.{j}"), x, text, heads, groups)
            mname = f"{prefix}down_blocks.{i}.motion_modules.{j}"
            x = _motion_module(blk.sub(f"motion_modules.{j}"), x, mm_heads, groups, pes, record, mname, guided(mname))
            skips.append(x)
        if i < len(chans) - 1:
            x = _conv(blk, "downsamplers.0.conv", x, stride=2, padding=1)
            skips.append(x)
    mid = sd.sub("mid_block")  # no motion module: motion_module_mid_block=false
    x = _resnet(mid.sub("resnets.0"), x, emb, groups, eps)
    x = _spatial_transformer(mid.sub("attentions.0"), x, text, heads, groups)
    x = _resnet(mid.sub("resnets.1"), x, emb, groups, eps)
    return x, skips


def _time_embedding(sd: _SD, timestep, sample: Tensor, dim: int) -> Tensor:
    if not torch.is_tensor(timestep):
        timestep = torch.tensor([timestep], dtype=torch.int64, device=sample.device)
    elif timestep.dim() == 0:
        timestep = timestep[None].to(sample.device)
    timestep = timestep.expand(sample.shape[0])
    t_emb = timestep_embedding(timestep, dim).to(sample.dtype)
    te = sd.sub("time_embedding")
    return te.linear("linear_2", F.silu(te.linear("linear_1", t_emb)))


def controlnet_forward(sd_flat: Dict[str, Tensor], cfg: dict, cn_kwargs: dict, sample: Tensor, timestep, text: Tensor,
                       controlnet_cond: Tensor, conditioning_mask: Tensor, conditioning_scale: float = 1.0):
    """models/sparse_controlnet.py:450-587 (SparseControlNetModel.forward; guess_mode False, no global pooling).

    `cfg` is the UNet config it was built from (from_unet, :317-370); `cn_kwargs` the controlnet_additional_kwargs of
    configs/sparsectrl/*.yaml. Returns (12 down residuals, mid residual), each `[b, c, f, h, w]`.
    """
    sd = _SD(sd_flat)
    chans = cfg["block_out_channels"]
    mm_heads = cn_kwargs["motion_module_kwargs"]["num_attention_heads"]
    text = text.repeat(sample.shape[0] // text.shape[0], 1, 1)  # :488
    emb = _time_embedding(sd, timestep, sample, chans[0]).to(sample.dtype)
    if cn_kwargs.get("set_noisy_sample_input_to_zero", False):  # :516-518: conv_in(0) == bias
        b, _, f, h, w = sample.shape
        x = sd["conv_in.bias"].reshape(1, -1, 1, 1, 1).expand(b, -1, f, h, w)
    else:
        x = _conv(sd, "conv_in", sample)
    cond = torch.cat([controlnet_cond, conditioning_mask], dim=1).to(torch.float16)  # :522-523 (fp16 hard-coded)
    ce = sd.sub("controlnet_cond_embedding")
    if cn_kwargs.get("use_simplified_condition_embedding", False):  # :181-184: one zero-initialised 3x3 conv
        e = _frames_op(lambda t: F.conv2d(t, sd["controlnet_cond_embedding.weight"], sd["controlnet_cond_embedding.bias"],
                                          padding=1), cond)
    else:  # SparseControlNetConditioningEmbedding, :49-82
        e = F.silu(_conv(ce, "conv_in", cond))
        n_blocks = 0
        while ce.has(f"blocks.{n_blocks}.weight"):
            n_blocks += 1
        for i in range(n_blocks):
            e = F.silu(_conv(ce, f"blocks.{i}", e, stride=2 if i % 2 == 1 else 1))
        e = _conv(ce, "conv_out", e)
    x = x + e  # :527
    x, skips = _down_and_mid(sd, cfg, x, emb, text, mm_heads, {}, None, lambda n: False)
    down = [_conv(sd, f"controlnet_down_blocks.{i}", s_, padding=0) * conditioning_scale for i, s_ in enumerate(skips)]
    mid = _conv(sd, "controlnet_mid_block", x, padding=0) * conditioning_scale
    return down, mid


def controlnet_condition(images: Tensor, image_index: Sequence[int], video_length: int, dtype, device):
    """utils/motionclone_functions.py:54-63 / :178-188: zero-filled condition + 1-channel mask with the given frames set.
    images `[1, c, n_img, h, w]`."""
    shp = list(images.shape)
    shp[2] = video_length
    cond = torch.zeros(shp, dtype=dtype, device=device)
    mask = torch.zeros([shp[0], 1] + shp[2:], dtype=dtype, device=device)
    cond[:, :, list(image_index)] = images.to(device=device, dtype=dtype)
    mask[:, :, list(image_index)] = 1
    return cond, mask


def unet_forward(sd_flat: Dict[str, Tensor], cfg: dict, sample: Tensor, timestep, text: Tensor,
                 record: Optional[dict] = None, guidance_blocks: Sequence[str] = ("up_blocks.1",),
                 only_motion_feature: bool = False, down_residuals: Optional[Sequence[Tensor]] = None,
                 mid_residual: Optional[Tensor] = None):
    """utils/motionclone_functions.py:478-662 (unet_customized_forward) over the topology of models/unet.py:42-249.

    `record` (dict) receives {module_name: (q, k)} for VersatileAttention modules whose name contains one of
    `guidance_blocks` (utils/xformer_attention.py:45-52). up_blocks beyond the last guidance block run under
    no_grad (:602, :629); with `only_motion_feature` the function returns 0 there (:627-628).
    """
    sd = _SD(sd_flat)
    groups, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    heads = cfg["attention_head_dim"]  # used as the head COUNT (models/unet_blocks.py:343-344)
    mm_heads = cfg["motion_module_kwargs"]["num_attention_heads"]
    chans = cfg["block_out_channels"]
    pes: Dict[int, Tensor] = {}
    guided = lambda n: any(g in n for g in guidance_blocks)  # noqa: E731  (utils/util.py:434-440)

    emb = _time_embedding(sd, timestep, sample, chans[0])  # :545-551

    x = _conv(sd, "conv_in", sample)
    x, skips = _down_and_mid(sd, cfg, x, emb, text, mm_heads, pes, record, guided)
    # NOTE: the reference adds the mid residual AFTER the mid block and the down residuals to the skip list (:582-598)
    if down_residuals is not None:
        skips = [s_ + (r.unsqueeze(2) if r.dim() == 4 else r) for s_, r in zip(skips, down_residuals)]
    if mid_residual is not None:
        x = x + (mid_residual.unsqueeze(2) if mid_residual.dim() == 4 else mid_residual)

    cut = int(guidance_blocks[-1].split(".")[-1])
    for i, btype in enumerate(cfg["up_block_types"]):  # models/unet_blocks.py:621-667, :735-760
        if i > cut and only_motion_feature:
            return 0
        ctx = torch.enable_grad() if (i <= cut and torch.is_grad_enabled()) else torch.no_grad()
        with ctx:
            blk = sd.sub(f"up_blocks.{i}")
            for j in range(cfg["layers_per_block"] + 1):
                x = torch.cat([x, skips.pop()], dim=1)
                x = _resnet(blk.sub(f"resnets.{j}"), x, emb, groups, eps)
                if btype.startswith("CrossAttn"):
                    x = _spatial_transformer(blk.sub(f"attentions.{j}"), x, text, heads, groups)
                mname = f"up_blocks.{i}.motion_modules.{j}"
                x = _motion_module(blk.sub(f"motion_modules.{j}"), x, mm_heads, groups, pes, record, mname,
                                   guided(mname))
            if i < len(chans) - 1:  # models/resnet.py:65 nearest 2x then conv
                x = _frames_op(lambda t: F.interpolate(t, scale_factor=2.0, mode="nearest"), x)
                x = _conv(blk, "upsamplers.0.conv", x)
    x = F.silu(_gn(sd, "conv_norm_out", x, groups, eps))
    return _conv(sd, "conv_out", x)


# ----------------------------------------------------------------------------------------------------------------
# the two entry points: motion-representation extraction and the guided sampling loop
# ----------------------------------------------------------------------------------------------------------------
def record_to_probs(record: Dict[str, Tuple[Tensor, Tensor]], heads: int) -> Dict[str, Tensor]:
    """utils/motionclone_functions.py:260-283 (get_temp_attn_prob)."""
    out = {}
    for name, (q, k) in record.items():
        scale = (q.shape[-1] // heads) ** -0.5
        out[name] = temporal_probs(q, k, heads, scale)
    return out


def _controlnet_residuals(cn: dict, cfg: dict, latents: Tensor, t: int, text: Tensor, images: Tensor):
    """utils/motionclone_functions.py:46-72 / :176-197. cn = dict(sd, kwargs, image_index, scale)."""
    cond, mask = controlnet_condition(images, cn["image_index"], latents.shape[2], latents.dtype, latents.device)
    with torch.no_grad():
        return controlnet_forward(cn["sd"], cfg, cn["kwargs"], latents, t, text, cond, mask, cn["scale"])


@torch.no_grad()
def obtain_motion_representation(sd, cfg, clip_latents: Tensor, clip_noise: Tensor, uncond_text: Tensor,
                                 add_noise_step: int = 400, guidance_blocks=("up_blocks.1",),
                                 controlnet: Optional[dict] = None, clip_pixels: Optional[Tensor] = None):
    """utils/motionclone_functions.py:25-82 with the VAE/CLIP outputs given (synthetic). With `controlnet`
    (dict(sd, kwargs, image_index, scale)) the condition comes from the CLIP itself: its latents (simplified
    embedding, :49) or its pixels `[f, 3, H, W]` in [-1, 1] mapped to [0, 1] (:51-52)."""
    acp = alphas_cumprod()
    noisy = add_noise(acp, int(add_noise_step), clip_latents, clip_noise)
    record: Dict[str, Tuple[Tensor, Tensor]] = {}
    down = mid = None
    if controlnet is not None:
        if controlnet["kwargs"].get("use_simplified_condition_embedding", False):
            images = clip_latents[:, :, list(controlnet["image_index"])]
        else:
            pix = (clip_pixels.unsqueeze(0).permute(0, 2, 1, 3, 4).to(clip_latents) + 1) / 2
            images = pix[:, :, list(controlnet["image_index"])]
        down, mid = _controlnet_residuals(controlnet, cfg, noisy, int(add_noise_step), uncond_text, images)
    unet_forward(sd, cfg, noisy, int(add_noise_step), uncond_text, record=record, guidance_blocks=guidance_blocks,
                 only_motion_feature=True, down_residuals=down, mid_residual=mid)
    probs = record_to_probs(record, cfg["motion_module_kwargs"]["num_attention_heads"])
    return {k: list(top1(p)) for k, p in probs.items()}, probs


def single_step(sd, cfg, icfg: dict, latents: Tensor, step_index: int, timesteps, acp: Tensor, text: Tensor,
                representation, reciprocal_div: bool = False, stats: Optional[dict] = None,
                controlnet: Optional[dict] = None) -> Tensor:
    """utils/motionclone_functions.py:173-257 (single_step_video). `controlnet` = dict(sd, kwargs, image_index, scale,
    images `[1, c, n_img, h, w]`) runs SparseCtrl at b=2 under no_grad (:176-197) and splits its residuals per pass."""
    t = int(timesteps[step_index])
    du = dc = dpair = mu = mc_ = mpair = None
    if controlnet is not None:
        dpair, mpair = _controlnet_residuals(controlnet, cfg, latents.expand(2, -1, -1, -1, -1), t, text,
                                             controlnet["images"])
        du, dc = [r[[0]].detach() for r in dpair], [r[[1]].detach() for r in dpair]  # :205-208
        mu, mc_ = mpair[[0]].detach(), mpair[[1]].detach()
    a_t, a_prev = ddim_scalars(acp, timesteps, step_index)
    mm_heads = cfg["motion_module_kwargs"]["num_attention_heads"]
    gb = tuple(icfg["motion_guidance_blocks"])
    if step_index < icfg["guidance_steps"]:
        control = latents.clone().detach().requires_grad_(True)
        with torch.no_grad():
            eps_u = unet_forward(sd, cfg, latents, t, text[[0]], guidance_blocks=gb, down_residuals=du, mid_residual=mu)
        record: Dict[str, Tuple[Tensor, Tensor]] = {}
        with torch.enable_grad():
            eps_c = unet_forward(sd, cfg, control, t, text[[1]], record=record, guidance_blocks=gb, down_residuals=dc,
                                 mid_residual=mc_)
            probs = record_to_probs(record, mm_heads)
            raw = motion_loss(probs, representation)
            loss = icfg["motion_guidance_weight"] * raw
            if step_index < icfg["warm_up_steps"]:  # :228-230 (two separate multiplies, as the reference does)
                loss = ((step_index + 1) / icfg["warm_up_steps"]) * loss
            if step_index > icfg["guidance_steps"] - icfg["cool_up_steps"]:  # :232-234
                loss = ((icfg["guidance_steps"] - step_index) / icfg["cool_up_steps"]) * loss
            grad = torch.autograd.grad(loss, control, allow_unused=True)[0]
        if stats is not None:
            stats.setdefault("loss_unscaled", []).append(raw.detach().float().cpu())
            stats.setdefault("grad", {})[step_index] = grad.detach().float().cpu()
        eps = cfg_combine(eps_c.detach(), eps_u, icfg["cfg_scale"])
        return ddim_guided_step(eps, control.detach(), grad.detach(), a_t, a_prev, 1.0, reciprocal_div).detach()
    with torch.no_grad():
        pair = unet_forward(sd, cfg, latents.expand(2, -1, -1, -1, -1), t, text, guidance_blocks=gb,
                            down_residuals=dpair, mid_residual=mpair)
        eps = cfg_combine(pair[[1]], pair[[0]], icfg["cfg_scale"])
        return ddim_guided_step(eps, latents, None, a_t, a_prev, 1.0, reciprocal_div).detach()


def sample_loop(sd, cfg, icfg: dict, latents: Tensor, text: Tensor, representation, reciprocal_div: bool = False,
                stats: Optional[dict] = None, max_steps: Optional[int] = None,
                controlnet: Optional[dict] = None) -> List[Tensor]:
    """utils/motionclone_functions.py:102-171 (sample_video) from prepared latents to final latents (VAE excluded)."""
    timesteps = uneven_timesteps(icfg["inference_steps"], icfg["guidance_steps"], icfg["guidance_scale"])
    acp = alphas_cumprod()
    per_step = []
    for i in range(len(timesteps) if max_steps is None else max_steps):
        latents = single_step(sd, cfg, icfg, latents, i, timesteps, acp, text, representation, reciprocal_div, stats,
                              controlnet)
        per_step.append(latents)
    return per_step
