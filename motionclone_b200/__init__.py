"""motionclone_b200 — B200-native (sm_100a) implementation of MotionClone's guided video-diffusion denoising loop,
behind the reference's own Python surface (VersatileAttention / CrossAttention / the nine bound functions).
See DESIGN.md; the C ABI underneath is include/motionclone_b200.h."""
from . import _lib  # noqa: F401

__all__ = ["build_pipeline"]


def build_pipeline(unet_config: dict, infer_config: dict, device="cuda", dtype=None, weight_seed: int = 42,
                   state_dict=None):
    """UNet3D (synthetic or given weights) + DDIMScheduler + AnimationPipeline with the nine functions bound
    (what t2v_video_sample.py:36-73 does)."""
    import torch

    from .guidance import bind_motionclone
    from .pipeline import AnimationPipeline, DDIMScheduler
    from .synthetic import NOISE_SCHEDULER_KWARGS, load_synthetic_weights
    from .unet3d import UNet3DConditionModel, _Config

    dtype = dtype or torch.float16
    unet = UNet3DConditionModel(**unet_config)
    if state_dict is not None:
        unet.load_state_dict(state_dict, strict=False)  # pos_encoder.pe is non-persistent (util.py:137)
    else:
        load_synthetic_weights(unet, weight_seed)
    unet = unet.to(device=device, dtype=dtype).to(memory_format=torch.channels_last).eval()
    pipe = AnimationPipeline(unet=unet, scheduler=DDIMScheduler(**NOISE_SCHEDULER_KWARGS))
    return bind_motionclone(pipe, _Config(dict(infer_config)))
