"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/ref_runner.py). TEST INFRASTRUCTURE.

Run in the build container only:  python oracle/gen_golden.py [tiny8 tiny16 c1 c2mini ...]
The fixtures are committed; the GPU box never needs /root/reference.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motionclone_b200.synthetic import (SPARSECTRL_IMAGE_KWARGS, SPARSECTRL_LATENT_KWARGS, UNET_SD15_CONFIG,  # noqa: E402
                                        UNET_TINY_CONFIG, synthetic_condition, synthetic_inputs)
from oracle.ref_runner import run_reference  # noqa: E402

BASE = dict(cfg_scale=7.5, negative_prompt="", warm_up_steps=10, cool_up_steps=10, motion_guidance_weight=2000,
            motion_guidance_blocks=["up_blocks.1"], add_noise_step=400)

CASES = {
    # name: (unet config name, inference cfg, input seed)
    "tiny8": ("tiny", dict(BASE, inference_steps=6, guidance_steps=3, guidance_scale=0.3, video_length=8, height=128,
                           width=128), 42),
    "tiny16": ("tiny", dict(BASE, inference_steps=5, guidance_steps=3, guidance_scale=0.4, video_length=16, height=128,
                            width=128, warm_up_steps=2, cool_up_steps=2), 52),
    # SparseCtrl (BASELINE.json configs[3], [4] topology at tiny widths): latent condition (i2v_rgb) / image condition (i2v_sketch)
    "tiny8_i2v_latent": ("tiny", dict(BASE, inference_steps=4, guidance_steps=2, guidance_scale=0.3, video_length=8, height=128,
                                      width=128, image_index=[0], controlnet_scale=1.0, sparsectrl="latent"), 62),
    "tiny8_i2v_image": ("tiny", dict(BASE, inference_steps=4, guidance_steps=2, guidance_scale=0.3, video_length=8, height=128,
                                     width=128, image_index=[0, 5], controlnet_scale=0.8, sparsectrl="image"), 72),
    # BASELINE.json configs[0]: t2v_camera, 8x256x256, 10 DDIM steps, SD1.5 widths (plumbing case, CPU-runnable)
    "c1": ("sd15", dict(BASE, inference_steps=10, guidance_steps=5, guidance_scale=0.3, video_length=8, height=256,
                        width=256), 42),
    # BASELINE.json configs[1] topology (t2v_object: 16 frames, SD1.5 + motion-module widths, guidance_scale 0.4) at 128x128
    # pixels and 4 DDIM steps: the head dims (40 / 80 / 160) and the frame count the bench config runs, CPU-runnable
    "c2mini": ("sd15", dict(BASE, inference_steps=4, guidance_steps=2, guidance_scale=0.4, video_length=16, height=128,
                            width=128, warm_up_steps=2, cool_up_steps=2), 82),
}


def main(names):
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name in names:
        ucfg_name, icfg, seed = CASES[name]
        ucfg = UNET_TINY_CONFIG if ucfg_name == "tiny" else UNET_SD15_CONFIG
        inp = synthetic_inputs(icfg["video_length"], icfg["height"], icfg["width"], ucfg["cross_attention_dim"], seed)
        t0 = time.time()
        cn_kwargs = None
        if icfg.get("sparsectrl"):
            kind = icfg["sparsectrl"]
            cn_kwargs = SPARSECTRL_LATENT_KWARGS if kind == "latent" else SPARSECTRL_IMAGE_KWARGS
            cond = synthetic_condition(kind, len(icfg["image_index"]), icfg["height"], icfg["width"], icfg["video_length"],
                                       seed + 5)
            inp.update(cond)
            paths = []
            if kind == "image":  # the reference opens image files (motionclone_functions.py:117): write lossless PNGs
                from PIL import Image
                for j, im in enumerate(cond["cond_images_u8"]):
                    pth = f"/tmp/_golden_{name}_cond{j}.png"
                    Image.fromarray(im.permute(1, 2, 0).numpy()).save(pth)
                    paths.append(pth)
            else:  # latent kind: the stub VAE returns the preset latents; the files only need to exist and open
                from PIL import Image
                for j in range(len(icfg["image_index"])):
                    pth = f"/tmp/_golden_{name}_cond{j}.png"
                    Image.fromarray(np.zeros((icfg["height"], icfg["width"], 3), dtype=np.uint8)).save(pth)
                    paths.append(pth)
            icfg = dict(icfg, condition_image_path_list=paths)
        out, pipe = run_reference(ucfg, icfg, inp, f"/tmp/_golden_{name}.pt", weight_seed=42, controlnet_kwargs=cn_kwargs)
        icfg = {k: v for k, v in icfg.items() if k != "condition_image_path_list"}
        arrays = {k: (v.numpy() if torch.is_tensor(v) else np.array(v)) for k, v in out.items()}
        arrays["meta"] = np.array(json.dumps(dict(case=name, unet=ucfg_name, infer=icfg, input_seed=seed,
                                                   weight_seed=42, torch=torch.__version__,
                                                   generator="oracle/gen_golden.py", reference="/root/reference @7724ee8",
                                                   seconds=round(time.time() - t0, 1))))
        if name == "c1":  # keep the fixture small: drop the full prob tensor and intermediate latents
            arrays.pop("extract_probs_0", None)
            arrays["latents_per_step"] = arrays["latents_per_step"][[0, 4, 5, 9]]
            arrays["latents_steps_kept"] = np.array([0, 4, 5, 9])
        path = os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz")
        if os.path.exists(path) and os.environ.get("GOLDEN_CHECK_STABLE", "1") == "1":
            old = np.load(path)  # regenerating must reproduce every tensor already committed, bit for bit
            for k in old.files:
                if k != "meta" and k in arrays:
                    assert np.array_equal(old[k], arrays[k]), f"{name}: regenerated '{k}' differs from the committed fixture"
        np.savez_compressed(path, **arrays)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB", round(time.time() - t0, 1), "s", flush=True)
        if ucfg_name == "tiny":
            shapes = {k: list(v.shape) for k, v in pipe.unet.state_dict().items()}
            with open(os.path.join(ROOT, "tests", "golden", f"ref_state_dict_shapes_{ucfg_name}.json"), "w") as f:
                json.dump(shapes, f, indent=0)
        if pipe.controlnet is not None:
            shapes = {k: list(v.shape) for k, v in pipe.controlnet.state_dict().items()}
            with open(os.path.join(ROOT, "tests", "golden", f"ref_state_dict_shapes_controlnet_{icfg['sparsectrl']}.json"), "w") as f:
                json.dump(shapes, f, indent=0)
        del pipe, out


if __name__ == "__main__":
    main(sys.argv[1:] or ["tiny8", "tiny16"])
