"""torch.profiler breakdown of one guided + one plain DDIM step at the bench shapes (GPU). Not a bench value."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import motionclone_b200 as mc
from motionclone_b200.synthetic import UNET_SD15_CONFIG, synthetic_inputs
from bench import CONFIGS
INFER = {k: v for k, v in CONFIGS["object"].items() if k not in ("workload", "distinct_prompts")}

dev = torch.device("cuda:0")
infer = dict(INFER)
pipe = mc.build_pipeline(UNET_SD15_CONFIG, infer, device=dev)
inp = synthetic_inputs(16, 512, 512, 768, 42)
h = lambda t: t.to(dev, torch.float16)
pipe.set_prompt_embeds(h(inp["text_embeddings"]))
pipe.input_config["video_latents"], pipe.input_config["video_noise"] = h(inp["clip_latents"]), h(inp["clip_noise"])
pipe.obtain_motion_representation()
pipe.text_embeddings = h(inp["text_embeddings"]); pipe.motion_scale = 2000; pipe.add_controlnet = False
lat = h(inp["noisy_latents"])
ts = pipe.scheduler.timesteps_host
for _ in range(2):
    pipe.single_step_video(lat, 0, int(ts[0]), {}); pipe.single_step_video(lat, 30, int(ts[30]), {})
torch.cuda.synchronize()
for name, idx in (("guided", 0), ("plain", 30)):
    t0 = time.time(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): pipe.single_step_video(lat, idx, int(ts[idx]), {})
    e1.record(); torch.cuda.synchronize()
    print(f"{name} step: {e0.elapsed_time(e1)/3:.1f} ms GPU, {1000*(time.time()-t0)/3:.1f} ms wall")
from torch.profiler import profile, ProfilerActivity
for name, idx in (("guided", 0), ("plain", 30)):
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        pipe.single_step_video(lat, idx, int(ts[idx]), {}); torch.cuda.synchronize()
    ev = [e for e in prof.key_averages() if e.device_time_total > 0 and getattr(e, "device_type", None) is not None]
    rows = sorted(((e.self_device_time_total, e.count, e.key) for e in prof.key_averages() if e.self_device_time_total > 0), reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"==== {name}: total device time {tot/1e3:.1f} ms over {sum(r[1] for r in rows)} events")
    for t, c, k in rows[:45]:
        print(f"{t/1e3:9.2f} ms {100*t/tot:5.1f}% x{c:5d}  {k[:110]}")
