"""One guided + one plain DDIM step at the bench shapes (16 x 512 x 512), no warm-up: the short command behind the ncu
launch list in profiles/ (`ncu --metrics gpu__time_duration.sum ...`). Not a bench value."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import motionclone_b200 as mc
from motionclone_b200.synthetic import UNET_SD15_CONFIG, synthetic_inputs
from bench import CONFIGS
INFER = {k: v for k, v in CONFIGS['object'].items() if k not in ('workload', 'distinct_prompts')}

dev = torch.device("cuda:0")
# eager launches (no CUDA-graph capture: its warm-up forwards would triple the list), like bench.py's roofline leg
pipe = mc.build_pipeline(UNET_SD15_CONFIG, dict(INFER), device=dev, use_cuda_graphs=False)
inp = synthetic_inputs(16, 512, 512, 768, 42)
h = lambda t: t.to(dev, torch.float16)
pipe.set_prompt_embeds(h(inp["text_embeddings"]))
g = torch.Generator().manual_seed(0)
hw = (512 // 32) * (512 // 32)
names = [f"up_blocks.1.motion_modules.{i}.temporal_transformer.transformer_blocks.0.attention_blocks.{j}"
         for i in range(3) for j in range(2)]
# a synthetic motion representation of the right shape (launch list only: skips the extraction forward)
pipe.motion_representation_dict = {n: [torch.rand(hw, 8, 16, 1, generator=g).to(dev, torch.float16),
                                       torch.randint(0, 16, (hw, 8, 16, 1), generator=g).to(dev, torch.uint8)] for n in names}
pipe.text_embeddings = h(inp["text_embeddings"]); pipe.motion_scale = 2000; pipe.add_controlnet = False
lat = h(inp["noisy_latents"])
ts = pipe.scheduler.timesteps_host
torch.cuda.synchronize()
print("MARK steps begin", flush=True)
pipe.single_step_video(lat, 0, int(ts[0]), {})    # guided
pipe.single_step_video(lat, 30, int(ts[30]), {})  # plain
torch.cuda.synchronize()
print("MARK steps end", flush=True)
