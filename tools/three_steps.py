"""Three denoising steps of the SD3-medium 1024^2 trajectory (what `ncu --metrics gpu__time_duration.sum` lists per launch:
profiles/r02_launch_shares.txt).  Same workload as bench.py, cut short so that the launch list stays readable."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpdm_b200.modeling_sd3_pnt import SD3_MEDIUM_TRANSFORMER_CONFIG, SD3PredictNextTimeStepModel  # noqa: E402

torch.manual_seed(1234)
model = SD3PredictNextTimeStepModel(transformer_config=SD3_MEDIUM_TRANSFORMER_CONFIG, torch_dtype=torch.bfloat16, device="cuda")
g = torch.Generator().manual_seed(0)
mk = lambda *s: torch.randn(*s, generator=g).cuda()
kw = dict(prompt_embeds=mk(1, 333, 4096), negative_prompt_embeds=mk(1, 333, 4096), pooled_prompt_embeds=mk(1, 2048),
          negative_pooled_prompt_embeds=mk(1, 2048), latents=mk(1, 16, 128, 128))
out = model(**kw, max_inference_steps=3, predict=True)
torch.cuda.synchronize()
print("steps", out.sigmas.shape[1], "sigma", out.sigmas.tolist())
