"""Workload for compute-sanitizer: the tiny config (2 joint blocks, hidden 384, 4 heads of 96, 256^2) through
every kernel family of the library -- adaptive trajectory (predict and device-side Beta draws), device-side prompt queue, TimePredictor
forward + backward + PPO loss + AdamW, VAE decode of the result -- each once, with small shapes so that the ~50x slowdown of the tool stays
affordable.  Run on the GPU box:
    compute-sanitizer --tool memcheck  --log-file profiles/r02_sanitizer_memcheck.txt  python tools/sanitize_tiny.py
    compute-sanitizer --tool racecheck --log-file profiles/r02_sanitizer_racecheck.txt python tools/sanitize_tiny.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpdm_b200.modeling_sd3_pnt import SD3PredictNextTimeStepModelRLOOWrapper  # noqa: E402
from tpdm_b200.rloo import rloo_update  # noqa: E402
from tpdm_b200.tpm_training import TimePredictorTrainer  # noqa: E402

tiny = dict(sample_size=32, patch_size=2, in_channels=16, num_layers=2, attention_head_dim=96, num_attention_heads=4,
            joint_attention_dim=4096, caption_projection_dim=384, pooled_projection_dim=2048, out_channels=16, pos_embed_max_size=96)
torch.manual_seed(0)
w = SD3PredictNextTimeStepModelRLOOWrapper(transformer_config=tiny, torch_dtype=torch.float32, device="cuda", min_sigma=0.05, max_inference_steps=4)
model = w.agent_model
g = torch.Generator().manual_seed(1)
mk = lambda *s: torch.randn(*s, generator=g).cuda()
data = dict(prompt_embeds=mk(2, 333, 4096), negative_prompt_embeds=mk(2, 333, 4096), pooled_prompt_embeds=mk(2, 2048),
            negative_pooled_prompt_embeds=mk(2, 2048))
lat = mk(2, 16, 32, 32)
out = model(**data, latents=lat, max_inference_steps=4, predict=True)
print("predict trajectory:", tuple(out.sigmas.shape), float(out.latents.abs().mean()))
out = model(**data, latents=lat, max_inference_steps=3, predict=False, generator=torch.Generator().manual_seed(2))
print("sampled trajectory:", tuple(out.sigmas.shape))
q = model.sample_queue(data["prompt_embeds"], data["negative_prompt_embeds"], data["pooled_prompt_embeds"], data["negative_pooled_prompt_embeds"],
                       latents=lat, slots=1, max_inference_steps=3, use_graph=False)
print("queue:", q.steps.tolist())
trainer = TimePredictorTrainer(model.time_predictor, grid=16, max_samples=2 * 4, lr=1e-3)
res = rloo_update(w, trainer, data, reward_fn=lambda latents, o: -(latents.float() ** 2).mean(dim=(1, 2, 3)), rloo_k=2, num_ppo_epochs=1,
                  micro_batch_size=2)
print("rloo update:", res["logs"][-1])
lp = w.logprobs(None, w.subset_outputs(w.sample({**w.rloo_repeat(dict(data), 1), "predict": False}), torch.tensor([0], device="cuda")))
lp.sum().backward()
print("autograd replay grad norm:", float(model.time_predictor.fc2.weight.grad.norm()))
from oracle import vae_oracle as V  # noqa: E402  (only for the tiny decoder's topology numbers)
from tpdm_b200.vae import AutoencoderKL  # noqa: E402

vc = V.tiny_vae_config()
vae = AutoencoderKL(block_out_channels=vc.block_out_channels, layers_per_block=vc.layers_per_block, norm_num_groups=vc.norm_num_groups,
                    device="cuda", dtype=torch.float32)
img = vae.decode_latents(out.latents.float(), "pt")
torch.cuda.synchronize()
print("vae decode:", tuple(img.shape), "finite:", bool(torch.isfinite(img).all()))
print("sanitize workload done")
