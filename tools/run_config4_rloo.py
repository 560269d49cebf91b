"""BASELINE.json configs[3]: SD3-medium 512^2 RLOO rollout (4 samples/prompt, batch 16 = 4 prompts x 4) with TimePredictor
fwd+bwd and an NCCL gradient all-reduce.  Launch:  python -m torch.distributed.run --nproc-per-node N tools/run_config4_rloo.py

Checks (every rank): the all-reduced flat gradient equals the sum of the per-rank gradients; parameters stay bit-identical
across ranks after the fused AdamW step.  Prints one JSON line with the timings."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpdm_b200.modeling_sd3_pnt import SD3_MEDIUM_TRANSFORMER_CONFIG, SD3PredictNextTimeStepModelRLOOWrapper  # noqa: E402
from tpdm_b200.rloo import rloo_advantages  # noqa: E402
from tpdm_b200.tpm_training import TimePredictorTrainer  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    small = "--small" in sys.argv
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(SD3_MEDIUM_TRANSFORMER_CONFIG, sample_size=64)
    if small:
        cfg.update(num_layers=2)
    torch.manual_seed(1234)                          # same weights on every rank (what the DDP broadcast would give)
    w = SD3PredictNextTimeStepModelRLOOWrapper(transformer_config=cfg, torch_dtype=torch.bfloat16, device=dev, min_sigma=0.01,
                                               init_alpha=2.5, init_beta=1.0, max_inference_steps=28)
    prompts, k = 4, 4
    g = torch.Generator().manual_seed(100 + rank)    # different prompts per rank (data parallel)
    data = dict(prompt_embeds=torch.randn(prompts, 333, 4096, generator=g).to(dev), negative_prompt_embeds=torch.randn(prompts, 333, 4096, generator=g).to(dev),
                pooled_prompt_embeds=torch.randn(prompts, 2048, generator=g).to(dev), negative_pooled_prompt_embeds=torch.randn(prompts, 2048, generator=g).to(dev))
    data = w.rloo_repeat(data, k)
    trainer = TimePredictorTrainer(w.agent_model.time_predictor, grid=32, max_samples=8 * 28, lr=1e-6)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    w.sample({**data, "predict": False, "generator": torch.Generator().manual_seed(1)})   # warm-up: weight packing, plan, first launches
    torch.cuda.synchronize()
    ev[0].record()
    out = w.sample({**data, "predict": False, "generator": torch.Generator().manual_seed(7 + rank * 100003)})
    ev[1].record()
    reward = -(out["latents"].float() ** 2).mean(dim=(1, 2, 3)).cpu()
    adv = rloo_advantages(reward, k).to(dev)
    x = out["hidden_states_combineds"].permute(0, 1, 3, 4, 2)
    idx = torch.arange(8, device=dev)                # one micro-batch of 8 rollouts
    st = trainer.ppo_update(out["sigmas"][idx], out["logprobs"][idx], x[idx], out["tembs"][idx], adv[idx], min_sigma=0.01, optimizer_step=False)
    torch.cuda.synchronize()
    # --- all-reduce check: recompute the local gradient, gather, compare with the reduced buffer
    reduced = trainer.grads.clone()
    if world > 1:
        ab = trainer.forward(x[idx].reshape(-1, 32, 32, 3072), out["tembs"][idx].reshape(-1, 1536))
        del ab
        # local gradient again (ppo_update already all-reduced trainer.grads in place)
        import ctypes as C
        from tpdm_b200 import _lib as L
        lib = L.load()
        mb, T = out["sigmas"][idx].shape
        new_lp, dz, stats = torch.empty(mb, T, device=dev), torch.empty(mb * T, 2, device=dev), torch.empty(4, device=dev)
        abv = trainer.forward(x[idx].reshape(-1, 32, 32, 3072), out["tembs"][idx].reshape(-1, 1536))
        sig, old, a = out["sigmas"][idx].float().contiguous(), out["logprobs"][idx].float().contiguous(), adv[idx].float().contiguous()
        L.check(lib.tpdm_ppo_clip_loss(L.ptr(abv), L.ptr(sig), L.ptr(old), L.ptr(a), mb, T, 0.01, 1e-3, 1, 0, 0.2, 1.0, L.ptr(new_lp), L.ptr(dz), L.ptr(stats), None, L.stream_ptr()))
        trainer.backward(dz)
        local_g = trainer.grads.clone()
        gathered = [torch.empty_like(local_g) for _ in range(world)]
        dist.all_gather(gathered, local_g)
        total = torch.stack(gathered).sum(0)
        err = float((total - reduced).norm() / (reduced.norm() + 1e-30))
        assert err < 1e-3, f"all-reduced gradient differs from the sum of per-rank gradients: {err}"   # atomics reorder sums
        trainer.grads.copy_(reduced)
    ev[2].record()
    gn = trainer.optimizer_step(grad_scale=1.0 / world)
    ev[3].record()
    torch.cuda.synchronize()
    if world > 1:
        ref = trainer.params.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, trainer.params), "parameters diverged across ranks after the optimizer step"
    if rank == 0:
        T = out["sigmas"].shape[1]
        print(json.dumps({"config": "SD3-medium 512^2 RLOO rollout, 4 prompts x rloo_k 4 per GPU, micro-batch 8" + (" [2 blocks]" if small else ""),
                          "n_gpus": world, "rollout_ms": ev[0].elapsed_time(ev[1]), "rollout_steps": T,
                          "rollouts_per_s_all_gpus": world * 16 / (ev[0].elapsed_time(ev[1]) / 1e3),
                          "ms_per_denoise_step": ev[0].elapsed_time(ev[1]) / T,
                          "mmdit_tflops": 67.44 * T / ev[0].elapsed_time(ev[1]) * 1e3 if not small else None,
                          "adamw_ms": ev[2].elapsed_time(ev[3]), "grad_norm": float(gn), "loss": float(st["loss"]),
                          "allreduce_bytes": trainer.grads.numel() * 4, "checks": "allreduce == sum of rank grads; params identical across ranks"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
