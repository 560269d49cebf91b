"""BASELINE.json configs[2]: SD3-medium 1024^2, 64 synthetic prompts sharded across the GPUs of one box with variable-length
adaptive trajectories (load-balance stress).  Launch:
    python -m torch.distributed.run --nproc-per-node N tools/run_config3_queue.py [--prompts 64] [--static]

Ranks claim prompts from one global ticket counter (tpdm_b200.work_queue.PromptQueue, an atomic fetch-add in the
torch.distributed store -- no collective on the data path); --static uses a round-robin split instead.  To get a spread of
trajectory lengths from random-init weights the TimePredictor's fc / conv2 weights are scaled up so that (alpha, beta)
depend on the hidden states (the reference init is bias dominated, SURVEY.md section 8d cfg 3)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpdm_b200.modeling_sd3_pnt import SD3_MEDIUM_TRANSFORMER_CONFIG, SD3PredictNextTimeStepModel  # noqa: E402
from tpdm_b200.work_queue import gather_results, max_over_ranks, sample_prompts  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    n_prompts = int(sys.argv[sys.argv.index("--prompts") + 1]) if "--prompts" in sys.argv else 64
    dynamic = "--static" not in sys.argv
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)
    model = SD3PredictNextTimeStepModel(transformer_config=SD3_MEDIUM_TRANSFORMER_CONFIG, torch_dtype=torch.bfloat16, device=dev,
                                        init_alpha=1.5, init_beta=0.5)
    with torch.no_grad():
        tp = model.time_predictor
        sc = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 4.0
        tp.fc2.weight.mul_(sc)
        tp.fc1.weight.mul_(4.0)
        tp.conv2.weight.mul_(2.0)
    model.get_engine()

    def inputs(i):
        g = torch.Generator().manual_seed(5000 + i)
        mk = lambda *s: torch.randn(*s, generator=g).to(dev)
        return dict(prompt_embeds=mk(1, 333, 4096), negative_prompt_embeds=mk(1, 333, 4096), pooled_prompt_embeds=mk(1, 2048),
                    negative_pooled_prompt_embeds=mk(1, 2048), latents=mk(1, 16, 128, 128))

    model(**inputs(10_000 + rank), max_inference_steps=28, predict=True)   # warm-up (plan, kernels)
    busy = [0.0]

    def run_one(i):
        t0 = time.perf_counter()
        out = model(**inputs(i), max_inference_steps=28, predict=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        busy[0] += dt
        return {"steps": int(out.sigmas.shape[1]), "rank": rank, "seconds": dt}

    if "--device-queue" in sys.argv:
        # every GPU holds all prompts and drains ONE ticket counter (CUDA IPC, system-scope atomics) with `slots` prompts in flight
        from tpdm_b200.work_queue import SharedTicket

        slots = int(sys.argv[sys.argv.index("--slots") + 1]) if "--slots" in sys.argv else 2
        allin = [inputs(i) for i in range(n_prompts)]
        cat = {k: torch.cat([a[k] for a in allin]) for k in allin[0]}
        del allin
        ticket = SharedTicket.create()
        model.sample_queue(cat["prompt_embeds"][:slots], cat["negative_prompt_embeds"][:slots], cat["pooled_prompt_embeds"][:slots],
                           cat["negative_pooled_prompt_embeds"][:slots], latents=cat["latents"][:slots], slots=slots, max_inference_steps=28)
        ticket.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = model.sample_queue(cat["prompt_embeds"], cat["negative_prompt_embeds"], cat["pooled_prompt_embeds"],
                                 cat["negative_pooled_prompt_embeds"], latents=cat["latents"], slots=slots, max_inference_steps=28, ticket=ticket,
                                 use_graph="--no-graph" not in sys.argv)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        makespan = max_over_ranks(wall, dev)
        steps = out.steps.clone()
        if world > 1:
            dist.all_reduce(steps)                 # each prompt was processed by exactly one rank
            counts = [torch.zeros_like(out.steps) for _ in range(world)]
            dist.all_gather(counts, (out.steps > 0).int())
            per_rank = [int(c.sum()) for c in counts]
            dsteps = [None] * world
            dist.all_gather_object(dsteps, out.device_steps)
        else:
            per_rank, dsteps = [int((out.steps > 0).sum())], [out.device_steps]
        if rank == 0:
            st = steps.tolist()
            assert all(s > 0 for s in st), "a prompt was not processed"
            assert sum(per_rank) == n_prompts, "a prompt was processed twice"
            hist = {}
            for s_ in st:
                hist[s_] = hist.get(s_, 0) + 1
            print(json.dumps({"config": f"SD3-medium 1024^2, {n_prompts} prompts, device-side queue, {slots} slots per GPU, shared ticket",
                              "n_gpus": world, "makespan_s": makespan, "images_per_s": n_prompts / makespan, "prompts_per_gpu": per_rank,
                              "device_steps_per_gpu": dsteps, "total_steps": sum(st), "steps_histogram": dict(sorted(hist.items())),
                              "ideal_device_steps": sum(st) / (world * slots)}))
        if world > 1:
            dist.destroy_process_group()
        return
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mine = sample_prompts(run_one, n_prompts, dynamic=dynamic, name="cfg3")
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    makespan = max_over_ranks(wall, dev)
    merged = gather_results(mine)
    busy_all = [None] * world
    if world > 1:
        dist.all_gather_object(busy_all, busy[0])
    else:
        busy_all = [busy[0]]
    if rank == 0:
        steps = [merged[i]["steps"] for i in sorted(merged)]
        secs = sum(merged[i]["seconds"] for i in merged)
        hist = {}
        for s in steps:
            hist[s] = hist.get(s, 0) + 1
        per_rank = [sum(1 for i in merged if merged[i]["rank"] == r) for r in range(world)]
        print(json.dumps({"config": f"SD3-medium 1024^2, {n_prompts} prompts, {'ticket queue' if dynamic else 'static round-robin'}", "n_gpus": world,
                          "makespan_s": makespan, "images_per_s": n_prompts / makespan, "lower_bound_s": secs / world,
                          "efficiency_vs_lower_bound": (secs / world) / makespan, "busy_fraction_per_gpu": [b / makespan for b in busy_all],
                          "prompts_per_gpu": per_rank, "steps_histogram": dict(sorted(hist.items())), "total_steps": sum(steps)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
