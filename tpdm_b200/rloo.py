"""Minimal RLOO update around the native pieces: the arithmetic of CommonRLOOTrainer.train for ONE update
(/root/reference/src/train/rloo_trainer.py:425-523) without the HF Trainer / trl / accelerate plumbing (out of scope,
SURVEY.md section 2.1 #6).  rollout (device-side Beta draws) -> reward -> RLOO advantage -> PPO epochs over micro-batches ->
native TPM backward + flat-buffer all-reduce + fused AdamW."""
from __future__ import annotations

from typing import Callable, Dict, List

import numpy as np
import torch

from .tpm_training import TimePredictorTrainer


def rloo_advantages(rlhf_reward: torch.Tensor, rloo_k: int) -> torch.Tensor:
    """rloo_trainer.py:458-461 (layout = rloo_k repeats x prompts, as rloo_repeat tiles the prompt list)."""
    r = rlhf_reward.reshape(rloo_k, -1)
    baseline = (r.sum(0) - r) / (rloo_k - 1)
    return (r - baseline).flatten()


def rloo_update(wrapper, trainer: TimePredictorTrainer, data: Dict, reward_fn: Callable, rloo_k: int = 2, num_ppo_epochs: int = 4,
                micro_batch_size: int = 8, cliprange: float = 0.2, gamma: float = 0.97, kl_coef: float = 0.0, seed: int = 0) -> Dict:
    """`wrapper`: SD3PredictNextTimeStepModelRLOOWrapper; `data`: dict with the four embedding tensors (and optionally
    'prompt'); `reward_fn(latents (B,C,h,w), outputs) -> (B,)` stands in for the reward model."""
    agent = wrapper.agent_model
    data = wrapper.rloo_repeat(dict(data), rloo_k)
    outputs = wrapper.sample({**data, "predict": False, "generator": torch.Generator().manual_seed(seed)})
    prob_masks = outputs["prob_masks"]
    last_reward = reward_fn(outputs["latents"], outputs).float().cpu()
    scores = []
    for i in range(prob_masks.shape[0]):       # discounted reward (modeling_sd3_pnt.py:838-841)
        last = int(outputs["last_valid_indices"][i])
        scores.append(sum(float(last_reward[i]) * gamma ** (last - j) for j in range(last + 1)) / (last + 1))
    scores = torch.tensor(scores)
    kl = wrapper.kl_divergence(outputs) if kl_coef != 0.0 else torch.zeros(scores.shape[0], 1)
    rlhf_reward = scores + (-kl_coef * kl).sum(1)
    advantages = rloo_advantages(rlhf_reward, rloo_k).to(agent.device)
    B = scores.shape[0]
    x = outputs["hidden_states_combineds"].permute(0, 1, 3, 4, 2)     # back to the NHWC storage it is a view of
    logs: List[Dict] = []
    rng = np.random.RandomState(seed)
    for _ in range(num_ppo_epochs):
        perm = rng.permutation(B)
        for s in range(0, B, micro_batch_size):
            idx = torch.as_tensor(perm[s: s + micro_batch_size], device=agent.device)
            st = trainer.ppo_update(outputs["sigmas"][idx], outputs["logprobs"][idx], x[idx], outputs["tembs"][idx], advantages[idx],
                                    min_sigma=agent.min_sigma, cliprange=cliprange, epsilon=agent.epsilon, relative=agent.relative)
            logs.append({k: float(v) for k, v in st.items() if k != "new_logprobs"})
    trainer.sync_to_module()
    return dict(scores=scores, advantages=advantages.cpu(), steps=(~prob_masks).sum(1).float().mean().item(), logs=logs)
