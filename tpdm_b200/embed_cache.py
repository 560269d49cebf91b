"""On-disk cache of text-encoder outputs.

The reference either runs its three text towers in-process (``pre_process=False``,
/root/reference/src/models/stable_diffusion_3/modeling_sd3_pnt.py:162-178) or expects the caller to hand over
``prompt_embeds`` / ``pooled_prompt_embeds`` (``pre_process=True``, :464-483; ``rloo_repeat`` tiles them, :777-786).  The
towers are outside this package; this cache is the hand-over format: whoever owns the towers (an offline job, the
reference itself) stores each prompt's embeddings once, and the native pipeline's ``forward(prompt=...)`` resolves prompts
through it.

Layout: one ``<sha256(prompt)[:32]>.safetensors`` per prompt under ``root`` holding
  ``prompt_embeds`` (T, 4096)  -- CLIP-L/G hidden states padded to 4096 and concatenated with T5 along tokens (:289-296)
  ``pooled_prompt_embeds`` (2048,)
in the dtype they were produced in, with the prompt text in the safetensors metadata.  The negative ("") prompt is stored
like any other prompt."""
from __future__ import annotations

import hashlib
import os
from typing import Dict, Iterable, List, Sequence, Tuple

import torch
from safetensors import safe_open
from safetensors.torch import save_file


class PromptEmbeddingCache:
    def __init__(self, root: str):
        self.root = root
        os.makedirs(root, exist_ok=True)

    @staticmethod
    def key(prompt: str) -> str:
        return hashlib.sha256(prompt.encode("utf-8")).hexdigest()[:32]

    def path(self, prompt: str) -> str:
        return os.path.join(self.root, self.key(prompt) + ".safetensors")

    def __contains__(self, prompt: str) -> bool:
        return os.path.isfile(self.path(prompt))

    def put(self, prompt: str, prompt_embeds: torch.Tensor, pooled_prompt_embeds: torch.Tensor) -> None:
        """prompt_embeds (T, D) or (1, T, D); pooled (P,) or (1, P).  Written atomically (rename)."""
        pe = prompt_embeds.detach().cpu()
        pp = pooled_prompt_embeds.detach().cpu()
        pe = pe[0] if pe.dim() == 3 else pe
        pp = pp[0] if pp.dim() == 2 else pp
        if pe.dim() != 2 or pp.dim() != 1:
            raise ValueError(f"expected prompt_embeds (T, D) and pooled (P,), got {tuple(prompt_embeds.shape)} / {tuple(pooled_prompt_embeds.shape)}")
        tmp = self.path(prompt) + f".tmp{os.getpid()}"
        save_file({"prompt_embeds": pe.contiguous(), "pooled_prompt_embeds": pp.contiguous()}, tmp, metadata={"prompt": prompt})
        os.replace(tmp, self.path(prompt))

    def put_many(self, prompts: Sequence[str], prompt_embeds: torch.Tensor, pooled_prompt_embeds: torch.Tensor) -> None:
        for i, p in enumerate(prompts):
            self.put(p, prompt_embeds[i], pooled_prompt_embeds[i])

    def get(self, prompt: str) -> Tuple[torch.Tensor, torch.Tensor]:
        path = self.path(prompt)
        if not os.path.isfile(path):
            raise KeyError(f"prompt not in the embedding cache {self.root!r}: {prompt!r}")
        with safe_open(path, framework="pt") as f:
            stored = (f.metadata() or {}).get("prompt")
            if stored is not None and stored != prompt:
                raise KeyError(f"hash collision in the embedding cache for {prompt!r}")
            return f.get_tensor("prompt_embeds"), f.get_tensor("pooled_prompt_embeds")

    def get_batch(self, prompts: Iterable[str], device=None, dtype=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> prompt_embeds (B, T, D), pooled (B, P); all prompts must share T (the reference pads to a fixed length)"""
        pes, pps = zip(*(self.get(p) for p in prompts))
        if len({t.shape for t in pes}) != 1:
            raise ValueError("cached prompt_embeds have different token counts; store them padded to one length")
        return torch.stack(pes).to(device=device, dtype=dtype), torch.stack(pps).to(device=device, dtype=dtype)

    def prompts(self) -> List[str]:
        out = []
        for name in sorted(os.listdir(self.root)):
            if name.endswith(".safetensors"):
                with safe_open(os.path.join(self.root, name), framework="pt") as f:
                    out.append((f.metadata() or {}).get("prompt", ""))
        return out
