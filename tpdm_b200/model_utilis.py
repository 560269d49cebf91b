"""Drop-in for /root/reference/src/models/model_utilis.py: output containers and the flow-matching Euler step."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Optional, Tuple, Union

import torch

from . import _lib as L
from .transformer_sd3 import BaseOutput


@dataclass
class CustomFlowMatchEulerDiscreteSchedulerOutput(BaseOutput):
    """model_utilis.py:12-23."""

    prev_sample: torch.FloatTensor


@dataclass
class CustomDiffusionModelOutput(BaseOutput):
    """model_utilis.py:25-45 (same field names; dict- and attribute-style access)."""

    init_noise_latents: torch.Tensor
    hidden_states_combineds: Optional[torch.Tensor]
    tembs: torch.Tensor
    images: Any
    last_valid_indices: Optional[List[int]]
    alphas: torch.Tensor
    betas: torch.Tensor
    sigmas: torch.Tensor
    logprobs: torch.Tensor
    prob_masks: torch.Tensor
    latents: Optional[torch.Tensor] = None  # tpdm_b200 addition: final (last valid) latents, (B, C, h, w)
    steps: Optional[torch.Tensor] = None    # sample_queue: denoising steps each prompt took (0: processed by another GPU)
    device_steps: Optional[int] = None      # sample_queue: steps this GPU executed


class CustomFlowMatchEulerDiscreteScheduler:
    """Only ``custom_step`` (model_utilis.py:52-74) is on the TPDM path; the diffusers base class is not needed."""

    def __init__(self, *args, **kwargs):
        self.config = dict(kwargs)

    @classmethod
    def from_pretrained(cls, *args, **kwargs):
        return cls()

    def custom_step(
        self,
        model_output: torch.FloatTensor,
        sigma_next: torch.Tensor,
        sigma: torch.Tensor,
        sample: torch.FloatTensor,
        return_dict: bool = True,
    ) -> Union[CustomFlowMatchEulerDiscreteSchedulerOutput, Tuple]:
        """prev = (fp32(sample) + (sigma_next - sigma)[:, None, None, None] * model_output).to(model_output.dtype)"""
        lib = L.load()
        if not model_output.is_cuda:
            raise RuntimeError("tpdm_b200.custom_step runs on CUDA only (no CPU path)")
        dev, f32 = model_output.device, torch.float32
        B = sample.shape[0]
        n = sample[0].numel()
        mo = model_output.to(f32).contiguous()
        x = sample.to(device=dev, dtype=f32).contiguous()
        sn = torch.as_tensor(sigma_next, device=dev, dtype=f32).reshape(-1).expand(B).contiguous()
        s0 = torch.as_tensor(sigma, device=dev, dtype=f32).reshape(-1).expand(B).contiguous()
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            L.check(lib.tpdm_euler_step(L.ptr(mo), L.ptr(sn), L.ptr(s0), L.ptr(x), L.ptr(out), B, n, L.stream_ptr()))
        prev_sample = out.to(model_output.dtype)
        if not return_dict:
            return (prev_sample,)
        return CustomFlowMatchEulerDiscreteSchedulerOutput(prev_sample=prev_sample)
