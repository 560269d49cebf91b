"""Training half of the TPDM path on top of libtpdm_b200.so: TimePredictor replay with
gradients, PPO-clip loss on summed log-probs, one flat-buffer gradient all-reduce over NCCL and a fused clip + AdamW
step.  Mirrors what CommonRLOOTrainer does for one micro-batch (/root/reference/src/train/rloo_trainer.py:485-523) for
the only trainable module, the TimePredictor (modeling_sd3_pnt.py:760-763)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib as L

NAMES = ("conv1.weight", "conv1.bias", "norm1.linear.weight", "norm1.linear.bias", "norm1.norm.weight", "norm1.norm.bias",
         "conv2.weight", "conv2.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")


def _to_packed(name: str, t: torch.Tensor) -> torch.Tensor:
    if name == "conv1.weight":   # [C1, 2D, 3, 3] -> [C1, 9, 2D]
        return t.permute(0, 2, 3, 1).reshape(-1)
    if name == "conv2.weight":   # [oc, c, 3, 3] -> [9, c, oc]
        return t.permute(2, 3, 1, 0).reshape(-1)
    return t.reshape(-1)


def _from_packed(name: str, flat: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    if name == "conv1.weight":
        c1, c_in = like.shape[0], like.shape[1]
        return flat.reshape(c1, 3, 3, c_in).permute(0, 3, 1, 2)
    if name == "conv2.weight":
        oc, c = like.shape[0], like.shape[1]
        return flat.reshape(3, 3, c, oc).permute(3, 2, 0, 1)
    return flat.reshape(like.shape)


class TimePredictorTrainer:
    """Owns the flat fp32 master copy of the TimePredictor parameters, their gradients and the AdamW state."""

    def __init__(self, time_predictor, grid: int, max_samples: int, lr: float = 1e-6, betas=(0.9, 0.99), eps: float = 1e-5,
                 weight_decay: float = 0.0, max_grad_norm: float = 1.0):
        lib = L.load()
        self.module = time_predictor
        sd = time_predictor.state_dict()
        dev = sd["fc2.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("TimePredictorTrainer runs on CUDA only (no CPU path)")
        self.device, self.g, self.max_samples = dev, grid, max_samples
        self.C1 = sd["conv1.weight"].shape[0]
        self.D = sd["conv1.weight"].shape[1] // 2
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        off = (C.c_longlong * 13)()
        L.check(lib.tpdm_tpm_param_offsets(self.D, self.C1, off))
        self.off = list(off)
        n = self.off[12]
        self.params = torch.zeros(n, device=dev, dtype=torch.float32)
        self.grads = torch.zeros_like(self.params)
        self.m, self.v = torch.zeros_like(self.params), torch.zeros_like(self.params)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.n_conv1 = sd["conv1.weight"].numel()
        self.conv1_bf16 = torch.zeros(self.n_conv1, device=dev, dtype=torch.bfloat16)
        self.load_from_module()
        with torch.cuda.device(dev):
            nbytes = lib.tpdm_tpm_trainer_workspace_bytes(self.D, self.C1, grid, max_samples)
            self.workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
            base = (self.workspace.data_ptr() + 1023) // 1024 * 1024
            h = L.vp()
            L.check(lib.tpdm_tpm_trainer_create(self.D, self.C1, grid, max_samples, float(time_predictor.epsilon), base, nbytes, C.byref(h)))
            self.handle = h
            L.check(lib.tpdm_tpm_trainer_bind(h, L.ptr(self.params), L.ptr(self.grads), L.ptr(self.conv1_bf16)))
        self.step_count = 0
        self._keep = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                L.load().tpdm_tpm_trainer_destroy(self.handle)
        except Exception:
            pass

    # ---- parameter hand-over -------------------------------------------------------------------------------------
    def load_from_module(self):
        sd = self.module.state_dict()
        for i, name in enumerate(NAMES):
            flat = _to_packed(name, sd[name].detach().to(device=self.device, dtype=torch.float32))
            self.params[self.off[i]: self.off[i] + flat.numel()].copy_(flat)
        self.conv1_bf16.copy_(self.params[: self.n_conv1])

    def tensors(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        sd = self.module.state_dict()
        return {name: _from_packed(name, flat[self.off[i]: self.off[i] + sd[name].numel()], sd[name]) for i, name in enumerate(NAMES)}

    def sync_to_module(self):
        """Write the fp32 master parameters back into the nn.Module (PyTorch layouts, module dtype)."""
        new = self.tensors(self.params)
        with torch.no_grad():
            for name, p in self.module.state_dict().items():
                p.copy_(new[name])

    def grad_dict(self) -> Dict[str, torch.Tensor]:
        return {k: v.clone() for k, v in self.tensors(self.grads).items()}

    # ---- forward / backward ----------------------------------------------------------------------------------------
    def forward(self, x_nhwc: torch.Tensor, temb: torch.Tensor) -> torch.Tensor:
        """x_nhwc: (ns, g, g, 2D) bf16 contiguous (or any view whose storage is NHWC-contiguous); temb: (ns, D)."""
        lib = L.load()
        ns = x_nhwc.shape[0]
        x = x_nhwc.to(torch.bfloat16).contiguous()
        t = temb.to(device=self.device, dtype=torch.float32).contiguous()
        if x.shape != (ns, self.g, self.g, 2 * self.D) or t.shape != (ns, self.D):
            raise ValueError(f"expected x (ns,{self.g},{self.g},{2 * self.D}) and temb (ns,{self.D}); got {tuple(x.shape)}, {tuple(t.shape)}")
        self._keep = (x, t)   # borrowed by the library until backward() has run
        ab = torch.empty(ns, 2, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(lib.tpdm_tpm_train_forward(self.handle, L.ptr(x), L.ptr(t), ns, L.ptr(ab), L.stream_ptr()))
        return ab

    def backward(self, dz: torch.Tensor) -> None:
        lib = L.load()
        d = dz.to(device=self.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self.device):
            L.check(lib.tpdm_tpm_train_backward(self.handle, L.ptr(d), L.stream_ptr()))

    # ---- one micro-batch of rloo_trainer.py:485-523 ------------------------------------------------------------------
    def ppo_update(self, sigmas: torch.Tensor, old_logprobs: torch.Tensor, tpm_inputs_nhwc: torch.Tensor, tembs: torch.Tensor,
                   advantages: torch.Tensor, min_sigma: float, cliprange: float = 0.2, epsilon: float = 1e-3, relative: bool = True,
                   optimizer_step: bool = True) -> Dict[str, torch.Tensor]:
        """sigmas / old_logprobs (mb, T); tpm_inputs_nhwc (mb, T, g, g, 2D) bf16; tembs (mb, T, D); advantages (mb,)."""
        lib = L.load()
        mb, T = sigmas.shape
        f32 = dict(device=self.device, dtype=torch.float32)
        ab = self.forward(tpm_inputs_nhwc.reshape(mb * T, self.g, self.g, 2 * self.D), tembs.reshape(mb * T, self.D))
        sig, old, adv = sigmas.to(**f32).contiguous(), old_logprobs.to(**f32).contiguous(), advantages.to(**f32).contiguous()
        new_lp = torch.empty(mb, T, **f32)
        dz = torch.empty(mb * T, 2, **f32)
        stats = torch.empty(4, **f32)
        with torch.cuda.device(self.device):
            L.check(lib.tpdm_ppo_clip_loss(L.ptr(ab), L.ptr(sig), L.ptr(old), L.ptr(adv), mb, T, float(min_sigma), float(epsilon),
                                           1 if relative else 0, float(cliprange), float(self.module.epsilon), L.ptr(new_lp), L.ptr(dz),
                                           L.ptr(stats), L.stream_ptr()))
        self.backward(dz)
        world = dist.get_world_size() if dist.is_initialized() else 1
        if world > 1:   # the only exchange step of the path: TPM gradients, one flat buffer (rloo_trainer.py:501 under DDP / ZeRO-0)
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM)
        out = dict(loss=stats[0], clipfrac=stats[1], approxkl=stats[2], ratio=stats[3], new_logprobs=new_lp)
        if optimizer_step:
            out["grad_norm"] = self.optimizer_step(grad_scale=1.0 / world)
        return out

    def optimizer_step(self, grad_scale: float = 1.0) -> torch.Tensor:
        lib = L.load()
        self.step_count += 1
        with torch.cuda.device(self.device):
            L.check(lib.tpdm_adamw_step(L.ptr(self.params), L.ptr(self.grads), L.ptr(self.m), L.ptr(self.v), self.params.numel(), float(self.lr),
                                        float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                                        float(self.max_grad_norm), self.step_count, float(grad_scale), L.ptr(self.sumsq),
                                        L.ptr(self.conv1_bf16), self.n_conv1, L.stream_ptr()))
        return self.sumsq.sqrt().float()
