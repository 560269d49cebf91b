"""Prompt sharding for multi-GPU sampling.

Trajectories are independent per prompt and TPDM makes their lengths differ, so a static split leaves GPUs idle.  Ranks
instead claim prompt indices from one global ticket counter.  The counter lives in the torch.distributed key-value store
(``store.add`` is an atomic fetch-add served by rank 0's TCPStore), i.e. there is no collective on the data path and a
rank never waits for another rank.  Works identically under the gloo and nccl backends and without torch.distributed
(single process)."""
from __future__ import annotations

from typing import Callable, Dict, Iterator, List, Optional

import torch
import torch.distributed as dist


class PromptQueue:
    """Contract: every rank constructs its PromptQueue objects of a given ``name`` in the same order (as with any collective).
    The n-th queue of that name counts under its own store key ``<name>/<n>``, so a second ``sample_prompts`` call in the same
    process group starts from zero again instead of inheriting the drained counter of the first."""

    _epochs: Dict[str, int] = {}

    def __init__(self, n_prompts: int, name: str = "tpdm_prompt_ticket", store=None):
        self.n = int(n_prompts)
        epoch = PromptQueue._epochs.get(name, 0)
        PromptQueue._epochs[name] = epoch + 1
        self.key = f"{name}/{epoch}"
        self._local = 0
        self.store = store
        if store is None and dist.is_available() and dist.is_initialized():
            self.store = dist.distributed_c10d._get_default_store()

    def claim(self) -> Optional[int]:
        """Next unclaimed prompt index, or None when the queue is drained."""
        if self.store is None:
            idx, self._local = self._local, self._local + 1
        else:
            idx = int(self.store.add(self.key, 1)) - 1
        return idx if idx < self.n else None

    def __iter__(self) -> Iterator[int]:
        while True:
            idx = self.claim()
            if idx is None:
                return
            yield idx


def expected_remaining_steps(sigma_after_first_step: torch.Tensor, min_sigma: float, max_steps: int) -> torch.Tensor:
    """Length estimate behind the longest-expected-first ticket order of the device queue (Engine._probe_and_order): a prompt whose
    first step took sigma from 1 to s keeps multiplying by about s, so it needs ceil(log(min_sigma / s) / log(s)) more steps until
    sigma_next < min_sigma (modeling_sd3_pnt.py:608), at most max_steps - 1; -1 marks a trajectory that already ended at its first step."""
    import math

    s = sigma_after_first_step.float()
    done = s < min_sigma
    r = s.clamp(1e-6, 1 - 1e-6)
    est = torch.ceil(math.log(min_sigma) / torch.log(r) - 1.0).clamp(0, max_steps - 1)
    return torch.where(done, torch.full_like(est, -1.0), est)


def longest_first_order(estimates: torch.Tensor) -> torch.Tensor:
    """Ticket t -> prompt id, longest expected trajectory first (stable, so every rank computes the same table); finished prompts last."""
    return torch.argsort(estimates, descending=True, stable=True).to(torch.int32)


def static_shard(n_prompts: int, rank: int, world: int) -> List[int]:
    """Round-robin split (what the bench's weak-scaling run uses: every rank owns prompts rank, rank+world, ...)."""
    return list(range(rank, n_prompts, world))


def sample_prompts(run_one: Callable[[int], Dict], n_prompts: int, dynamic: bool = True, name: str = "tpdm_prompt_ticket") -> Dict[int, Dict]:
    """Runs ``run_one(prompt_index)`` for this rank's share of ``n_prompts`` and returns {index: result}."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    indices = PromptQueue(n_prompts, name) if dynamic else static_shard(n_prompts, rank, world)
    return {i: run_one(i) for i in indices}


def gather_results(local: Dict[int, Dict]) -> Optional[Dict[int, Dict]]:
    """Bookkeeping only (step counts, timings): gathers the per-rank dicts on rank 0."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, out, dst=0)
    if dist.get_rank() != 0:
        return None
    merged: Dict[int, Dict] = {}
    for d in out:
        merged.update(d)
    return merged


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class SharedTicket:
    """One int32 ticket counter in the memory of local rank 0's GPU, mapped into every process of the box through CUDA IPC.
    `tpdm_queue_*` claims prompts from it with system-scope atomics over NVLink, so the GPUs of one node drain a single
    prompt list with no host round trip and no collective.  Usage (all ranks, same order):
        ticket = SharedTicket.create()        # collective: broadcasts the IPC handle through torch.distributed
        model.sample_queue(..., ticket=ticket)
    Needs peer access between the GPUs (NVLink / NVSwitch boxes); with world size 1 it is a plain device counter."""

    def __init__(self, ptr: int, owner: bool, opened=None):
        self._ptr, self._owner, self._opened = ptr, owner, opened

    def data_ptr(self) -> int:
        return self._ptr

    @classmethod
    def create(cls, group=None) -> "SharedTicket":
        from cuda.bindings import runtime as rt

        def ok(res):
            err, *rest = res
            if int(err) != 0:
                raise RuntimeError(f"CUDA runtime error {err} while setting up the shared ticket")
            return rest[0] if len(rest) == 1 else rest

        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        rank = dist.get_rank(group) if multi else 0
        if rank == 0:
            ptr = ok(rt.cudaMalloc(4))
            ok(rt.cudaMemset(ptr, 0, 4))
            ok(rt.cudaDeviceSynchronize())
            handle = bytes(ok(rt.cudaIpcGetMemHandle(ptr)).reserved) if multi else b""
        else:
            ptr, handle = 0, b""
        if not multi:
            return cls(int(ptr), True)
        box = [handle]
        dist.broadcast_object_list(box, src=0, group=group)
        if rank == 0:
            out = cls(int(ptr), True)
        else:
            h = rt.cudaIpcMemHandle_t()
            h.reserved = box[0]
            remote = ok(rt.cudaIpcOpenMemHandle(h, rt.cudaIpcMemLazyEnablePeerAccess))
            out = cls(int(remote), False, opened=remote)
        dist.barrier(group)
        return out

    def reset(self, group=None) -> None:
        """collective: the owner zeroes the counter, everybody waits"""
        from cuda.bindings import runtime as rt

        if self._owner:
            rt.cudaMemset(self._ptr, 0, 4)
            rt.cudaDeviceSynchronize()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.barrier(group)

    def value(self) -> int:
        from cuda.bindings import runtime as rt
        import numpy as np

        buf = np.zeros(1, dtype=np.int32)
        rt.cudaMemcpy(buf.ctypes.data, self._ptr, 4, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        return int(buf[0])
