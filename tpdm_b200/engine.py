"""Host-side engine: packs module weights for libtpdm_b200.so, owns shape-bound plans (workspace + prebuilt TMA
descriptors) and drives the adaptive loop.  PyTorch is used for device memory and streams only; every arithmetic
operation of the path runs inside the C-ABI library (no eager fallback exists).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import torch

from . import _lib as L


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(dtype=torch.float32).contiguous()


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(dtype=torch.bfloat16).contiguous()


def _pad_heads_rows(w: torch.Tensor, heads: int, d: int, dp: int) -> torch.Tensor:
    """[heads*d, ...] -> [heads*dp, ...] with zero rows for the padded head dims."""
    if d == dp:
        return w
    rest = w.shape[1:]
    out = w.new_zeros((heads, dp) + tuple(rest))
    out[:, :d] = w.reshape((heads, d) + tuple(rest))
    return out.reshape((heads * dp,) + tuple(rest))


class PackedWeights:
    """Device tensors in the layout include/tpdm_b200.h documents, plus the ctypes structs pointing at them."""

    def __init__(self):
        self.keep = []           # owning references
        self.by_name = {}
        self.struct = L.TpdmWeights()
        self.blocks = None

    def _set(self, struct, name, tensor):
        if struct is self.struct and name in self.by_name:   # refresh: same storage, same pointers (plans stay valid)
            self.by_name[name].copy_(tensor)
            return
        self.keep.append(tensor)
        if struct is self.struct:
            self.by_name[name] = tensor
        setattr(struct, name, tensor.data_ptr())


def pack_transformer(pw: PackedWeights, sd: Dict[str, torch.Tensor], cfg, device) -> None:
    """diffusers-layout state dict of CustomSD3Transformer2DModel -> tpdm_weights (MMDiT part)."""
    H, d = cfg.num_attention_heads, cfg.attention_head_dim
    D = H * d
    dp = 64 if d <= 64 else 128
    Lyr = cfg.num_layers
    dual_layers = set(int(i) for i in (getattr(cfg, "dual_attention_layers", ()) or ()))
    g = lambda k: sd[k].to(device)
    s = pw.struct
    pw._set(s, "patch_w", _f32(g("pos_embed.proj.weight").reshape(D, -1)))
    pw._set(s, "patch_b", _f32(g("pos_embed.proj.bias")))
    pw._set(s, "pos_table", _f32(g("pos_embed.pos_embed").reshape(-1, D)))
    for short, base in (("t", "time_text_embed.timestep_embedder"), ("p", "time_text_embed.text_embedder")):
        pw._set(s, f"{short}_w1", _f32(g(base + ".linear_1.weight")))
        pw._set(s, f"{short}_b1", _f32(g(base + ".linear_1.bias")))
        pw._set(s, f"{short}_w2", _f32(g(base + ".linear_2.weight")))
        pw._set(s, f"{short}_b2", _f32(g(base + ".linear_2.bias")))
    pw._set(s, "ctx_w", _bf16(g("context_embedder.weight")))
    pw._set(s, "ctx_b", _f32(g("context_embedder.bias")))
    pw._set(s, "proj_w", _bf16(g("proj_out.weight")))
    pw._set(s, "proj_b", _f32(g("proj_out.bias")))
    ada_w, ada_b = [], []
    blocks = (L.TpdmBlockWeights * Lyr)()
    for i in range(Lyr):
        p = f"transformer_blocks.{i}."
        last = i == Lyr - 1
        ada_w += [g(p + "norm1.linear.weight"), g(p + "norm1_context.linear.weight")]
        ada_b += [g(p + "norm1.linear.bias"), g(p + "norm1_context.linear.bias")]
        b = blocks[i]

        def fused(names, attn="attn"):
            w = torch.cat([_pad_heads_rows(g(p + f"{attn}.{n}.weight"), H, d, dp) for n in names], 0)
            bias = torch.cat([_pad_heads_rows(g(p + f"{attn}.{n}.bias"), H, d, dp) for n in names], 0)
            return _bf16(w), _f32(bias)

        def out_proj(name, attn="attn"):
            w = g(p + f"{attn}.{name}.weight")  # [D, H*d] -> [D, H*dp]
            if d != dp:
                w = _pad_heads_rows(w.t().contiguous(), H, d, dp).t().contiguous()
            return _bf16(w), _f32(g(p + f"{attn}.{name}.bias"))

        for field, (w, bias) in (("qkv", fused(("to_q", "to_k", "to_v"))), ("cqkv", fused(("add_q_proj", "add_k_proj", "add_v_proj"))),
                                 ("out", out_proj("to_out.0"))):
            pw._set(b, field + "_w", w)
            pw._set(b, field + "_b", bias)
        pw._set(b, "ff1_w", _bf16(g(p + "ff.net.0.proj.weight")))
        pw._set(b, "ff1_b", _f32(g(p + "ff.net.0.proj.bias")))
        pw._set(b, "ff2_w", _bf16(g(p + "ff.net.2.weight")))
        pw._set(b, "ff2_b", _f32(g(p + "ff.net.2.bias")))
        if not last:
            w, bias = out_proj("to_add_out")
            pw._set(b, "cout_w", w)
            pw._set(b, "cout_b", bias)
            pw._set(b, "cff1_w", _bf16(g(p + "ff_context.net.0.proj.weight")))
            pw._set(b, "cff1_b", _f32(g(p + "ff_context.net.0.proj.bias")))
            pw._set(b, "cff2_w", _bf16(g(p + "ff_context.net.2.weight")))
            pw._set(b, "cff2_b", _f32(g(p + "ff_context.net.2.bias")))
        dual = i in dual_layers
        if dual:  # SD3.5 attn2 (transformer_sd3.py:138): image-token self-attention with its own projections
            for field, (w, bias) in (("qkv2", fused(("to_q", "to_k", "to_v"), "attn2")), ("out2", out_proj("to_out.0", "attn2"))):
                pw._set(b, field + "_w", w)
                pw._set(b, field + "_b", bias)
        if cfg.qk_norm == "rms_norm":
            names = [("norm_q", "attn.norm_q"), ("norm_k", "attn.norm_k"), ("norm_added_q", "attn.norm_added_q"),
                     ("norm_added_k", "attn.norm_added_k")]
            if dual:
                names += [("norm_q2", "attn2.norm_q"), ("norm_k2", "attn2.norm_k")]
            for field, name in names:
                w = g(p + f"{name}.weight")
                wp = w.new_zeros(dp)
                wp[:d] = w
                pw._set(b, field, _f32(wp))
    ada_w.append(g("norm_out.linear.weight"))
    ada_b.append(g("norm_out.linear.bias"))
    pw._set(s, "adaln_w", _bf16(torch.cat(ada_w, 0)))
    pw._set(s, "adaln_b", _f32(torch.cat(ada_b, 0)))
    assert pw.keep[-1].numel() == 12 * D * Lyr - 2 * D + 3 * D * len(dual_layers)
    pw.blocks = blocks
    s.blocks = C.cast(blocks, C.POINTER(L.TpdmBlockWeights))


def pack_time_predictor(pw: PackedWeights, sd: Dict[str, torch.Tensor], device) -> None:
    """TimePredictor state dict (modeling_sd3_pnt.py:85-126 names) -> tpdm_weights (TPM part)."""
    g = lambda k: sd[k].to(device)
    s = pw.struct
    c1 = g("conv1.weight")  # [C1, 2D, 3, 3] -> [C1, 9, 2D]
    pw._set(s, "tpm_conv1_w", _bf16(c1.permute(0, 2, 3, 1).reshape(c1.shape[0], -1)))
    pw._set(s, "tpm_conv1_b", _f32(g("conv1.bias")))
    pw._set(s, "tpm_lin_w", _f32(g("norm1.linear.weight")))
    pw._set(s, "tpm_lin_b", _f32(g("norm1.linear.bias")))
    pw._set(s, "tpm_gn_w", _f32(g("norm1.norm.weight")))
    pw._set(s, "tpm_gn_b", _f32(g("norm1.norm.bias")))
    c2 = g("conv2.weight")  # [oc, c, 3, 3] -> [9, c, oc]
    pw._set(s, "tpm_conv2_w", _f32(c2.permute(2, 3, 1, 0).reshape(9, c2.shape[1], c2.shape[0])))
    pw._set(s, "tpm_conv2_b", _f32(g("conv2.bias")))
    pw._set(s, "tpm_fc1_w", _f32(g("fc1.weight")))
    pw._set(s, "tpm_fc1_b", _f32(g("fc1.bias")))
    pw._set(s, "tpm_fc2_w", _f32(g("fc2.weight")))
    pw._set(s, "tpm_fc2_b", _f32(g("fc2.bias")))


class Plan:
    """A shape-bound workspace with typed torch views over the library's state buffers."""

    def __init__(self, engine: "Engine", batch: int, cfg_pairs: bool, latent: int, n_text: int, max_steps: int):
        lib = L.load()
        self.engine, self.batch, self.cfg_pairs, self.latent, self.n_text, self.max_steps = engine, batch, cfg_pairs, latent, n_text, max_steps
        nbytes = lib.tpdm_plan_workspace_bytes(engine.ctx, batch, int(cfg_pairs), latent, latent, n_text, max_steps)
        if nbytes == 0:
            L.check(L.TPDM_ERR_SHAPE if lib.tpdm_last_error() else L.TPDM_ERR_ARG)
        self.workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=engine.device)
        base = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        self._base_off = base - self.workspace.data_ptr()
        handle = L.vp()
        L.check(lib.tpdm_plan_create(engine.ctx, batch, int(cfg_pairs), latent, latent, n_text, max_steps, base, nbytes, C.byref(handle)))
        self.handle = handle
        self.state = None
        if engine.has_mmdit and engine.has_tpm and cfg_pairs:
            st = L.TpdmSampleState()
            L.check(lib.tpdm_sample_state_get(handle, C.byref(st)))
            cfg, D, g = engine.cfg, engine.D, latent // 2
            Cc = cfg["in_channels"]
            f32, i32 = torch.float32, torch.int32
            self.state = dict(
                latents=self._view(st.latents, f32, (batch, Cc, latent, latent)),
                velocity=self._view(st.velocity, f32, (batch, Cc, latent, latent)),
                sigma_hist=self._view(st.sigma_hist, f32, (batch, max_steps + 1)),
                alphas=self._view(st.alphas, f32, (batch, max_steps)),
                betas=self._view(st.betas, f32, (batch, max_steps)),
                logprobs=self._view(st.logprobs, f32, (batch, max_steps)),
                prob_masks=self._view(st.prob_masks, i32, (batch, max_steps)),
                all_done=self._view(st.all_done, i32, (max_steps,)),
                tembs=self._view(st.tembs, f32, (max_steps, batch, D)),
                tpm_input=self._view(st.tpm_input, torch.bfloat16, (batch, g, g, 2 * D)),
                history_latents=self._view(st.history_latents, f32, (max_steps, batch, Cc, latent, latent)),
            )

    def _view(self, ptr: int, dtype: torch.dtype, shape) -> torch.Tensor:
        off = ptr - self.workspace.data_ptr()
        n = 1
        for s in shape:
            n *= s
        nb = n * torch.empty((), dtype=dtype).element_size()
        return self.workspace[off: off + nb].view(dtype).view(shape)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                L.load().tpdm_plan_destroy(self.handle)
        except Exception:
            pass


class Engine:
    """One per (transformer, time_predictor) pair and device."""

    def __init__(self, cfg: dict, device: torch.device, transformer_sd: Optional[Dict[str, torch.Tensor]] = None, transformer_cfg=None,
                 tpm_sd: Optional[Dict[str, torch.Tensor]] = None, min_sigma: float = 0.001, relative: bool = True,
                 prediction_type: str = "alpha_beta", epsilon: float = 1e-3, tpm_epsilon: float = 1.0, tpm_channels: int = 128):
        lib = L.load()
        if device.type != "cuda":
            raise RuntimeError("tpdm_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if prediction_type not in ("alpha_beta", "mode_concentration"):
            raise ValueError(f"unknown prediction_type {prediction_type!r}")
        self.cfg, self.device = dict(cfg), device
        self.D = cfg["num_attention_heads"] * cfg["attention_head_dim"]
        c = L.TpdmConfig(
            num_layers=cfg["num_layers"], num_heads=cfg["num_attention_heads"], head_dim=cfg["attention_head_dim"],
            joint_attention_dim=cfg["joint_attention_dim"], pooled_projection_dim=cfg["pooled_projection_dim"],
            in_channels=cfg["in_channels"], out_channels=cfg["out_channels"], patch_size=cfg["patch_size"],
            pos_embed_max_size=cfg["pos_embed_max_size"], qk_norm=1 if cfg.get("qk_norm") == "rms_norm" else 0,
            tpm_channels=tpm_channels, prediction_type=0 if prediction_type == "alpha_beta" else 1, relative=1 if relative else 0,
            min_sigma=min_sigma, epsilon=epsilon, tpm_epsilon=tpm_epsilon,
            dual_attention_mask=sum(1 << int(i) for i in set(cfg.get("dual_attention_layers", ()) or ())))
        with torch.cuda.device(device):
            handle = L.vp()
            L.check(lib.tpdm_create(C.byref(c), C.byref(handle)))
            self.ctx = handle
            self.weights = PackedWeights()
            self.has_mmdit = transformer_sd is not None
            self.has_tpm = tpm_sd is not None
            if self.has_mmdit:
                pack_transformer(self.weights, transformer_sd, transformer_cfg, device)
            if self.has_tpm:
                pack_time_predictor(self.weights, tpm_sd, device)
            L.check(lib.tpdm_set_weights(self.ctx, C.byref(self.weights.struct)))
        self.plans: Dict[Tuple, Plan] = {}
        self.min_sigma = min_sigma

    def __del__(self):
        try:
            self.plans.clear()
            if getattr(self, "ctx", None):
                L.load().tpdm_destroy(self.ctx)
        except Exception:
            pass

    def refresh_time_predictor(self, tpm_sd: Dict[str, torch.Tensor]) -> None:
        """Copy new TimePredictor values into the already packed device tensors (pointers and TMA descriptors unchanged)."""
        if not self.has_tpm:
            raise RuntimeError("engine was built without a TimePredictor")
        with torch.cuda.device(self.device):
            pack_time_predictor(self.weights, tpm_sd, self.device)

    def plan(self, batch: int, cfg_pairs: bool, latent: int, n_text: int, max_steps: int = 1) -> Plan:
        key = (batch, bool(cfg_pairs), latent, n_text, max_steps)
        p = self.plans.get(key)
        if p is None:
            if len(self.plans) >= 4:  # bound the HBM held by cached workspaces
                self.plans.pop(next(iter(self.plans)))
            with torch.cuda.device(self.device):
                p = Plan(self, batch, cfg_pairs, latent, n_text, max_steps)
            self.plans[key] = p
        return p

    # ---- CustomSD3Transformer2DModel.forward -------------------------------------------------------------------
    def mmdit_forward(self, hidden_states, encoder_hidden_states, pooled_projections, timestep):
        lib = L.load()
        Bt, Cc, h, w = hidden_states.shape
        if h != w:
            raise ValueError(f"only square latents are supported (got {h}x{w})")
        n_text = encoder_hidden_states.shape[1]
        plan = self.plan(Bt, False, h, n_text, 1)
        dev, f32 = self.device, torch.float32
        lat = hidden_states.to(device=dev, dtype=f32).contiguous()
        enc = encoder_hidden_states.to(device=dev, dtype=f32).contiguous()
        pooled = pooled_projections.to(device=dev, dtype=f32).contiguous()
        ts = timestep.to(device=dev, dtype=f32).reshape(-1)
        if ts.numel() == 1:
            ts = ts.expand(Bt)
        ts = ts.contiguous()
        if enc.shape[0] != Bt or pooled.shape[0] != Bt or ts.shape[0] != Bt:
            raise ValueError("batch sizes of hidden_states / encoder_hidden_states / pooled_projections / timestep differ")
        N = (h // 2) * (w // 2)
        out = torch.empty(Bt, self.cfg["out_channels"], h, w, device=dev, dtype=f32)
        temb = torch.empty(Bt, self.D, device=dev, dtype=f32)
        h1 = torch.empty(Bt, N, self.D, device=dev, dtype=f32)
        h2 = torch.empty(Bt, N, self.D, device=dev, dtype=f32)
        with torch.cuda.device(dev):
            L.check(lib.tpdm_mmdit_forward(plan.handle, L.ptr(lat), L.ptr(ts), L.ptr(enc), L.ptr(pooled), L.ptr(out), L.ptr(temb),
                                           L.ptr(h1), L.ptr(h2), L.stream_ptr()))
        return out, temb, h1, h2

    # ---- TimePredictor.forward ---------------------------------------------------------------------------------
    def tpm_forward(self, x, temb):
        lib = L.load()
        B, C2, g, g2 = x.shape
        if g != g2 or C2 != 2 * self.D:
            raise ValueError(f"TimePredictor input must be (B, {2 * self.D}, g, g); got {tuple(x.shape)}")
        plan = self.plan(B, False, 2 * g, 8, 1)
        dev, f32 = self.device, torch.float32
        xx = x.to(device=dev, dtype=f32).contiguous()
        tt = temb.to(device=dev, dtype=f32).contiguous()
        out = torch.empty(B, 2, device=dev, dtype=f32)
        with torch.cuda.device(dev):
            L.check(lib.tpdm_tpm_forward(plan.handle, L.ptr(xx), L.ptr(tt), L.ptr(out), L.stream_ptr()))
        return out

    # ---- the adaptive loop -------------------------------------------------------------------------------------
    def sample(self, latents, prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds,
               max_inference_steps: int, guidance_scale: float, predict: bool, ratios=None, seed: int = 0,
               record_tpm_inputs: bool = False, record_velocity: bool = False, use_graph: Optional[bool] = None):
        """Runs tpdm_sample_step until every sigma_next < min_sigma (checked one step late through a pinned flag, the
        speculative extra step is skipped on the device) or max_inference_steps.  Returns a dict of device tensors.
        ``use_graph`` (default: on, ``TPDM_SAMPLE_GRAPH=0`` turns it off): every step index is replayed from a CUDA graph
        captured the first time it runs on the plan (~190 launches become one ``cudaGraphLaunch``); the trajectory then runs on a
        side stream, because the legacy default stream cannot be captured, and the caller's stream waits for it."""
        lib = L.load()
        if use_graph is None:
            use_graph = os.environ.get("TPDM_SAMPLE_GRAPH", "1") != "0"
        B, Cc, h, w = latents.shape
        dev, f32 = self.device, torch.float32
        plan = self.plan(B, True, h, prompt_embeds.shape[1], max_inference_steps)
        st = plan.state
        cvt = lambda t: t.to(device=dev, dtype=f32).contiguous()
        lat, pe, ne, pp, npp = cvt(latents), cvt(prompt_embeds), cvt(negative_prompt_embeds), cvt(pooled_prompt_embeds), cvt(negative_pooled_prompt_embeds)
        rt = None
        if not predict and ratios is not None:
            rt = cvt(ratios)
            if rt.shape != (B, max_inference_steps):
                raise ValueError(f"ratios must have shape {(B, max_inference_steps)}")
        rec = vel = None
        g = h // 2
        if record_tpm_inputs:
            rec = torch.empty(B, max_inference_steps, g, g, 2 * self.D, device=dev, dtype=torch.bfloat16)
        if record_velocity:
            vel = torch.empty(B, max_inference_steps, Cc, h, w, device=dev, dtype=f32)
        done_host = getattr(self, "_done_host", None)
        if done_host is None or done_host.numel() < max_inference_steps:
            done_host = self._done_host = torch.zeros(max(64, max_inference_steps), dtype=torch.int32).pin_memory()
        events = getattr(self, "_step_events", None)
        if events is None:
            events = self._step_events = [torch.cuda.Event(), torch.cuda.Event()]
        executed = max_inference_steps
        step_fn = lib.tpdm_sample_step_graph if use_graph else lib.tpdm_sample_step
        side = None
        if use_graph:
            side = getattr(self, "_sample_stream", None)
            if side is None:
                side = self._sample_stream = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.device(dev), torch.cuda.stream(side):
            stream = L.stream_ptr()
            L.check(lib.tpdm_sample_begin(plan.handle, L.ptr(lat), L.ptr(ne), L.ptr(pe), L.ptr(npp), L.ptr(pp), float(guidance_scale),
                                          1 if predict else 0, L.ptr(rt) if rt is not None else None, int(seed) & (2**64 - 1), stream))
            for step in range(max_inference_steps):
                L.check(step_fn(plan.handle, step, stream))
                if rec is not None:
                    rec[:, step].copy_(st["tpm_input"])
                if vel is not None:
                    vel[:, step].copy_(st["velocity"])
                done_host[step: step + 1].copy_(st["all_done"][step: step + 1], non_blocking=True)
                events[step & 1].record()
                if step >= 1:       # the flag of the previous step: the device never waits for the host
                    events[(step - 1) & 1].synchronize()
                    if int(done_host[step - 1]) != 0:
                        executed = step  # steps [0, step) are real; step `step` was skipped on the device
                        break
            else:
                events[(max_inference_steps - 1) & 1].synchronize()
        if side is not None:
            torch.cuda.current_stream(dev).wait_stream(side)
            for t in (lat, pe, ne, pp, npp, rt, rec, vel):
                if t is not None:
                    t.record_stream(side)
        T = executed
        out = dict(
            steps=T,
            sigmas=st["sigma_hist"][:, 1: T + 1].clone(), alphas=st["alphas"][:, :T].clone(), betas=st["betas"][:, :T].clone(),
            logprobs_raw=st["logprobs"][:, :T].clone(), prob_masks=st["prob_masks"][:, :T].clone().bool(),
            tembs=st["tembs"][:T].permute(1, 0, 2).clone(), history_latents=st["history_latents"][:T].permute(1, 0, 2, 3, 4).clone(),
        )
        if rec is not None:
            out["tpm_inputs_nhwc"] = rec[:, :T]
        if vel is not None:
            out["velocities"] = vel[:, :T]
        return out

    def _probe_and_order(self, lat, pe, ne, pp, npp, guidance_scale, max_steps, out_lat, out_steps, out_sig):
        """Longest-expected-first scheduling for the device queue.  Greedy list scheduling loses the tail: the last GPUs to finish
        are whoever drew a long prompt late (86.6 % of the sum-of-steps / N bound on 8 GPUs, profiles/r01_config3_devq_n8.log).
        Trajectory lengths are not known in advance, but the FIRST TimePredictor call already says how fast sigma falls for a
        prompt.  So every GPU runs denoising step 0 for its static share of the prompts (one batched call; the step is not wasted:
        it IS step 0 of those trajectories), the post-step latents and sigmas are all-gathered (1 MB per prompt over NVLink -- the
        one exchange of this scheduler, a few hundred microseconds), every rank sorts the prompts by the expected number of
        remaining steps, ceil(log(min_sigma / sigma_1) / log(sigma_1)), and the queue hands them out longest first with every
        prompt entering at step 1.  Ticket claiming stays dynamic, so a wrong estimate costs balance, never correctness: each
        prompt still follows exactly its own trajectory.  Returns (latents after step 0 [P], order, n_queued, sigma_1 [P])."""
        import torch.distributed as dist

        lib = L.load()
        P = lat.shape[0]
        dev = self.device
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        if P % world != 0:
            raise ValueError(f"schedule='lpt' needs the number of prompts ({P}) to be a multiple of the number of ranks ({world})")
        mine = torch.arange(rank, P, world, device=dev)
        nb = mine.numel()
        plan = self.plan(nb, True, lat.shape[-1], pe.shape[1], 1)
        st = plan.state
        sel = lambda t: t[mine].contiguous()
        l0, pe0, ne0, pp0, npp0 = sel(lat), sel(pe), sel(ne), sel(pp), sel(npp)
        with torch.cuda.device(dev):
            stream = L.stream_ptr()
            L.check(lib.tpdm_sample_begin(plan.handle, L.ptr(l0), L.ptr(ne0), L.ptr(pe0), L.ptr(npp0), L.ptr(pp0), float(guidance_scale), 1, None,
                                          0, stream))
            L.check(lib.tpdm_sample_step(plan.handle, 0, stream))
        packed = torch.cat([st["latents"].reshape(nb, -1), st["sigma_hist"][:, 1:2]], dim=1).contiguous()     # (nb, C*h*w + 1)
        if world > 1:
            allp = torch.empty(world, nb, packed.shape[1], device=dev, dtype=packed.dtype)
            dist.all_gather_into_tensor(allp, packed)
            allp = allp.permute(1, 0, 2).reshape(P, -1)          # prompt p = rank + k * world  ->  row k * world + rank = p
        else:
            allp = packed
        lat1 = allp[:, :-1].reshape(lat.shape).contiguous()
        sig1 = allp[:, -1].contiguous()
        from .work_queue import expected_remaining_steps, longest_first_order

        done = sig1 < self.min_sigma                              # the whole trajectory was one step (:608)
        order = longest_first_order(expected_remaining_steps(sig1, self.min_sigma, max_steps)).contiguous()
        n_queued = int((~done).sum().item())
        out_sig[:, 0] = 1.0
        out_sig[:, 1] = sig1
        probed_here = torch.zeros(P, dtype=torch.bool, device=dev)
        probed_here[mine] = True
        fin = done & probed_here                                  # finished at the probe: reported by the rank that ran it
        out_steps[fin] = 1
        out_lat[fin] = lat1[fin]
        return lat1, order, n_queued, sig1

    def sample_queue(self, latents, prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds,
                     slots: int, max_inference_steps: int, guidance_scale: float, ticket: Optional[torch.Tensor] = None,
                     out_latents: Optional[torch.Tensor] = None, use_graph: bool = True, schedule: str = "fifo"):
        """Device-side prompt queue (``tpdm_queue_*``): all P prompts are resident, ``slots`` of them are in flight; a slot
        whose trajectory ends takes the next ticket on the device.  ``ticket`` (int32 device tensor of one element, zeroed by
        its owner) may be shared between GPUs, which then split one prompt list between them.  Returns device tensors:
        latents (P, C, h, w), steps (P,) (0 for prompts another GPU processed), sigmas (P, max_steps + 1)."""
        lib = L.load()
        P, Cc, h, w = latents.shape
        dev, f32 = self.device, torch.float32
        if schedule not in ("fifo", "lpt"):
            raise ValueError(f"unknown queue schedule {schedule!r} ('fifo': tickets in prompt order, 'lpt': longest expected trajectory first)")
        if not 1 <= slots <= P:
            raise ValueError(f"slots must be in [1, {P}]")
        cvt = lambda t: t.to(device=dev, dtype=f32).contiguous()
        lat, pe, ne, pp, npp = cvt(latents), cvt(prompt_embeds), cvt(negative_prompt_embeds), cvt(pooled_prompt_embeds), cvt(negative_pooled_prompt_embeds)
        if ticket is None:
            ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        out_lat = torch.zeros(P, Cc, h, w, device=dev, dtype=f32) if out_latents is None else out_latents
        out_steps = torch.zeros(P, dtype=torch.int32, device=dev)
        out_sig = torch.zeros(P, max_inference_steps + 1, device=dev, dtype=f32)
        order = init_sigma = None
        n_queued, init_step = P, 0
        if schedule == "lpt" and max_inference_steps > 1:
            lat, order, n_queued, init_sigma = self._probe_and_order(lat, pe, ne, pp, npp, guidance_scale, max_inference_steps, out_lat, out_steps, out_sig)
            init_step = 1
            if n_queued == 0:       # every trajectory ended at its first step
                return dict(latents=out_lat, steps=out_steps, sigmas=out_sig, device_steps=1)
        plan = self.plan(slots, True, h, prompt_embeds.shape[1], max_inference_steps)
        nbytes = lib.tpdm_queue_workspace_bytes(plan.handle, P)
        ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        base = (ws.data_ptr() + 1023) // 1024 * 1024
        active_host = torch.ones(64, dtype=torch.int32).pin_memory()
        # the step is replayed from a CUDA graph (captured on the first step), which needs a non-default stream
        qs = getattr(self, "_queue_stream", None)
        if qs is None:
            qs = self._queue_stream = torch.cuda.Stream(device=dev)
        qs.wait_stream(torch.cuda.current_stream(dev))
        step_fn = lib.tpdm_queue_step_graph if use_graph else lib.tpdm_queue_step
        with torch.cuda.device(dev), torch.cuda.stream(qs):
            stream = L.stream_ptr()
            L.check(lib.tpdm_queue_begin(plan.handle, P, L.ptr(lat), L.ptr(ne), L.ptr(pe), L.ptr(npp), L.ptr(pp), float(guidance_scale), base,
                                         nbytes, ticket.data_ptr(), L.ptr(out_lat), L.ptr(out_steps), L.ptr(out_sig),
                                         L.ptr(order) if order is not None else None, int(n_queued),
                                         L.ptr(init_sigma) if init_sigma is not None else None, int(init_step), stream))
            act_p, slot_p = L.vp(), L.vp()
            L.check(lib.tpdm_queue_status(plan.handle, C.byref(act_p), C.byref(slot_p)))
            off = act_p.value - ws.data_ptr()          # the counters live in the queue workspace
            active = ws[off: off + 4].view(torch.int32)
            ring = active_host.numel()
            events = [None] * ring
            limit = (P + slots - 1) // slots * max_inference_steps + 2 * max_inference_steps   # upper bound on device steps
            steps_run = 0
            drained = False
            for it in range(limit):
                L.check(step_fn(plan.handle, stream))
                steps_run += 1
                active_host[it % ring: it % ring + 1].copy_(active, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                events[it % ring] = ev
                if it >= 1:                      # look at the flag one step late: the device never waits for the host
                    events[(it - 1) % ring].synchronize()
                    if int(active_host[(it - 1) % ring]) == 0:
                        drained = True
                        break
            if not drained:
                torch.cuda.current_stream().synchronize()
                if int(active.item()) != 0:
                    raise RuntimeError("sample_queue: the queue did not drain within the step bound")
            torch.cuda.current_stream().synchronize()
        torch.cuda.current_stream(dev).wait_stream(qs)
        self._queue_keepalive = (ws, lat, pe, ne, pp, npp, ticket, order, init_sigma)
        return dict(latents=out_lat, steps=out_steps, sigmas=out_sig, device_steps=steps_run + (1 if init_step else 0))
