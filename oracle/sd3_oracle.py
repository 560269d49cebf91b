"""CPU ORACLE (test infrastructure, NOT the product path).

Plain-PyTorch fp32 restatement of the TPDM adaptive denoising path of jinkyu032/TPDM.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
this module; ``tpdm_b200`` never does.

PARITY PINNING.  The reference ships no tests, golden vectors or known-answer values for this path
 and its arithmetic lives in ``diffusers>=0.31.0`` (requirements.txt:5), which is neither
vendored under /root/reference nor installed here.  So:
  * everything that IS in-tree and pure torch (TimePredictor, CustomAdaGroupNormZeroSingle,
    reshape_hidden_states_to_2d, custom_step body, get_ref_beta, get_kl_beta) is pinned: ``oracle/ref_extract.py``
    AST-extracts those definitions from /root/reference and ``oracle/make_golden.py`` asserts the restatements
    below are bit-identical to them before it writes ``tests/golden``;
  * the diffusers classes (PatchEmbed, CombinedTimestepTextProjEmbeddings, JointTransformerBlock, Attention +
    JointAttnProcessor2_0, AdaLayerNormZero, AdaLayerNormContinuous, FeedForward, RMSNorm) are restated from the
    published diffusers 0.31 algorithm and anchored on the reference's call sites -- **parity unpinned** for
    those rows (DESIGN.md says the same).  The sincos table is cross-checked against the MAE implementation
    that transformers ships (tests/test_oracle.py).

State-dict names follow the diffusers layout so one state dict feeds oracle and CUDA path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------------------------
@dataclass
class SD3Config:
    """Mirrors CustomSD3Transformer2DModel.__init__ kwargs (transformer_sd3.py:91-108)."""

    sample_size: int = 128
    patch_size: int = 2
    in_channels: int = 16
    num_layers: int = 18
    attention_head_dim: int = 64
    num_attention_heads: int = 18
    joint_attention_dim: int = 4096
    caption_projection_dim: int = 1152
    pooled_projection_dim: int = 2048
    out_channels: int = 16
    pos_embed_max_size: int = 96
    qk_norm: Optional[str] = None
    dual_attention_layers: tuple = ()   # () for SD3.0; (0..12) for SD3.5-medium (transformer_sd3.py:104-106)

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


def tiny_config(qk_norm: Optional[str] = None) -> SD3Config:
    """BASELINE.json configs[0]: 2 joint blocks, hidden 384, 4 heads, 256^2 (32x32 latent)."""
    return SD3Config(sample_size=32, num_layers=2, attention_head_dim=96, num_attention_heads=4,
                     caption_projection_dim=384, pos_embed_max_size=96, qk_norm=qk_norm)


def sd3_medium_config(sample_size: int = 128, qk_norm: Optional[str] = None) -> SD3Config:
    """BASELINE.json configs[1..4]: SD3-medium (24 blocks, hidden 1536, 24 heads), pos_embed_max_size 192."""
    return SD3Config(sample_size=sample_size, num_layers=24, attention_head_dim=64, num_attention_heads=24,
                     caption_projection_dim=1536, pos_embed_max_size=192, qk_norm=qk_norm)


# --------------------------------------------------------------------------------------------------------------
# diffusers.models.embeddings restatements  (call sites: transformer_sd3.py:114-125, 334, 336)
# --------------------------------------------------------------------------------------------------------------
def get_1d_sincos(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    """[sin(p*w_j), cos(p*w_j)], w_j = 10000^(-j/(embed_dim/2)); float64 (diffusers get_1d_sincos_pos_embed_from_grid)."""
    omega = np.arange(embed_dim // 2, dtype=np.float64)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def get_2d_sincos_pos_embed(embed_dim: int, grid_size: int, base_size: int) -> np.ndarray:
    """(grid_size^2, embed_dim) table.  Coordinates are i / (grid_size/base_size); channels [0, D/2) encode the
    COLUMN coordinate and [D/2, D) the row coordinate (meshgrid 'w goes first', MAE layout)."""
    gh = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size)
    gw = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])
    emb_a = get_1d_sincos(embed_dim // 2, grid[0])
    emb_b = get_1d_sincos(embed_dim // 2, grid[1])
    return np.concatenate([emb_a, emb_b], axis=1)


class PatchEmbed(nn.Module):
    """Conv2d(k=s=patch) + flatten + centre-cropped sincos table (diffusers PatchEmbed with pos_embed_max_size)."""

    def __init__(self, height, width, patch_size, in_channels, embed_dim, pos_embed_max_size):
        super().__init__()
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=(patch_size, patch_size), stride=patch_size, bias=True)
        self.patch_size = patch_size
        self.base_size = height // patch_size
        self.pos_embed_max_size = pos_embed_max_size
        table = get_2d_sincos_pos_embed(embed_dim, pos_embed_max_size, base_size=self.base_size)
        self.register_buffer("pos_embed", torch.from_numpy(table).float().unsqueeze(0), persistent=True)

    def cropped_pos_embed(self, height: int, width: int) -> torch.Tensor:
        height, width = height // self.patch_size, width // self.patch_size
        if height > self.pos_embed_max_size or width > self.pos_embed_max_size:
            raise ValueError(f"Height/width ({height},{width}) exceed pos_embed_max_size {self.pos_embed_max_size}.")
        top = (self.pos_embed_max_size - height) // 2
        left = (self.pos_embed_max_size - width) // 2
        sp = self.pos_embed.reshape(1, self.pos_embed_max_size, self.pos_embed_max_size, -1)
        sp = sp[:, top: top + height, left: left + width, :]
        return sp.reshape(1, -1, sp.shape[-1])

    def forward(self, latent: torch.Tensor) -> torch.Tensor:
        height, width = latent.shape[-2:]
        latent = self.proj(latent).flatten(2).transpose(1, 2)  # BCHW -> BNC, token n = y*g + x
        return (latent + self.cropped_pos_embed(height, width)).to(latent.dtype)


def get_timestep_embedding(timesteps: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """[cos, sin](t * 10000^(-i/half)), i < half  (flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device) / half
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class _TwoLayer(nn.Module):
    """linear_1 -> SiLU -> linear_2 (TimestepEmbedding and PixArtAlphaTextProjection(act_fn='silu'))."""

    def __init__(self, d_in, d_hidden):
        super().__init__()
        self.linear_1 = nn.Linear(d_in, d_hidden)
        self.linear_2 = nn.Linear(d_hidden, d_hidden)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class CombinedTimestepTextProjEmbeddings(nn.Module):
    def __init__(self, embedding_dim, pooled_projection_dim):
        super().__init__()
        self.timestep_embedder = _TwoLayer(256, embedding_dim)
        self.text_embedder = _TwoLayer(pooled_projection_dim, embedding_dim)

    def forward(self, timestep, pooled_projection):
        t = get_timestep_embedding(timestep, 256).to(pooled_projection.dtype)
        return self.timestep_embedder(t) + self.text_embedder(pooled_projection)


# --------------------------------------------------------------------------------------------------------------
# diffusers normalisation / attention / feed-forward restatements
# --------------------------------------------------------------------------------------------------------------
class AdaLayerNormZero(nn.Module):
    """chunk order: shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp."""

    def __init__(self, dim):
        super().__init__()
        self.linear = nn.Linear(dim, 6 * dim)
        self.dim = dim

    def forward(self, x, emb):
        emb = self.linear(F.silu(emb))
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = emb.chunk(6, dim=1)
        x = F.layer_norm(x, (self.dim,), eps=1e-6) * (1 + scale_msa[:, None]) + shift_msa[:, None]
        return x, gate_msa, shift_mlp, scale_mlp, gate_mlp


class AdaLayerNormContinuous(nn.Module):
    """chunk order: SCALE first, then SHIFT."""

    def __init__(self, dim, cond_dim):
        super().__init__()
        self.linear = nn.Linear(cond_dim, 2 * dim)
        self.dim = dim

    def forward(self, x, cond):
        emb = self.linear(F.silu(cond).to(x.dtype))
        scale, shift = emb.chunk(2, dim=1)
        return F.layer_norm(x, (self.dim,), eps=1e-6) * (1 + scale)[:, None, :] + shift[:, None, :]


class RMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        var = x.float().pow(2).mean(-1, keepdim=True)
        return (x * torch.rsqrt(var + self.eps)) * self.weight


class _GELUProj(nn.Module):
    def __init__(self, d_in, d_out):
        super().__init__()
        self.proj = nn.Linear(d_in, d_out)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class FeedForward(nn.Module):
    """net.0 = GELU(tanh) projection D->4D, net.1 = Dropout(0), net.2 = Linear 4D->D."""

    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, 4 * dim), nn.Identity(), nn.Linear(4 * dim, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class SD35AdaLayerNormZeroX(nn.Module):
    """diffusers SD35AdaLayerNormZeroX (norm1 of a dual-attention block): 9 chunks
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp, shift_msa2, scale_msa2, gate_msa2; both modulated
    copies come from the same LayerNorm(x)."""

    def __init__(self, dim):
        super().__init__()
        self.linear = nn.Linear(dim, 9 * dim)
        self.dim = dim

    def forward(self, x, emb):
        emb = self.linear(F.silu(emb))
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp, shift_msa2, scale_msa2, gate_msa2 = emb.chunk(9, dim=1)
        n = F.layer_norm(x, (self.dim,), eps=1e-6)
        x1 = n * (1 + scale_msa[:, None]) + shift_msa[:, None]
        x2 = n * (1 + scale_msa2[:, None]) + shift_msa2[:, None]
        return x1, gate_msa, shift_mlp, scale_mlp, gate_mlp, x2, gate_msa2


class SelfAttention(nn.Module):
    """attn2 of a dual-attention block: diffusers Attention(cross_attention_dim=None, bias=True, qk_norm=...) with
    JointAttnProcessor2_0 and no encoder states -> plain self-attention over the image tokens."""

    def __init__(self, dim, heads, dim_head, qk_norm):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])
        if qk_norm == "rms_norm":
            self.norm_q, self.norm_k = RMSNorm(dim_head), RMSNorm(dim_head)
        self.qk_norm = qk_norm

    def forward(self, x):
        b, s, _ = x.shape
        split = lambda t: t.view(b, s, self.heads, self.dim_head).transpose(1, 2)
        q, k, v = split(self.to_q(x)), split(self.to_k(x)), split(self.to_v(x))
        if self.qk_norm is not None:
            q, k = self.norm_q(q), self.norm_k(k)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        return self.to_out[0](o.transpose(1, 2).reshape(b, s, self.heads * self.dim_head))


class JointAttention(nn.Module):
    """diffusers Attention(added_kv_proj_dim=dim, bias=True) driven by JointAttnProcessor2_0:
    image tokens first, text tokens second; plain softmax(QK^T/sqrt(d))V, no mask."""

    def __init__(self, dim, heads, dim_head, context_pre_only, qk_norm):
        super().__init__()
        self.heads, self.dim_head, self.context_pre_only = heads, dim_head, context_pre_only
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.add_k_proj, self.add_v_proj, self.add_q_proj = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])
        if not context_pre_only:
            self.to_add_out = nn.Linear(dim, dim)
        if qk_norm == "rms_norm":
            self.norm_q, self.norm_k = RMSNorm(dim_head), RMSNorm(dim_head)
            self.norm_added_q, self.norm_added_k = RMSNorm(dim_head), RMSNorm(dim_head)
        elif qk_norm is not None:
            raise ValueError(f"unknown qk_norm: {qk_norm}")
        self.qk_norm = qk_norm

    def _split(self, x):
        b, s, _ = x.shape
        return x.view(b, s, self.heads, self.dim_head).transpose(1, 2)

    def forward(self, hidden_states, encoder_hidden_states):
        n_img = hidden_states.shape[1]
        q, k, v = self._split(self.to_q(hidden_states)), self._split(self.to_k(hidden_states)), self._split(self.to_v(hidden_states))
        cq = self._split(self.add_q_proj(encoder_hidden_states))
        ck = self._split(self.add_k_proj(encoder_hidden_states))
        cv = self._split(self.add_v_proj(encoder_hidden_states))
        if self.qk_norm is not None:
            q, k, cq, ck = self.norm_q(q), self.norm_k(k), self.norm_added_q(cq), self.norm_added_k(ck)
        q, k, v = torch.cat([q, cq], dim=2), torch.cat([k, ck], dim=2), torch.cat([v, cv], dim=2)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(o.shape[0], -1, self.heads * self.dim_head)
        o_img, o_ctx = o[:, :n_img], o[:, n_img:]
        o_img = self.to_out[0](o_img)
        o_ctx = self.to_add_out(o_ctx) if not self.context_pre_only else None
        return o_img, o_ctx


class JointTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, context_pre_only=False, qk_norm=None, use_dual_attention=False):
        super().__init__()
        self.context_pre_only = context_pre_only
        self.use_dual_attention = use_dual_attention
        self.dim = dim
        self.norm1 = SD35AdaLayerNormZeroX(dim) if use_dual_attention else AdaLayerNormZero(dim)
        self.norm1_context = AdaLayerNormContinuous(dim, dim) if context_pre_only else AdaLayerNormZero(dim)
        self.attn = JointAttention(dim, heads, dim_head, context_pre_only, qk_norm)
        if use_dual_attention:
            self.attn2 = SelfAttention(dim, heads, dim_head, qk_norm)
        self.ff = FeedForward(dim)
        if not context_pre_only:
            self.ff_context = FeedForward(dim)

    def forward(self, hidden_states, encoder_hidden_states, temb):
        if self.use_dual_attention:
            n, gate_msa, shift_mlp, scale_mlp, gate_mlp, n2, gate_msa2 = self.norm1(hidden_states, temb)
        else:
            n, gate_msa, shift_mlp, scale_mlp, gate_mlp = self.norm1(hidden_states, temb)
        if self.context_pre_only:
            nc = self.norm1_context(encoder_hidden_states, temb)
        else:
            nc, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = self.norm1_context(encoder_hidden_states, temb)
        attn_out, ctx_attn_out = self.attn(n, nc)
        hidden_states = hidden_states + gate_msa.unsqueeze(1) * attn_out
        if self.use_dual_attention:
            hidden_states = hidden_states + gate_msa2.unsqueeze(1) * self.attn2(n2)
        m = F.layer_norm(hidden_states, (self.dim,), eps=1e-6) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
        hidden_states = hidden_states + gate_mlp.unsqueeze(1) * self.ff(m)
        if self.context_pre_only:
            encoder_hidden_states = None
        else:
            encoder_hidden_states = encoder_hidden_states + c_gate_msa.unsqueeze(1) * ctx_attn_out
            mc = F.layer_norm(encoder_hidden_states, (self.dim,), eps=1e-6) * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
            encoder_hidden_states = encoder_hidden_states + c_gate_mlp.unsqueeze(1) * self.ff_context(mc)
        return encoder_hidden_states, hidden_states


# --------------------------------------------------------------------------------------------------------------
# CustomSD3Transformer2DModel.forward  (transformer_sd3.py:299-409)
# --------------------------------------------------------------------------------------------------------------
class OracleSD3Transformer(nn.Module):
    def __init__(self, cfg: SD3Config):
        super().__init__()
        self.cfg = cfg
        d = cfg.inner_dim
        if d != cfg.caption_projection_dim:
            raise ValueError("inner_dim must equal caption_projection_dim")
        self.pos_embed = PatchEmbed(cfg.sample_size, cfg.sample_size, cfg.patch_size, cfg.in_channels, d, cfg.pos_embed_max_size)
        self.time_text_embed = CombinedTimestepTextProjEmbeddings(d, cfg.pooled_projection_dim)
        self.context_embedder = nn.Linear(cfg.joint_attention_dim, cfg.caption_projection_dim)
        self.transformer_blocks = nn.ModuleList([
            JointTransformerBlock(d, cfg.num_attention_heads, cfg.attention_head_dim,
                                  context_pre_only=(i == cfg.num_layers - 1), qk_norm=cfg.qk_norm,
                                  use_dual_attention=i in tuple(cfg.dual_attention_layers))
            for i in range(cfg.num_layers)])
        self.norm_out = AdaLayerNormContinuous(d, d)
        self.proj_out = nn.Linear(d, cfg.patch_size * cfg.patch_size * cfg.out_channels)

    def forward(self, hidden_states, encoder_hidden_states, pooled_projections, timestep, return_blocks=False):
        height, width = hidden_states.shape[-2:]
        hidden_states = self.pos_embed(hidden_states)                       # :334
        hidden_states_1 = hidden_states.clone()                             # :335
        temb = self.time_text_embed(timestep, pooled_projections)           # :336
        encoder_hidden_states = self.context_embedder(encoder_hidden_states)  # :337
        per_block = []
        for block in self.transformer_blocks:                               # :339-365
            encoder_hidden_states, hidden_states = block(hidden_states, encoder_hidden_states, temb)
            if return_blocks:
                per_block.append(hidden_states)
        hidden_states = self.norm_out(hidden_states, temb)                  # :372
        hidden_states_2 = hidden_states.clone()                             # :373
        hidden_states = self.proj_out(hidden_states)                        # :374
        p, c = self.cfg.patch_size, self.cfg.out_channels                   # :377-399 unpatchify
        h, w = height // p, width // p
        hidden_states = hidden_states.reshape(hidden_states.shape[0], h, w, p, p, c)
        hidden_states = torch.einsum("nhwpqc->nchpwq", hidden_states)
        output = hidden_states.reshape(hidden_states.shape[0], c, h * p, w * p)
        if return_blocks:
            return output, temb, hidden_states_1, hidden_states_2, per_block
        return output, temb, hidden_states_1, hidden_states_2


# --------------------------------------------------------------------------------------------------------------
# in-tree reference pieces, restated (pinned bit-exact against AST-extracted reference by make_golden.py)
# --------------------------------------------------------------------------------------------------------------
def reshape_hidden_states_to_2d(hidden_states: torch.Tensor, height: int, width: int, patch_size: int = 2) -> torch.Tensor:
    """modeling_sd3_pnt.py:33-54.  NOTE the scramble: token n lands at pixel
    (y, x) = (2*(n // (2g)) + (n % 4) // 2,  2*((n % (2g)) // 4) + n % 2)  for a g x g grid."""
    b, _, c = hidden_states.shape
    x = hidden_states.reshape(b, height // patch_size, width // patch_size, patch_size, patch_size, c)
    x = torch.einsum("nhwpqc->nchpwq", x)
    return x.reshape(b, c, height, width)


class OracleAdaGroupNormZeroSingle(nn.Module):
    """modeling_sd3_pnt.py:56-83: [shift, scale] = Linear(SiLU(emb)); GroupNorm(1, C, eps 1e-6)*(1+scale)+shift."""

    def __init__(self, input_dim, embedding_dim):
        super().__init__()
        self.linear = nn.Linear(input_dim, 2 * embedding_dim)
        self.norm = nn.GroupNorm(1, embedding_dim, eps=1e-6)

    def forward(self, x, emb):
        emb = self.linear(F.silu(emb))
        shift, scale = emb.chunk(2, dim=1)
        return self.norm(x) * (1 + scale[:, :, None, None]) + shift[:, :, None, None]


class OracleTimePredictor(nn.Module):
    """modeling_sd3_pnt.py:85-126."""

    def __init__(self, conv_out_channels=128, in_channels=3072, projection_dim=2, init_alpha=1.5, init_beta=0.5):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, conv_out_channels, kernel_size=(3, 3), padding=1)
        self.conv2 = nn.Conv2d(conv_out_channels, conv_out_channels, kernel_size=(3, 3), padding=1, stride=2)
        self.fc1 = nn.Linear(conv_out_channels, 128)
        self.fc2 = nn.Linear(128, projection_dim)
        self.norm1 = OracleAdaGroupNormZeroSingle(in_channels // 2, conv_out_channels)
        self.epsilon = 1.0
        self.init_alpha, self.init_beta = init_alpha, init_beta
        self._init_weights()

    def forward(self, x, temb):
        x = self.conv1(x)
        x = self.norm1(x, temb)
        x = F.silu(x)
        x = self.conv2(x)
        x = F.adaptive_avg_pool2d(x, (16, 16))
        x = F.adaptive_max_pool2d(x, (1, 1)).view(x.size(0), -1)
        x = F.silu(self.fc1(x))
        x = self.fc2(x)
        return torch.exp(x) + self.epsilon

    def _init_weights(self):
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.normal_(m.weight, std=0.02)
                if m.bias is not None and isinstance(m, nn.Conv2d):
                    nn.init.constant_(m.bias, 0)
        nn.init.constant_(self.fc1.bias, 0)
        nn.init.constant_(self.fc2.bias[0], self.init_alpha)
        nn.init.constant_(self.fc2.bias[1], self.init_beta)


def custom_step(model_output, sigma_next, sigma, sample):
    """model_utilis.py:61-69: fp32 Euler update, cast back to the model-output dtype."""
    sample = sample.to(torch.float32)
    delta = (sigma_next - sigma).view(-1, 1, 1, 1)
    return (sample + delta * model_output).to(model_output.dtype)


EX = math.exp(1)


def get_ref_beta(sigmas_1: torch.Tensor, num_steps: int = 28):
    """reference_distributions.py:9-19."""
    t_1 = sigmas_1 / (EX + (1 - EX) * sigmas_1)
    t_2 = torch.clamp(t_1 - 1.0 / num_steps, 1e-3)
    sigmas_2 = EX / (EX + 1 / t_2 - 1)
    mode = sigmas_2 / sigmas_1
    return mode * (20 - 2) + 1, (1 - mode) * (20 - 2) + 1


def get_kl_beta(beta1, alpha1, beta2, alpha2):
    """train_utilis.py:6-20 (argument names as in the reference, which swaps them)."""
    B1 = torch.special.gammaln(alpha1) + torch.special.gammaln(beta1) - torch.special.gammaln(alpha1 + beta1)
    B2 = torch.special.gammaln(alpha2) + torch.special.gammaln(beta2) - torch.special.gammaln(alpha2 + beta2)
    return ((B2 - B1) + (alpha1 - alpha2) * torch.special.digamma(alpha1) + (beta1 - beta2) * torch.special.digamma(beta1)
            - (alpha1 - alpha2 + beta1 - beta2) * torch.special.digamma(alpha1 + beta1))


def beta_log_prob(alpha, beta, x):
    """torch.distributions.Beta(alpha, beta).log_prob(x) written out (Dirichlet log-density)."""
    return ((alpha - 1) * torch.log(x) + (beta - 1) * torch.log1p(-x)
            + torch.lgamma(alpha + beta) - torch.lgamma(alpha) - torch.lgamma(beta))


# --------------------------------------------------------------------------------------------------------------
# the adaptive loop: SD3PredictNextTimeStepModel.forward (modeling_sd3_pnt.py:447-668), VAE / text towers removed
# --------------------------------------------------------------------------------------------------------------
class OraclePipeline(nn.Module):
    def __init__(self, cfg: SD3Config, min_sigma=0.001, init_alpha=1.5, init_beta=0.5, relative=True,
                 prediction_type="alpha_beta"):
        super().__init__()
        self.cfg = cfg
        self.transformer = OracleSD3Transformer(cfg)
        self.time_predictor = OracleTimePredictor(128, cfg.caption_projection_dim * 2, 2, init_alpha, init_beta)
        self.min_sigma, self.relative, self.epsilon, self.prediction_type = min_sigma, relative, 1e-3, prediction_type
        self.requires_grad_(False).eval()

    @torch.no_grad()
    def forward(self, prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds,
                latents, max_inference_steps=28, guidance_scale=7.0, predict=True, ratios=None,
                record_velocity=False) -> Dict[str, torch.Tensor]:
        """``ratios`` (B, T) injects the Beta draws when predict=False so that two implementations follow the
        same trajectory (the reference calls beta_dist.sample(), :569).  Grid side is derived from the latent
        (the reference hard-codes 64, :35-36,550-551)."""
        batch_size = prompt_embeds.shape[0]
        g = latents.shape[-1] // self.cfg.patch_size
        init_noise_latents = latents.clone()
        prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)                     # :505
        pooled = torch.cat([negative_pooled_prompt_embeds, pooled_prompt_embeds], dim=0)              # :506
        sigma = torch.ones(batch_size, dtype=latents.dtype, device=latents.device)                    # :508
        sigmas, logprobs, prob_masks, alphas, betas = ([[] for _ in range(batch_size)] for _ in range(5))
        hcs, tembs, vels, hist = [], [], [], []
        for step in range(max_inference_steps):                                                       # :522
            latent_model_input = torch.cat([latents] * 2)                                             # :524
            timestep = sigma.repeat(2) * 1000                                                         # :526
            noise_pred, temb, h1, h2 = self.transformer(latent_model_input, prompt_embeds, pooled, timestep)
            nu, nt = noise_pred.chunk(2); noise_pred = nu + guidance_scale * (nt - nu)                # :537-538
            tu, tt = temb.chunk(2); temb = tu + guidance_scale * (tt - tu)                            # :539-540
            h1u, h1t = h1.chunk(2); h1 = h1u + guidance_scale * (h1t - h1u)                           # :541-544
            h2u, h2t = h2.chunk(2); h2 = h2u + guidance_scale * (h2t - h2u)                           # :545-548
            hc = torch.cat([reshape_hidden_states_to_2d(h1, g, g), reshape_hidden_states_to_2d(h2, g, g)], dim=1)
            hcs.append(hc); tembs.append(temb)
            if record_velocity:
                vels.append(noise_pred.clone())
            time_preds = self.time_predictor(hc, temb)                                                # :556
            sigma_next = torch.zeros_like(sigma)
            for i, (p1, p2) in enumerate(time_preds):                                                 # :558-590
                if self.prediction_type == "alpha_beta":
                    alpha, beta = p1, p2
                else:
                    alpha, beta = p1 * (p2 - 2) + 1, (1 - p1) * (p2 - 2) + 1
                if predict:
                    ratio = (alpha - 1) / (alpha + beta - 2)          # Beta.mode
                else:
                    ratio = ratios[i, step].to(alpha.dtype)
                ratio = (ratio.clamp(self.epsilon, 1 - self.epsilon) if self.relative
                         else ratio.clamp(self.epsilon, sigma[i]).clamp(0, 1 - self.epsilon))
                sigma_next[i] = sigma[i] * ratio if self.relative else sigma[i] - ratio
                sigmas[i].append(sigma_next[i])   # a VIEW, as in the reference: the :585 write below shows through
                logprobs[i].append(beta_log_prob(alpha, beta, ratio))
                if sigma[i] < self.min_sigma:
                    prob_masks[i].append(torch.tensor(1))
                    if predict:
                        sigma_next[i] = 0.0
                else:
                    prob_masks[i].append(torch.tensor(0))
                alphas[i].append(alpha); betas[i].append(beta)
            latents = custom_step(noise_pred, sigma_next, sigma, latents)                             # :592-598
            hist.append(latents.clone())
            if (sigma_next < self.min_sigma).all():                                                   # :608
                break
            sigma = sigma_next
        stack = lambda xs: torch.stack([torch.stack(x) for x in xs])
        sigmas, logprobs, alphas, betas = stack(sigmas), stack(logprobs), stack(alphas), stack(betas)
        prob_masks = stack(prob_masks).bool().to(logprobs.device)   # the reference builds the mask entries on the host (:583,:588)
        logprobs = torch.masked_fill(logprobs, prob_masks, 1.0)                                       # :621
        hist = torch.stack(hist, dim=1)                                                               # (B, T, C, h, w)
        last_valid = torch.stack([torch.where(~prob_masks[i])[0][-1] for i in range(batch_size)])     # :647
        out = dict(init_noise_latents=init_noise_latents, sigmas=sigmas, logprobs=logprobs, prob_masks=prob_masks,
                   alphas=alphas, betas=betas, tembs=torch.stack(tembs, dim=1),
                   hidden_states_combineds=torch.stack(hcs, dim=1), history_latents=hist,
                   last_valid_indices=last_valid,
                   final_latents=torch.stack([hist[i, last_valid[i]] for i in range(batch_size)]))
        if record_velocity:
            out["velocities"] = torch.stack(vels, dim=1)
        return out

    def only_predict_logprobs(self, fix_sigmas, fix_hidden_states_combineds, fix_tembs):
        """modeling_sd3_pnt.py:670-726 (differentiable w.r.t. time_predictor parameters)."""
        if fix_sigmas is None:
            raise ValueError("fix_sigmas must be provided")
        if fix_hidden_states_combineds is None:
            raise ValueError("fix_hidden_states_combineds must be provided")
        bsz, steps = fix_sigmas.shape[:2]
        sigma = torch.ones(bsz, dtype=fix_sigmas.dtype, device=fix_sigmas.device)
        logprobs, masks = [[] for _ in range(bsz)], [[] for _ in range(bsz)]
        for step in range(steps):
            tp = self.time_predictor(fix_hidden_states_combineds[:, step], fix_tembs[:, step])
            sigma_next = torch.zeros_like(sigma)
            for i, (alpha, beta) in enumerate(tp):
                sigma_next[i] = fix_sigmas[i][step]
                if sigma[i] < self.min_sigma:
                    logprobs[i].append(torch.zeros((), dtype=tp.dtype, device=tp.device)); masks[i].append(torch.tensor(1))
                    continue
                ratio = sigma_next[i] / sigma[i] if self.relative else sigma[i] - sigma_next[i]
                ratio = torch.clamp(ratio, min=self.epsilon, max=1 - self.epsilon)
                logprobs[i].append(beta_log_prob(alpha, beta, ratio)); masks[i].append(torch.tensor(0))
            sigma = sigma_next
        lp = torch.stack([torch.stack(x) for x in logprobs])
        mk = torch.stack([torch.stack(x) for x in masks]).bool().to(lp.device)
        return {"logprobs": torch.masked_fill(lp, mk, 1.0)}


# --------------------------------------------------------------------------------------------------------------
# RLOO pieces (rows R2, R3 of SURVEY.md §8a)
# --------------------------------------------------------------------------------------------------------------
def rloo_advantage(rlhf_reward: torch.Tensor, rloo_k: int) -> torch.Tensor:
    """rloo_trainer.py:458-461."""
    r = rlhf_reward.reshape(rloo_k, -1)
    baseline = (r.sum(0) - r) / (rloo_k - 1)
    return (r - baseline).flatten()


def ppo_clip_loss(new_logprobs, old_logprobs, advantage, cliprange=0.2):
    """rloo_trainer.py:485-495."""
    ratio = torch.exp(new_logprobs.sum(1) - old_logprobs.sum(1))
    l1 = -advantage * ratio
    l2 = -advantage * torch.clamp(ratio, 1.0 - cliprange, 1.0 + cliprange)
    return torch.max(l1, l2).mean()


def discounted_reward(last_reward: float, last_idx: int, gamma: float) -> float:
    """modeling_sd3_pnt.py:838-841."""
    return sum(last_reward * gamma ** (last_idx - i) for i in range(last_idx + 1)) / (last_idx + 1)


# --------------------------------------------------------------------------------------------------------------
# deterministic synthetic weights / inputs
# --------------------------------------------------------------------------------------------------------------
def build_pipeline(cfg: SD3Config, seed: int = 1234, **kw) -> OraclePipeline:
    """PyTorch default nn.Linear/Conv2d init under manual_seed(seed) for the MMDiT; the reference's own
    _init_weights for the TPM (modeling_sd3_pnt.py:117-126)."""
    torch.manual_seed(seed)
    return OraclePipeline(cfg, **kw)


def synthetic_inputs(cfg: SD3Config, batch: int, seed: int = 0, n_text: int = 333, latent_size: Optional[int] = None):
    g = torch.Generator().manual_seed(seed)
    ls = latent_size or cfg.sample_size
    return dict(
        prompt_embeds=torch.randn(batch, n_text, cfg.joint_attention_dim, generator=g),
        negative_prompt_embeds=torch.randn(batch, n_text, cfg.joint_attention_dim, generator=g),
        pooled_prompt_embeds=torch.randn(batch, cfg.pooled_projection_dim, generator=g),
        negative_pooled_prompt_embeds=torch.randn(batch, cfg.pooled_projection_dim, generator=g),
        latents=torch.randn(batch, cfg.in_channels, ls, ls, generator=g),
    )


def mmdit_flops(cfg: SD3Config, bt: int, n_img: int, n_txt: int) -> float:
    """Algorithmic FLOPs of one MMDiT forward."""
    d, s, L = cfg.inner_dim, n_img + n_txt, cfg.num_layers
    per_block = 2 * bt * s * 12 * d * d + 4 * bt * s * s * d
    last = 2 * bt * (n_img * 12 * d * d + n_txt * 3 * d * d) + 4 * bt * s * s * d
    embed = 2 * bt * (n_txt * cfg.joint_attention_dim * d + n_img * (64 * d + d * 64))
    return (L - 1) * per_block + last + embed
