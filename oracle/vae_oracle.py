"""TEST INFRASTRUCTURE ONLY -- fp32 PyTorch restatement of the VAE *decode* step that follows the adaptive denoising loop
(SURVEY.md section 8(f) rank 1).  Nothing under tpdm_b200/ may import this file.

Reference call sites (relative to /root/reference):
  * src/models/stable_diffusion_3/modeling_sd3_pnt.py:144-146  ``AutoencoderKL.from_pretrained(..., subfolder="vae")``
  * :631 / :653-655  ``latents = latents / vae.config.scaling_factor + vae.config.shift_factor``;
                     ``image = vae.decode(latents, return_dict=False)[0]``; ``image_processor.postprocess(image, "pil")``
  * :181-184         ``vae_scale_factor = 2 ** (len(block_out_channels) - 1)``

PARITY UNPINNED: ``AutoencoderKL`` lives in diffusers (>= 0.31, requirements.txt:5), which is neither vendored in the
reference nor installed here, and the reference ships no golden image / known-answer test for the decode.  The classes
below restate the published diffusers 0.31 algorithm for the SD3 VAE configuration (``latent_channels=16``,
``block_out_channels=(128, 256, 512, 512)``, ``layers_per_block=2``, ``norm_num_groups=32``, ``mid_block_add_attention``,
``use_quant_conv=False``, ``use_post_quant_conv=False``, ``scaling_factor=1.5305``, ``shift_factor=0.0609``):

  Decoder.forward      conv_in -> UNetMidBlock2D(resnet, attention, resnet) -> 4 x UpDecoderBlock2D -> GroupNorm -> SiLU -> conv_out
  UpDecoderBlock2D     (layers_per_block + 1) ResnetBlock2D, then Upsample2D (nearest x2 + conv3x3) except in the last block;
                       channels run over reversed(block_out_channels), each block's first resnet takes the previous width
  ResnetBlock2D        GroupNorm(32, eps 1e-6) -> SiLU -> conv3x3 -> GroupNorm -> SiLU -> conv3x3, + input (1x1 conv_shortcut
                       when the width changes), output_scale_factor 1, no time embedding
  Attention (mid)      one head of width C: GroupNorm(32, eps 1e-6) on (B, C, HW) -> to_q/to_k/to_v -> softmax(q k^T / sqrt(C)) v
                       -> to_out[0] -> + input
  VaeImageProcessor.postprocess   (image / 2 + 0.5).clamp(0, 1) -> HWC -> uint8 by (x * 255).round()

Module / parameter names follow the diffusers state-dict layout (``decoder.up_blocks.1.resnets.0.conv_shortcut.weight`` ...)
so that one state dict feeds both this oracle and the CUDA path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class VAEConfig:
    latent_channels: int = 16
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 1.5305
    shift_factor: float = 0.0609


def sd3_vae_config() -> VAEConfig:
    return VAEConfig()


def tiny_vae_config() -> VAEConfig:
    """two resolution levels, 64/128 channels, 16 groups (>= 4 channels per group): every code path (shortcut conv,
    upsampler, attention) in a few MFLOP"""
    return VAEConfig(block_out_channels=(64, 128), layers_per_block=1, norm_num_groups=16)


class OracleResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class OracleVAEAttention(nn.Module):
    def __init__(self, channels: int, groups: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=1e-6)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Identity()])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x.view(b, c, h * w)).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        p = torch.softmax(q @ k.transpose(1, 2) / (c ** 0.5), dim=-1)
        o = self.to_out[0](p @ v)
        return x + o.transpose(1, 2).reshape(b, c, h, w)


class OracleMidBlock(nn.Module):
    def __init__(self, channels: int, groups: int):
        super().__init__()
        self.resnets = nn.ModuleList([OracleResnetBlock2D(channels, channels, groups) for _ in range(2)])
        self.attentions = nn.ModuleList([OracleVAEAttention(channels, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class OracleUpsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class OracleUpDecoderBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, num_layers: int, groups: int, add_upsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([OracleResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([OracleUpsample2D(cout)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class OracleDecoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        ch = list(reversed(cfg.block_out_channels))
        g = cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, ch[0], 3, padding=1)
        self.mid_block = OracleMidBlock(ch[0], g)
        blocks, prev = [], ch[0]
        for i, c in enumerate(ch):
            blocks.append(OracleUpDecoderBlock2D(prev, c, cfg.layers_per_block + 1, g, add_upsample=i != len(ch) - 1))
            prev = c
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(g, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class OracleAutoencoderKL(nn.Module):
    """decode-only AutoencoderKL (SD3: no quant / post-quant convs)"""

    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.cfg = cfg
        self.decoder = OracleDecoder(cfg)
        self.requires_grad_(False).eval()

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        return self.decoder(z)

    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        """modeling_sd3_pnt.py:653-654: un-scale, decode -> image in [-1, 1] nominal range, (B, 3, 8h, 8w)."""
        return self.decode(latents / self.cfg.scaling_factor + self.cfg.shift_factor)


def postprocess_uint8(image: torch.Tensor) -> torch.Tensor:
    """VaeImageProcessor.postprocess(..., output_type='pil') up to the PIL wrapper: (B, 3, H, W) -> uint8 (B, H, W, 3)."""
    x = (image / 2 + 0.5).clamp(0, 1)
    return (x.permute(0, 2, 3, 1) * 255).round().to(torch.uint8)


def build_vae(cfg: VAEConfig, seed: int = 4321) -> OracleAutoencoderKL:
    """PyTorch default Conv2d / Linear / GroupNorm init under manual_seed(seed) (no checkpoint is reachable offline)."""
    torch.manual_seed(seed)
    return OracleAutoencoderKL(cfg)


def decode_flops(cfg: VAEConfig, h: int, w: int) -> float:
    """algorithmic FLOPs of one decode of an (h, w) latent (convolutions + attention GEMMs, 2 per MAC)"""
    ch = list(reversed(cfg.block_out_channels))
    conv = lambda cin, cout, hh, ww, k=3: 2.0 * hh * ww * cin * cout * k * k
    f = conv(cfg.latent_channels, ch[0], h, w)
    f += 4 * conv(ch[0], ch[0], h, w)                                            # two mid resnets
    n = h * w
    f += 4 * 2.0 * n * ch[0] * ch[0] + 2 * 2.0 * n * n * ch[0]                   # q, k, v, out + QK^T + PV
    prev = ch[0]
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block + 1):
            cin = prev if j == 0 else c
            f += conv(cin, c, h, w) + conv(c, c, h, w) + (conv(cin, c, h, w, 1) if cin != c else 0)
        prev = c
        if i != len(ch) - 1:
            h, w = 2 * h, 2 * w
            f += conv(c, c, h, w)
    return f + conv(ch[-1], cfg.out_channels, h, w)
