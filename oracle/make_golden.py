"""Mint tests/golden/*.safetensors.  Run in the build container:  python -m oracle.make_golden

1. Pins the oracle restatements bit-exactly against the definitions AST-extracted from /root/reference
   (TimePredictor, CustomAdaGroupNormZeroSingle, reshape_hidden_states_to_2d, custom_step, get_ref_beta, get_kl_beta)
   -- aborts if any differs.
2. Writes small fixtures produced by the *reference's own code* (tpm_ref.safetensors, pieces_ref.safetensors) and
   by the fp32 oracle for BASELINE.json configs[0] (tiny_traj.safetensors, tiny_block.safetensors).

Inputs / weights are not stored (they are regenerated from seeds, see sd3_oracle.build_pipeline /
synthetic_inputs); a per-tensor checksum of both IS stored so a test can tell an RNG drift from a real mismatch.
"""
from __future__ import annotations

import os
import sys

import torch
from safetensors.torch import save_file

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_extract  # noqa: E402
from oracle import sd3_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def checksum(tensors) -> torch.Tensor:
    return torch.stack([t.double().abs().sum() for t in tensors]).to(torch.float64)


def pin_against_reference():
    R = ref_extract.load_reference_pieces()
    torch.manual_seed(7)
    ref_tpm = R["TimePredictor"](128, 128, 2, 1.5, 0.5)
    torch.manual_seed(7)
    ora_tpm = O.OracleTimePredictor(128, 128, 2, 1.5, 0.5)
    sd_r, sd_o = ref_tpm.state_dict(), ora_tpm.state_dict()
    assert list(sd_r.keys()) == list(sd_o.keys()), "TPM state-dict names differ from the reference"
    assert all(torch.equal(sd_r[k], sd_o[k]) for k in sd_r), "TPM init differs from the reference"
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 128, 16, 16, generator=g)
    temb = torch.randn(2, 64, generator=g)
    y_r, y_o = ref_tpm(x, temb), ora_tpm(x, temb)
    assert torch.equal(y_r, y_o), "TPM forward differs from the reference"
    # grads of sum(log(alpha,beta)) w.r.t. TPM parameters (row R1 / K20)
    y_r.log().sum().backward()
    grads = {f"grad.{k}": p.grad.clone() for k, p in ref_tpm.named_parameters()}
    h = torch.randn(2, 256, 384, generator=g)
    assert torch.equal(R["reshape_hidden_states_to_2d"](h, 16, 16), O.reshape_hidden_states_to_2d(h, 16, 16))
    # scramble map: which token lands on which pixel
    idx = torch.arange(64, dtype=torch.float32).reshape(1, 64, 1)
    scramble8 = R["reshape_hidden_states_to_2d"](idx, 8, 8)[0, 0]
    mo, x0 = torch.randn(2, 16, 8, 8, generator=g), torch.randn(2, 16, 8, 8, generator=g)
    s, sn = torch.tensor([1.0, 0.6]), torch.tensor([0.7, 0.33])
    e_r = R["custom_step_body"](mo, sn, s, x0)
    assert torch.equal(e_r, O.custom_step(mo, sn, s, x0))
    sig = torch.linspace(0.02, 1.0, 29)
    a_r, b_r = R["get_ref_beta"](sig)
    a_o, b_o = O.get_ref_beta(sig)
    assert torch.equal(a_r, a_o) and torch.equal(b_r, b_o)
    kl = R["get_kl_beta"](torch.tensor(2.0), torch.tensor(5.0), torch.tensor(3.0), torch.tensor(4.0))
    assert torch.equal(kl, O.get_kl_beta(torch.tensor(2.0), torch.tensor(5.0), torch.tensor(3.0), torch.tensor(4.0)))
    save_file({"x": x, "temb": temb, "alpha_beta": y_r.detach(), **{k: v for k, v in sd_r.items()}, **grads},
              os.path.join(OUT, "tpm_ref.safetensors"))
    save_file({"scramble8": scramble8, "euler.model_output": mo, "euler.sample": x0, "euler.sigma": s,
               "euler.sigma_next": sn, "euler.prev": e_r, "refbeta.sigma": sig, "refbeta.alpha": a_r, "refbeta.beta": b_r,
               "kl_2_5_3_4": kl.reshape(1)}, os.path.join(OUT, "pieces_ref.safetensors"))
    print("pinned: TPM / reshape / custom_step / get_ref_beta / get_kl_beta are bit-identical to /root/reference")


def tiny_fixtures():
    for qk in (None, "rms_norm"):
        cfg = O.tiny_config(qk_norm=qk)
        pipe = O.build_pipeline(cfg)
        inp = O.synthetic_inputs(cfg, batch=2)
        out = pipe(**inp, max_inference_steps=8, guidance_scale=7.0, predict=True, record_velocity=True)
        tag = "tiny_traj" if qk is None else "tiny_traj_qknorm"
        fx = {
            "weights_checksum": checksum(pipe.state_dict().values()),
            "inputs_checksum": checksum(inp.values()),
            "sigmas": out["sigmas"], "alphas": out["alphas"], "betas": out["betas"], "logprobs": out["logprobs"],
            "prob_masks": out["prob_masks"].to(torch.uint8), "tembs": out["tembs"],
            "velocities": out["velocities"], "final_latents": out["final_latents"],
            "last_valid_indices": out["last_valid_indices"],
        }
        if qk is None:
            # predict=False with injected ratios (RLOO rollout shape, BASELINE config 4 in miniature)
            g = torch.Generator().manual_seed(5)
            ratios = torch.rand(2, 8, generator=g) * 0.6 + 0.2
            out_s = pipe(**inp, max_inference_steps=8, predict=False, ratios=ratios)
            fx.update({"sample.ratios": ratios, "sample.sigmas": out_s["sigmas"], "sample.logprobs": out_s["logprobs"],
                       "sample.final_latents": out_s["final_latents"]})
            # one MMDiT forward with per-block taps at sigma = (1.0, 0.37)
            lat2 = torch.cat([inp["latents"]] * 2)
            pe = torch.cat([inp["negative_prompt_embeds"], inp["prompt_embeds"]])
            pp = torch.cat([inp["negative_pooled_prompt_embeds"], inp["pooled_prompt_embeds"]])
            ts = torch.tensor([1000.0, 370.0, 1000.0, 370.0])
            with torch.no_grad():
                v, temb, h1, h2, blocks = pipe.transformer(lat2, pe, pp, ts, return_blocks=True)
            save_file({"timestep": ts, "velocity": v, "temb": temb, "h1_row0": h1[:, 0].contiguous(),
                       "h2_mean": h2.mean(dim=1), "h2_tok17": h2[:, 17].contiguous(),
                       "block0_tok5": blocks[0][:, 5].contiguous(), "block1_tok5": blocks[1][:, 5].contiguous()},
                      os.path.join(OUT, "tiny_block.safetensors"))
        save_file({k: v.contiguous() for k, v in fx.items()}, os.path.join(OUT, f"{tag}.safetensors"))
        print(tag, "sigmas[0] =", [round(float(s), 4) for s in out["sigmas"][0]])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if not ref_extract.available():
        raise SystemExit("needs /root/reference (build container only)")
    pin_against_reference()
    tiny_fixtures()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
