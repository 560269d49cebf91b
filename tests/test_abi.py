"""CPU tests: the C-ABI library builds, loads and exports every symbol include/tpdm_b200.h declares; host-side error
behaviour that needs no GPU."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tpdm_b200 import build
    from tpdm_b200 import _lib as L

    build.build()
    return L.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "tpdm_b200.h")).read()
    declared = set(re.findall(r"\b(tpdm_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"tpdm_status"}
    assert len(declared) >= 18
    from tpdm_b200 import _lib as L

    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_struct_sizes(lib):
    from tpdm_b200 import _lib as L

    assert lib.tpdm_abi_version() == 3
    assert C.sizeof(L.TpdmConfig) == 13 * 4 + 3 * 4 + 8          # 64 bytes, the uint64 mask is naturally aligned
    assert C.sizeof(L.TpdmBlockWeights) == 26 * 8
    assert C.sizeof(L.TpdmWeights) == (17 + 1 + 12) * 8
    assert C.sizeof(L.TpdmSampleState) == 11 * 8


def test_sass_is_blackwell_native():
    """UTCHMMA = tcgen05.mma, UTMALDG = TMA, LDTM/STTM = tcgen05.ld/st (B200_PROFILING.md)."""
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    from tpdm_b200 import _lib as L

    sass = subprocess.run([exe, "-sass", L.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass  # no legacy mma.sync path


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(lib):
    """Without a GPU the product path must fail loudly, never compute on the CPU."""
    from tpdm_b200 import _lib as L
    from tpdm_b200.modeling_sd3_pnt import TimePredictor
    from tpdm_b200.transformer_sd3 import CustomSD3Transformer2DModel

    cfg = L.TpdmConfig(num_layers=1, num_heads=1, head_dim=64, joint_attention_dim=4096, pooled_projection_dim=2048, in_channels=16,
                       out_channels=16, patch_size=2, pos_embed_max_size=96, qk_norm=0, tpm_channels=128, prediction_type=0, relative=1,
                       min_sigma=1e-3, epsilon=1e-3, tpm_epsilon=1.0)
    h = L.vp()
    st = lib.tpdm_create(C.byref(cfg), C.byref(h))
    assert st == L.TPDM_ERR_CUDA and b"no CPU fallback" in lib.tpdm_last_error()
    m = CustomSD3Transformer2DModel(sample_size=32, num_layers=1, attention_head_dim=64, num_attention_heads=1, caption_projection_dim=64)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 16, 32, 32), torch.zeros(1, 8, 4096), torch.zeros(1, 2048), torch.zeros(1))
    with pytest.raises(RuntimeError):
        TimePredictor(128, 128)(torch.zeros(1, 128, 16, 16), torch.zeros(1, 64))


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "tpdm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"


def test_state_dict_names_match_reference_layout():
    """Same parameter names as the diffusers layout the reference checkpoints use."""
    from oracle import sd3_oracle as O
    from tpdm_b200.modeling_sd3_pnt import TimePredictor
    from tpdm_b200.transformer_sd3 import CustomSD3Transformer2DModel

    for qk in (None, "rms_norm"):
        cfg = O.tiny_config(qk_norm=qk)
        ora = O.OracleSD3Transformer(cfg)
        ours = CustomSD3Transformer2DModel(sample_size=32, num_layers=2, attention_head_dim=96, num_attention_heads=4,
                                           caption_projection_dim=384, pos_embed_max_size=96, qk_norm=qk)
        a, b = ora.state_dict(), ours.state_dict()
        assert set(a) == set(b), set(a) ^ set(b)
        assert all(a[k].shape == b[k].shape for k in a)
        assert torch.equal(a["pos_embed.pos_embed"], b["pos_embed.pos_embed"])
    # SD3.5 dual-attention blocks (transformer_sd3.py:104-106,138): attn2.* and a 9-chunk norm1 in the named layers only
    cfg = O.tiny_config(qk_norm="rms_norm")
    cfg.dual_attention_layers = (0,)
    a = O.OracleSD3Transformer(cfg).state_dict()
    b = CustomSD3Transformer2DModel(sample_size=32, num_layers=2, attention_head_dim=96, num_attention_heads=4, caption_projection_dim=384,
                                    pos_embed_max_size=96, qk_norm="rms_norm", dual_attention_layers=(0,)).state_dict()
    assert set(a) == set(b) and all(a[k].shape == b[k].shape for k in a)
    assert a["transformer_blocks.0.norm1.linear.weight"].shape[0] == 9 * 384 and a["transformer_blocks.1.norm1.linear.weight"].shape[0] == 6 * 384
    assert "transformer_blocks.0.attn2.norm_q.weight" in b and "transformer_blocks.1.attn2.to_q.weight" not in b
    with pytest.raises(ValueError):
        CustomSD3Transformer2DModel(sample_size=32, num_layers=2, attention_head_dim=96, num_attention_heads=4, caption_projection_dim=384,
                                    dual_attention_layers=(2,))
    tp = TimePredictor(128, 768)
    assert set(tp.state_dict()) == set(O.OracleTimePredictor(128, 768).state_dict())
    assert sorted(tp.state_dict()) == sorted(
        ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias",
         "norm1.linear.weight", "norm1.linear.bias", "norm1.norm.weight", "norm1.norm.bias"])
    assert float(tp.fc2.bias[0]) == 1.5 and float(tp.fc2.bias[1]) == 0.5


def test_output_container_access_styles():
    from tpdm_b200.model_utilis import CustomDiffusionModelOutput

    o = CustomDiffusionModelOutput(init_noise_latents=torch.zeros(1), hidden_states_combineds=None, tembs=torch.zeros(1), images=[],
                                   last_valid_indices=[], alphas=torch.ones(1), betas=torch.ones(1), sigmas=torch.ones(1, 2),
                                   logprobs=torch.zeros(1), prob_masks=torch.zeros(1).bool())
    assert o.sigmas is o["sigmas"] and o.get("images") == [] and "alphas" in dict(o.items())
    assert o.get("hidden_states_combineds", None) is None


def test_vae_module_matches_diffusers_state_dict_layout():
    """the product AutoencoderKL holds exactly the decoder tensors of the diffusers layout (names shared with the oracle)"""
    from oracle import vae_oracle as V
    from tpdm_b200.vae import AutoencoderKL

    for cfg in (V.sd3_vae_config(), V.tiny_vae_config()):
        ora = V.build_vae(cfg)
        vae = AutoencoderKL(latent_channels=cfg.latent_channels, block_out_channels=cfg.block_out_channels,
                            layers_per_block=cfg.layers_per_block, norm_num_groups=cfg.norm_num_groups)
        a, b = ora.state_dict(), vae.state_dict()
        assert set(a) == set(b) and all(a[k].shape == b[k].shape for k in a)
    assert "decoder.up_blocks.2.resnets.0.conv_shortcut.weight" in a or "decoder.up_blocks.1.resnets.0.conv_shortcut.weight" in a
    assert vae.upscale == 2 and AutoencoderKL().upscale == 8


def test_vae_has_no_cpu_path():
    import pytest
    import torch
    from tpdm_b200.vae import AutoencoderKL

    vae = AutoencoderKL(block_out_channels=(64, 128), layers_per_block=1, norm_num_groups=16)
    with pytest.raises((RuntimeError, ValueError, OSError)):
        vae.decode(torch.zeros(1, 16, 8, 8))


def test_vae_oracle_shapes_and_flops():
    import torch
    from oracle import vae_oracle as V

    t = V.build_vae(V.tiny_vae_config())
    img = t.decode_latents(torch.randn(1, 16, 8, 8, generator=torch.Generator().manual_seed(0)))
    assert img.shape == (1, 3, 16, 16) and bool(torch.isfinite(img).all())
    u8 = V.postprocess_uint8(img)
    assert u8.shape == (1, 16, 16, 3) and u8.dtype == torch.uint8
    assert abs(V.decode_flops(V.sd3_vae_config(), 128, 128) / 1e12 - 10.47) < 0.05


def test_vae_create_validates_topology_without_a_gpu(lib):
    """tpdm_vae_create is host-only: bad topologies are refused with TPDM_ERR_SHAPE / TPDM_ERR_ARG and a message; a valid SD3
    configuration reports 2 + 4 * 3 resnets."""
    from tpdm_b200 import _lib as L

    def cfg(channels, groups=32, layers=2, latent=16):
        c = L.TpdmVaeConfig(latent_channels=latent, out_channels=3, num_levels=len(channels), layers_per_block=layers,
                            norm_num_groups=groups, scaling_factor=1.5305, shift_factor=0.0609)
        for i, v in enumerate(channels):
            c.block_out_channels[i] = v
        return c

    h = L.vp()
    assert lib.tpdm_vae_create(C.byref(cfg((128, 256, 512, 512))), C.byref(h)) == 0
    assert lib.tpdm_vae_num_resnets(h) == 14
    assert lib.tpdm_vae_workspace_bytes(h, 128, 128) > 3 * 2**30        # 4 activation buffers of 512 MiB + attention scratch
    assert lib.tpdm_vae_workspace_bytes(h, 0, 128) == 0
    assert lib.tpdm_vae_destroy(h) == 0
    bad = L.vp()
    assert lib.tpdm_vae_create(C.byref(cfg((96, 192))), C.byref(bad)) == L.TPDM_ERR_SHAPE      # not multiples of 64
    assert b"block_out_channels" in lib.tpdm_last_error()
    assert lib.tpdm_vae_create(C.byref(cfg((64, 128), groups=32)), C.byref(bad)) == L.TPDM_ERR_SHAPE   # 2 channels per group
    assert lib.tpdm_vae_create(C.byref(cfg((128,), latent=80)), C.byref(bad)) == L.TPDM_ERR_SHAPE
    assert lib.tpdm_vae_create(None, C.byref(bad)) == L.TPDM_ERR_ARG
