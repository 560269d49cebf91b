"""BASELINE config 5: one SD3-medium MMDiT forward at 2048^2 (latent 256x256 -> 16 384 image tokens + 333 text tokens,
S = 16 717), batch 1 with CFG (Bt = 2), bf16.  Prints one JSON line: step ms, MMDiT TFLOP/s and the per-class (GEMM /
attention) device times from the library's per-launch CUDA-event brackets.  Run on a B200:
    python tools/run_config5_2048.py [iters]"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpdm_b200 import _lib as L  # noqa: E402
from tpdm_b200.modeling_sd3_pnt import SD3_MEDIUM_TRANSFORMER_CONFIG  # noqa: E402
from tpdm_b200.transformer_sd3 import CustomSD3Transformer2DModel  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    lib = L.load()
    torch.manual_seed(1234)
    cfg = dict(SD3_MEDIUM_TRANSFORMER_CONFIG)
    cfg["sample_size"] = 256
    model = CustomSD3Transformer2DModel(**cfg).to(device="cuda", dtype=torch.bfloat16)
    g = torch.Generator(device="cuda").manual_seed(5)
    lat = torch.randn(1, 16, 256, 256, device="cuda", generator=g).repeat(2, 1, 1, 1)
    enc = torch.randn(2, 333, 4096, device="cuda", generator=g)
    pooled = torch.randn(2, 2048, device="cuda", generator=g)
    ts = torch.tensor([500.0, 500.0], device="cuda")
    fwd = lambda: model(lat, enc, pooled, ts, return_dict=False)
    for _ in range(2):
        fwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fwd()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    L.check(lib.tpdm_profile_start(4096))
    fwd()
    torch.cuda.synchronize()
    pms, pfl, pct = (C.c_double * 2)(), (C.c_double * 2)(), (C.c_longlong * 2)()
    L.check(lib.tpdm_profile_stop(pms, pfl, pct, 2))
    D, Ln, N, T, Bt = 1536, 24, 16384, 333, 2
    S = N + T
    flops = 0.0
    for i in range(Ln):
        lin = (N + T) * 12 * D * D if i < Ln - 1 else N * 12 * D * D + T * 3 * D * D
        flops += 2.0 * Bt * lin + 4.0 * Bt * S * S * D
    line = dict(config="cfg5 SD3-medium 2048^2, B=1 with CFG (Bt=2), S=16717, bf16, one MMDiT forward through "
                       "CustomSD3Transformer2DModel.forward (includes the sigma-independent text branch)",
                ms_per_forward=ms, algorithmic_tflop=flops / 1e12, tflops=flops / ms / 1e9, frac_of_sustained_peak=flops / ms / 1e9 / 1371.6,
                gemm=dict(ms=pms[0], launches=int(pct[0]), tflops=pfl[0] / pms[0] / 1e9),
                attention=dict(ms=pms[1], launches=int(pct[1]), tflops=pfl[1] / pms[1] / 1e9),
                finite=bool(torch.isfinite(out[0]).all()))
    print(json.dumps(line))


if __name__ == "__main__":
    main()
