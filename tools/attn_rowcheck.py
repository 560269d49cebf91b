"""Per-row error of the attention kernel against fp32 SDPA (worst rows and where they sit), for several logit scales."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpdm_b200 import _lib as L  # noqa: E402

lib = L.load()
torch.manual_seed(0)
for (Bt, S, H, d, qs, ks, vs) in ((2, 4429, 24, 64, 1.0, 1.0, 1.0), (2, 4429, 24, 64, 0.3, 0.3, 1.0), (2, 4429, 24, 64, 3.0, 3.0, 1.0), (1, 1357, 24, 64, 1.0, 1.0, 1.0)):
    qkv = torch.randn(Bt, S, 3, H, d, device="cuda")
    qkv[:, :, 0] *= qs
    qkv[:, :, 1] *= ks
    qkv[:, :, 2] = qkv[:, :, 2] * vs + 0.5          # a mean in V makes normalisation errors visible
    qkv = qkv.bfloat16().contiguous()
    out = torch.zeros(Bt, S, H, d, device="cuda", dtype=torch.bfloat16)
    q, k, v = (qkv[:, :, i].float().transpose(1, 2) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2)     # Bt, S, H, d
    L.check(lib.tpdm_joint_attention(L.ptr(qkv), L.ptr(out), Bt, S, H, 64, d, 0, None))
    torch.cuda.synchronize()
    err = (out.float() - ref)
    row_rel = err.norm(dim=-1) / ref.norm(dim=-1)                 # Bt, S, H
    mean_bias = float((out.float().mean() - ref.mean()) / ref.abs().mean())
    worst = row_rel.flatten().topk(5)
    idx = [divmod(int(i), S * H) for i in worst.indices]
    pos = [(b, r // H, r % H) for b, r in idx]
    print(f"S={S} scales q{qs} k{ks}: rel-L2 {float(err.norm() / ref.norm()):.3e}  mean row rel {float(row_rel.mean()):.3e}  worst rows {[f'{float(x):.2e}' for x in worst.values]} at (b, token, head) {pos}  relative mean bias {mean_bias:+.2e}")
    by_tile = row_rel.mean(dim=(0, 2)).reshape(-1)[: (S // 128) * 128].reshape(-1, 128).mean(1)
    print("   mean row error per 128-row query tile (first 6, last 3):", [f"{float(x):.2e}" for x in by_tile[:6]], [f"{float(x):.2e}" for x in by_tile[-3:]],
          " tail rows:", f"{float(row_rel[:, (S // 128) * 128:].mean()):.2e}")
