"""Timeline of ONE attention CTA from the instrumented build (python -m tpdm_b200.build --variant trace TPDM_ATTN_TRACE):
clock64() stamps of softmax warp 4 and of the two MMA-issuing warps, per 64-key half-tile.  Run on the GPU box:
    TPDM_B200_LIB=tpdm_b200/_build/libtpdm_trace.so python tools/attn_trace.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpdm_b200 import _lib as L  # noqa: E402

lib = L.load()
Bt, S, H, d = 2, 4429, 24, 64
qkv = torch.randn(Bt, S, 3, H, d, device="cuda").bfloat16().contiguous()
out = torch.zeros(Bt, S, H, d, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    L.check(lib.tpdm_joint_attention(L.ptr(qkv), L.ptr(out), Bt, S, H, 64, d, 0, None))
torch.cuda.synchronize()
buf = (C.c_longlong * (3 * 2048))()
lib.tpdm_attn_trace_read.argtypes = [C.POINTER(C.c_longlong), C.c_int]
assert lib.tpdm_attn_trace_read(buf, 3 * 2048) == 0
sm, qk, pv = ([buf[r * 2048 + i] for i in range(2048)] for r in range(3))
n_half = (S + 63) // 64
names = ["wait s_full/pv_done", "ld va", "exp va (+ld vb)", "wait vb", "exp vb + check", "fence+s_free", "st wait+p_full"]
print(f"softmax warp 4, half-tiles 20..27 (cycles per phase; period = top(i+1) - top(i))")
tot = [0.0] * 8
cnt = 0
for i in range(8, n_half - 2):
    e = sm[8 * i: 8 * i + 8]
    nxt = sm[8 * (i + 1)]
    if 0 in e or nxt == 0:
        continue
    ph = [e[k + 1] - e[k] for k in range(7)] + [nxt - e[0]]
    for k in range(8):
        tot[k] += ph[k]
    cnt += 1
    if 20 <= i < 28:
        print(f"  i={i:3d}: " + "  ".join(f"{names[k]}={ph[k]:4d}" for k in range(7)) + f"  | period={ph[7]}")
print("mean over", cnt, "half-tiles: " + "  ".join(f"{names[k]}={tot[k] / cnt:6.1f}" for k in range(7)) + f"  | period={tot[7] / cnt:.1f}")
for role, name, arr in ((1, "QK issuer", qk), (2, "PV issuer", pv)):
    w = s = c = 0.0
    for i in range(8, n_half - 2):
        a = arr[4 * i: 4 * i + 3]
        if 0 in a:
            continue
        w += a[1] - a[0]
        s += a[2] - a[1]
        c += 1
    print(f"{name}: mean wait {w / c:.1f} cycles, issue+commit {s / c:.1f} cycles per half-tile")
# lag between the producer and consumer of S and P
lag_s = [sm[8 * i + 1] - qk[4 * i + 2] for i in range(8, n_half - 2) if sm[8 * i + 1] and qk[4 * i + 2]]
lag_p = [pv[4 * i + 1] - sm[8 * i + 7] for i in range(8, n_half - 2) if pv[4 * i + 1] and sm[8 * i + 7]]
print(f"S(i) issued -> softmax(i) past its waits: mean {sum(lag_s) / len(lag_s):.0f} cycles (min {min(lag_s)}, max {max(lag_s)})")
print(f"p_full(i) arrive -> PV issuer past its waits: mean {sum(lag_p) / len(lag_p):.0f} cycles (min {min(lag_p)}, max {max(lag_p)})")
