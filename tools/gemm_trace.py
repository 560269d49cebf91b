"""Timeline of the CTA-pair GEMM's MMA-issuing warp from the instrumented build
(python -m tpdm_b200.build --variant gtrace TPDM_GEMM_TRACE gemm2_tcgen05.cu):
    TPDM_B200_LIB=tpdm_b200/_build/libtpdm_gtrace.so python tools/gemm_trace.py
Per k-block (64 deep = 4 MMAs of 256 x 256 x 16): cycles spent waiting for the TMA data (full barrier) and issuing + committing; per tile:
cycles waiting for a free accumulator (epilogue)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpdm_b200 import _lib as L  # noqa: E402

lib = L.load()
lib.tpdm_gemm_trace_read.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.c_int]
buf, n = (C.c_longlong * 4096)(), C.c_int(0)
torch.manual_seed(0)
for (N, K, epi, name) in ((4608, 1536, 0, "QKV"), (6144, 1536, 2, "FF1+GELU"), (1536, 6144, 3, "FF2 gate+residual"), (1536, 1536, 3, "out-proj gate+residual")):
    batch, rows = 2, 4429
    A = (torch.randn(batch, rows, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias, gate = torch.randn(N, device="cuda"), torch.randn(batch, N, device="cuda")
    out = torch.zeros(batch, rows, N, device="cuda", dtype=torch.float32 if epi in (1, 3) else torch.bfloat16)
    for _ in range(2):
        L.check(lib.tpdm_gemm_bf16(L.ptr(A), L.ptr(W), L.ptr(bias), L.ptr(gate), L.ptr(out), batch, rows, N, K, epi, None))
    torch.cuda.synchronize()
    lib.tpdm_gemm_trace_read(buf, C.byref(n), 1)
    L.check(lib.tpdm_gemm_bf16(L.ptr(A), L.ptr(W), L.ptr(bias), L.ptr(gate), L.ptr(out), batch, rows, N, K, epi, None))
    torch.cuda.synchronize()
    lib.tpdm_gemm_trace_read(buf, C.byref(n), 1)
    v = [buf[i] for i in range(n.value)]
    i, tiles, waits, issues, acc_waits, first_waits = 0, 0, [], [], [], []
    t_begin = t_end = None
    while i < len(v):
        if v[i] == -1:
            acc_waits.append(v[i + 2] - v[i + 1])
            nkb = v[i + 3]
            i += 4
            tiles += 1
            first = True
            continue
        t0, t1, t2 = v[i], v[i + 1], v[i + 2]
        t_begin = t0 if t_begin is None else t_begin
        t_end = t2
        (first_waits if first else waits).append(t1 - t0)
        first = False
        issues.append(t2 - t1)
        i += 3
    nk = len(issues)
    mean = lambda x: sum(x) / max(1, len(x))
    span = t_end - t_begin
    print(f"{name:24s} N={N} K={K}: {tiles} tiles, {nk} k-blocks of cluster 0 in {span} cycles = {span / nk:.0f} per k-block (MMA time at full rate 512)")
    print(f"    full-barrier wait: mean {mean(waits):.0f} cycles (first k-block of a tile {mean(first_waits):.0f}); issue 4 MMAs + commit: mean {mean(issues):.0f}; "
          f"accumulator wait per tile: mean {mean(acc_waits):.0f}, max {max(acc_waits)}")
    big = sorted(waits)[-max(1, len(waits) // 20):]
    print(f"    worst 5 % of the full-barrier waits: mean {mean(big):.0f} cycles; share of the span spent waiting on data {100 * (sum(waits) + sum(first_waits)) / span:.1f} %, "
          f"issuing {100 * sum(issues) / span:.1f} %, waiting for an accumulator {100 * sum(acc_waits) / span:.1f} %")
