"""torch.matmul (cuBLAS / cuBLASLt) on the four SD3-medium block GEMM shapes, one launch each after a warm-up: the library bar next to
`tools/gpu_diag.py gemm_ncu` (run both under ncu --set full and compare tensor-pipe activity, DRAM / L2 / shared-memory traffic, grid)."""
import sys

import torch

torch.manual_seed(0)
rows = 2 * 4429
for (N, K) in ((4608, 1536), (6144, 1536), (1536, 6144), (1536, 1536)):
    A = (torch.randn(rows, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    out = torch.empty(rows, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2 if "--once" not in sys.argv else 1):
        torch.matmul(A, W.t(), out=out)
    torch.cuda.synchronize()
    print(f"matmul {rows}x{N}x{K} done", flush=True)
