"""Diagnostic sweep of the tcgen05 kernels against torch fp32 on the GPU box (prints, never asserts)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpdm_b200 import _lib as L  # noqa: E402

lib = L.load()
dev = "cuda"
torch.manual_seed(0)


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def gemm_case(batch, rows, N, K, epi, time_it=False):
    A = (torch.randn(batch, rows, K, device=dev) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    gate = torch.randn(batch, N, device=dev)
    acc = A.float() @ W.float().t() + bias
    if epi == 0:
        out = torch.zeros(batch, rows, N, device=dev, dtype=torch.bfloat16); ref = acc
    elif epi == 1:
        out = torch.zeros(batch, rows, N, device=dev); ref = acc
    elif epi == 2:
        out = torch.zeros(batch, rows, N, device=dev, dtype=torch.bfloat16); ref = torch.nn.functional.gelu(acc, approximate="tanh")
    else:
        out = torch.randn(batch, rows, N, device=dev); ref = out + gate[:, None, :] * acc
    st = lib.tpdm_gemm_bf16(L.ptr(A), L.ptr(W), L.ptr(bias), L.ptr(gate), L.ptr(out), batch, rows, N, K, epi, None)
    torch.cuda.synchronize()
    msg = f"gemm b={batch} rows={rows} N={N} K={K} epi={epi}: status={st} rel={rel(out, ref):.3e} maxabs={float((out.float()-ref).abs().max()):.3e}"
    if time_it:
        ms = timeit(lambda: lib.tpdm_gemm_bf16(L.ptr(A), L.ptr(W), L.ptr(bias), L.ptr(gate), L.ptr(out), batch, rows, N, K, epi, None))
        msg += f"  {ms*1e3:.1f} us  {2*batch*rows*N*K/ms/1e9:.1f} TFLOP/s"
    print(msg, flush=True)
    if rel(out, ref) > 2e-2:
        print("   out[0,0,:8] =", out[0, 0, :8].float().tolist())
        print("   ref[0,0,:8] =", ref[0, 0, :8].float().tolist())
        print("   out[0,5,64:72] =", out[0, min(5, rows - 1), 64:72].float().tolist())
        print("   ref[0,5,64:72] =", ref[0, min(5, rows - 1), 64:72].float().tolist())


def attn_case(Bt, S, H, d, q_rows=0, time_it=False):
    dp = 64 if d <= 64 else 128
    qkv = torch.zeros(Bt, S, 3, H, dp, device=dev)
    qkv[..., :d] = torch.randn(Bt, S, 3, H, d, device=dev)
    qkv = qkv.bfloat16().contiguous()
    out = torch.zeros(Bt, S, H, dp, device=dev, dtype=torch.bfloat16)
    q, k, v = (qkv[:, :, i, :, :d].float().transpose(1, 2) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2)  # Bt,S,H,d
    st = lib.tpdm_joint_attention(L.ptr(qkv), L.ptr(out), Bt, S, H, dp, d, q_rows, None)
    torch.cuda.synchronize()
    rows = q_rows if q_rows else S
    o = out[:, :rows, :, :d]
    msg = f"attn Bt={Bt} S={S} H={H} d={d} q_rows={q_rows}: status={st} rel={rel(o, ref[:, :rows]):.3e} pad_abs={float(out[..., d:].float().abs().max()) if dp > d else 0:.1e}"
    if time_it:
        ms = timeit(lambda: lib.tpdm_joint_attention(L.ptr(qkv), L.ptr(out), Bt, S, H, dp, d, q_rows, None))
        msg += f"  {ms*1e3:.1f} us  {4*Bt*H*S*S*d/ms/1e9:.1f} TFLOP/s"
    print(msg, flush=True)
    if rel(o, ref[:, :rows]) > 2e-2:
        print("   out[0,0,0,:8] =", out[0, 0, 0, :8].float().tolist())
        print("   ref[0,0,0,:8] =", ref[0, 0, 0, :8].float().tolist())
        print("   out[0,S-1,0,:8] =", out[0, S - 1, 0, :8].float().tolist())
        print("   ref[0,S-1,0,:8] =", ref[0, S - 1, 0, :8].float().tolist())


def conv_case(B, g, Cc, N):
    x = torch.randn(B, Cc, g, g, device=dev).bfloat16()
    w = (torch.randn(N, Cc, 3, 3, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    ref = torch.nn.functional.conv2d(x.float(), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(B, g * g, N)
    xn = x.permute(0, 2, 3, 1).contiguous()
    wp = w.permute(0, 2, 3, 1).reshape(N, 9 * Cc).contiguous()
    out = torch.zeros(B, g * g, N, device=dev)
    st = lib.tpdm_conv3x3_nhwc(L.ptr(xn), L.ptr(wp), L.ptr(bias), L.ptr(out), B, g, Cc, N, None)
    torch.cuda.synchronize()
    print(f"conv3x3 B={B} g={g} C={Cc} N={N}: status={st} rel={rel(out, ref):.3e}", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "gemm"):
        gemm_case(1, 128, 256, 64, 1)
        gemm_case(1, 128, 256, 256, 1)
        gemm_case(1, 256, 512, 512, 0)
        gemm_case(2, 333, 384, 384, 0)
        gemm_case(2, 333, 1152, 384, 2)
        gemm_case(2, 256, 384, 1536, 3)
        gemm_case(2, 256, 64, 384, 1)
        gemm_case(2, 4096, 4608, 1536, 0, True)
        gemm_case(2, 4096, 6144, 1536, 2, True)
        gemm_case(2, 4096, 1536, 6144, 3, True)
        gemm_case(2, 4096, 1536, 1536, 3, True)
        gemm_case(2, 333, 1536, 1536, 3, True)
        gemm_case(2, 4096, 1536, 6144, 1, True)
    if which in ("all", "conv"):
        conv_case(2, 16, 128, 128)
        conv_case(1, 64, 3072, 128)
        conv_case(1, 128, 256, 128)
    if which in ("all", "attn"):
        attn_case(1, 128, 1, 64)
        attn_case(1, 256, 2, 64)
        attn_case(2, 589, 4, 96)
        attn_case(1, 1357, 4, 64)
        attn_case(1, 1357, 4, 64, q_rows=1024)
        attn_case(2, 4429, 24, 64, 0, True)
    if which == "gemm_ncu":   # one launch per SD3-medium block GEMM shape, for an `ncu --set full` capture
        gemm_case(2, 4429, 4608, 1536, 0)
        gemm_case(2, 4429, 6144, 1536, 2)
        gemm_case(2, 4429, 1536, 6144, 3)
        gemm_case(2, 4429, 1536, 1536, 3)
    if which == "vae":   # SD3 VAE decode of a 128x128 latent (1024^2 image): time + GEMM-class share
        import ctypes as C
        from tpdm_b200.vae import AutoencoderKL
        torch.manual_seed(4321)
        vae = AutoencoderKL(device="cuda", dtype=torch.float32)
        side = int(sys.argv[2]) if len(sys.argv) > 2 else 128
        lat = torch.randn(1, 16, side, side, device="cuda")
        fn = lambda: vae.decode_latents(lat, "uint8")
        fn()
        ms = timeit(fn, 5)
        L.check(lib.tpdm_profile_start(4096))
        fn()
        torch.cuda.synchronize()
        pms, pfl, pct = (C.c_double * 2)(), (C.c_double * 2)(), (C.c_longlong * 2)()
        L.check(lib.tpdm_profile_stop(pms, pfl, pct, 2))
        flops = 10.472e12 * (side / 128) ** 2 if side == 128 else float("nan")
        print(f"vae decode {side*8}^2: {ms:.2f} ms  ({flops/ms/1e9:.0f} TFLOP/s algorithmic); GEMM-class kernels {pms[0]:.2f} ms over {pct[0]} launches "
              f"({pfl[0]/pms[0]/1e9:.0f} TFLOP/s), everything else {ms - pms[0]:.2f} ms; workspace {vae._workspace.numel()/2**30:.2f} GiB", flush=True)
    if which == "gemm_sustained":   # 4 s back to back: our FF1 GEMM vs torch.matmul (cuBLAS) on the same shape, with clocks
        import subprocess, threading
        rows, N, K = (int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (8858, 6144, 1536)
        A = (torch.randn(rows, K, device=dev) * 0.5).bfloat16()
        W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        bias = torch.zeros(N, device=dev)
        out = torch.empty(rows, N, device=dev, dtype=torch.bfloat16)
        def ours():
            lib.tpdm_gemm_bf16(L.ptr(A), L.ptr(W), L.ptr(bias), None, L.ptr(out), 1, rows, N, K, 0, None)
        def cublas():
            torch.matmul(A, W.t(), out=out)
        for name, fn in (("tpdm gemm2", ours), ("torch.matmul", cublas), ("tpdm gemm2", ours), ("torch.matmul", cublas)):
            clocks = []
            stop = [False]
            def sample():
                while not stop[0]:
                    r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True)
                    try:
                        c, pw = r.stdout.strip().split(",")
                        clocks.append((float(c), float(pw)))
                    except Exception:
                        pass
                    time.sleep(0.2)
            th = threading.Thread(target=sample, daemon=True)
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            th.start()
            n, t0 = 0, time.perf_counter()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            while time.perf_counter() - t0 < 4.0:
                for _ in range(200):
                    fn()
                n += 200
                torch.cuda.synchronize()
            e1.record()
            torch.cuda.synchronize()
            stop[0] = True
            th.join()
            ms = e0.elapsed_time(e1)
            late = clocks[len(clocks) // 2:]
            mhz = sorted(c for c, _ in late)[len(late) // 2] if late else float("nan")
            pw = max((p for _, p in late), default=float("nan"))
            print(f"{name:13s} {rows}x{N}x{K}: {2.0*rows*N*K*n/ms/1e9:7.1f} TFLOP/s sustained over {ms/1e3:.1f} s, median SM clock {mhz:.0f} MHz, max power {pw:.0f} W", flush=True)
    if which == "attn_sustained":   # 4 s of the SD3-medium attention kernel back to back, and of cuDNN SDPA on the same shape: clocks and power
        import subprocess, threading
        Bt, S, H, d = 2, 4429, 24, 64
        qkv = torch.randn(Bt, S, 3, H, d, device=dev).bfloat16().contiguous()
        out = torch.zeros(Bt, S, H, d, device=dev, dtype=torch.bfloat16)
        qf, kf, vf = (qkv[:, :, i].transpose(1, 2).contiguous() for i in range(3))
        def ours():
            lib.tpdm_joint_attention(L.ptr(qkv), L.ptr(out), Bt, S, H, 64, d, 0, None)
        def sdpa():
            torch.nn.functional.scaled_dot_product_attention(qf, kf, vf)
        for name, fn in (("tpdm attention", ours), ("torch SDPA", sdpa), ("tpdm attention", ours)):
            clocks, stop = [], [False]
            def sample():
                while not stop[0]:
                    r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True)
                    try:
                        c, pw = r.stdout.strip().split(",")
                        clocks.append((float(c), float(pw)))
                    except Exception:
                        pass
                    time.sleep(0.2)
            th = threading.Thread(target=sample, daemon=True)
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            th.start()
            n, t0 = 0, time.perf_counter()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            while time.perf_counter() - t0 < 4.0:
                for _ in range(100):
                    fn()
                n += 100
                torch.cuda.synchronize()
            e1.record()
            torch.cuda.synchronize()
            stop[0] = True
            th.join()
            ms = e0.elapsed_time(e1)
            late = clocks[len(clocks) // 2:]
            mhz = sorted(c for c, _ in late)[len(late) // 2] if late else float("nan")
            pw = sorted(p for _, p in late)[len(late) // 2] if late else float("nan")
            print(f"{name:15s} Bt={Bt} H={H} S={S} d={d}: {ms / n * 1e3:7.1f} us per launch = {4.0*Bt*H*S*S*d*n/ms/1e9:7.1f} TFLOP/s sustained over {ms/1e3:.1f} s, "
                  f"median SM clock {mhz:.0f} MHz, median power {pw:.0f} W", flush=True)
    if which == "ln":
        for (batch, rows) in ((2, 4096), (2, 333)):
            D = 1536
            x = torch.randn(batch, rows, D, device="cuda")
            mod = torch.randn(batch, 2 * D, device="cuda") * 0.1
            out = torch.empty(batch, rows, D, device="cuda", dtype=torch.bfloat16)
            fn = lambda: L.check(lib.tpdm_ln_modulate(L.ptr(x), L.ptr(mod), L.ptr(mod) + 4 * D, 2 * D, L.ptr(out), batch, rows, D,
                                                        torch.cuda.current_stream().cuda_stream))
            fn()
            ref = torch.nn.functional.layer_norm(x, (D,), eps=1e-6) * (1 + mod[:, None, D:]) + mod[:, None, :D]
            big = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
            ms = timeit(fn, 20)
            def cold():
                big.zero_()
                fn()
            ms_cold = timeit(cold, 10) - timeit(lambda: big.zero_(), 10)
            gb = batch * rows * D * 6 / 1e9
            print(f"ln batch={batch} rows={rows}: rel={rel(out, ref):.2e}  warm {ms*1e3:.1f} us ({gb/ms*1e3:.0f} GB/s)  "
                  f"after L2 flush {ms_cold*1e3:.1f} us ({gb/ms_cold*1e3:.0f} GB/s)", flush=True)
    if which == "gemm_k64":
        for epi in (0, 1, 2, 3):
            gemm_case(2, 4096, 1536, 64, epi, True)      # epilogue-bound: almost no MMA work
        for epi in (0, 3):
            gemm_case(2, 4096, 1536, 512, epi, True)
    if which == "gemm_epi":
        for epi in (0, 1, 2, 3):
            gemm_case(2, 4096, 1536, 1536, epi, True)
        for epi in (0, 3):
            gemm_case(1, 9472, 1536, 1536, epi, True)   # 74 x 6 = 444 tiles = 3 full waves
            gemm_case(1, 4736, 1536, 1536, epi, True)   # 222 tiles = 1.5 waves
            gemm_case(1, 2368 * 8, 1536, 1536, epi, True)   # 6 waves
    if which == "wgrad":
        cases = ((1, 16, 256),) if len(sys.argv) > 2 else ((1, 16, 256), (3, 16, 256), (2, 32, 512), (1, 64, 3072))
        for (ns, g, Cc) in cases:
            M = 128
            x = torch.randn(ns, Cc, g, g, device=dev).bfloat16()
            dy = torch.randn(ns, M, g, g, device=dev).bfloat16()
            xs = torch.zeros(ns, 3, Cc, g, g, device=dev, dtype=torch.bfloat16)
            xs[:, 1] = x
            xs[:, 0, :, :, 1:] = x[:, :, :, :-1]
            xs[:, 2, :, :, :-1] = x[:, :, :, 1:]
            w = torch.zeros(M, Cc, 3, 3, device=dev, requires_grad=True)
            torch.nn.functional.conv2d(x.float(), w, padding=1).backward(dy.float())
            ref = w.grad.permute(0, 2, 3, 1).reshape(M, 9 * Cc)
            out = torch.full((M, 9 * Cc), 7.0, device=dev)
            st = lib.tpdm_conv3x3_wgrad(L.ptr(dy.reshape(ns, M, g * g).contiguous()), L.ptr(xs), L.ptr(out), ns, g, Cc, M, None)
            torch.cuda.synchronize()
            print(f"wgrad ns={ns} g={g} C={Cc}: status={st} rel={rel(out, ref):.3e}", flush=True)
    if which == "attn_lib":   # library bars for the joint attention of one SD3-medium block: (Bt, H, S, d) = (2, 24, 4429, 64), bf16
        from torch.nn.attention import SDPBackend, sdpa_kernel
        for (Bt, H, S, d) in ((2, 24, 4429, 64), (2, 24, 16717, 64), (32, 24, 1357, 64)):
            flops = 4.0 * Bt * H * S * S * d
            q, k, v = (torch.randn(Bt, H, S, d, device=dev, dtype=torch.bfloat16) for _ in range(3))
            for name, be in (("torch SDPA cuDNN", SDPBackend.CUDNN_ATTENTION), ("torch SDPA flash", SDPBackend.FLASH_ATTENTION),
                             ("torch SDPA mem-efficient", SDPBackend.EFFICIENT_ATTENTION)):
                try:
                    with sdpa_kernel(be):
                        ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), 20)
                    print(f"{name:26s} Bt={Bt} H={H} S={S} d={d}: {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s", flush=True)
                except Exception as e:
                    print(f"{name:26s} Bt={Bt} H={H} S={S} d={d}: unavailable ({str(e).splitlines()[0][:90]})", flush=True)
            try:
                from flash_attn import flash_attn_func
                qf, kf, vf = (t.transpose(1, 2).contiguous() for t in (q, k, v))
                ms = timeit(lambda: flash_attn_func(qf, kf, vf), 20)
                print(f"{'flash_attn 2 (mma.sync)':26s} Bt={Bt} H={H} S={S} d={d}: {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s", flush=True)
            except Exception as e:
                print(f"flash_attn unavailable ({str(e).splitlines()[0][:90]})", flush=True)
            qkv = torch.randn(Bt, S, 3, H, d, device=dev).bfloat16().contiguous()
            out = torch.zeros(Bt, S, H, d, device=dev, dtype=torch.bfloat16)
            ms = timeit(lambda: lib.tpdm_joint_attention(L.ptr(qkv), L.ptr(out), Bt, S, H, 64, d, 0, None), 20)
            print(f"{'tpdm joint_attention':26s} Bt={Bt} H={H} S={S} d={d}: {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s", flush=True)
            del q, k, v, qkv, out
    if which == "attn_time":   # the SD3-medium 1024^2 block shape only, timed (kernel experiments: TPDM_B200_LIB=... variants)
        attn_case(2, 4429, 24, 64, 0, True)
    if which == "attn_big":
        attn_case(2, 4429, 24, 64)
        attn_case(2, 4429, 24, 64)
    if which == "gemm_big":
        gemm_case(2, 4096, 6144, 1536, 2)
        gemm_case(2, 4096, 1536, 1536, 3)
        gemm_case(2, 4096, 1536, 6144, 3)
    print("diag done", flush=True)
