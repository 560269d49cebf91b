"""Builds tpdm_b200/libtpdm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m tpdm_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtpdm_b200.so")
SOURCES = ["runtime.cu", "gemm_tcgen05.cu", "gemm2_tcgen05.cu", "attention_tcgen05.cu", "kernels.cu", "tpm_train.cu", "vae.cu", "api.cu"]
HEADERS = ["common.cuh", "gemm_epilogue.cuh", "host.h", "kernels.h", os.path.join("..", "..", "include", "tpdm_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    obj_dir = os.path.join(HERE, "_build")
    os.makedirs(obj_dir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC, *FLAGS, "-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc {src} failed ---\n{out}\n")
        elif verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        with open(os.path.join(obj_dir, src + ".log"), "w") as f:
            f.write(out)
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        subprocess.check_call([NVCC, "-shared", "-o", LIB, *objs, "-lcudart"])
    return LIB


def build_variant(name: str, defines, sources=None) -> str:
    """Diagnostic build: the whole library with extra -D flags into tpdm_b200/_build/libtpdm_<name>.so (load it by setting
    TPDM_B200_LIB).  Used for the instrumented attention kernel (tools/attn_trace.py) and kernel experiments."""
    obj_dir = os.path.join(HERE, "_build", name)
    os.makedirs(obj_dir, exist_ok=True)
    lib = os.path.join(HERE, "_build", f"libtpdm_{name}.so")
    procs, objs = [], []
    for src in SOURCES:
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(o)
        flags = [f"-D{d}" for d in defines] if (sources is None or src in sources) else []
        base = os.path.join(HERE, "_build", src.replace(".cu", ".o"))
        if not flags and os.path.exists(base):      # untouched translation unit: reuse the product object
            objs[-1] = base
            continue
        procs.append((src, subprocess.Popen([NVCC, *FLAGS, *flags, "-c", os.path.join(CSRC, src), "-o", o], stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(f"--- nvcc {src} failed ---\n{out}\n")
            raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-o", lib, *objs, "-lcudart"])
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        only = sys.argv[i + 3].split(",") if len(sys.argv) > i + 3 else ["attention_tcgen05.cu"]
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2].split(","), only))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
