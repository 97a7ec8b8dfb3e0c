"""Builds libgd_b200.so (all CUDA kernels + the C-ABI) for sm_100a with nvcc; no GPU needed.
Translation units are compiled in parallel (one nvcc -c per .cu) and linked into the shared library."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["api.cu", "gemm_tcgen05.cu", "gemm_resid_ln.cu", "gemm_ln_a.cu", "elementwise.cu", "attention.cu", "attention_tc.cu", "speech_kernels.cu"]
OUT = os.path.join(os.path.dirname(HERE), "libgd_b200.so")
OBJ_DIR = os.path.join(HERE, "_obj")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--compiler-options", "-fPIC"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "..", "include", "gd_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_trace():
    """Profiling build with in-kernel %globaltimer stamps (-DGD_TRACE) -> libgd_b200_trace.so next to the product library;
    loaded through GD_LIB by profiles/kernel_timeline.py only."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out = os.path.join(os.path.dirname(HERE), "libgd_b200_trace.so")
    obj_dir = os.path.join(HERE, "_obj_trace")
    os.makedirs(obj_dir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(obj_dir, src[:-3] + ".o")
        subprocess.run([nvcc] + FLAGS + ["-DGD_TRACE", "-c", os.path.join(HERE, src), "-o", obj], check=True, cwd=HERE)
        return obj
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["-lcudart"], check=True, cwd=HERE)
    return out


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        subprocess.run([nvcc] + FLAGS + extra + ["-c", os.path.join(HERE, src), "-o", obj], check=True, cwd=HERE)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs + ["-lcudart"], check=True, cwd=HERE)
    return OUT


if __name__ == "__main__":
    if "--trace" in sys.argv:
        print(build_trace())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
