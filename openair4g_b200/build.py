"""Builds the in-tree CUDA library openair4g_b200/lib/liboai_turbo_b200.so for sm_100a.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU
box with the repo snapshot.
"""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.environ.get("OAI_TURBO_LIB") or os.path.join(LIBDIR, "liboai_turbo_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = glob.glob(os.path.join(CSRC, "*")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: build an experimental variant (tools/ use it for tuning sweeps)."""
    target = out or LIB
    if not force and out is None and not is_stale():
        return LIB
    os.makedirs(os.path.dirname(target), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] + sources()
    subprocess.run(cmd, check=True, cwd=CSRC)
    return target


if __name__ == "__main__":
    print(build(force=True, verbose=True))
