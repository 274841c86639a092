"""In-tree build of the CUDA extension (nvcc cross-compiles sm_100a without a GPU)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRCS = [os.path.join(PKG, "csrc", f) for f in ("rb_kernels.cu", "rb_head.cu", "rb_head_tc.cu", "rb_peer.cu")]
DEPS = SRCS + [os.path.join(PKG, "csrc", "rb_internal.cuh")]
HDR = os.path.join(ROOT, "include", "rainbow_b200.h")
SO = os.path.join(PKG, "librainbow_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include")]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in DEPS + [HDR])


def build(force=False, verbose=False):
    """Compile rainbow_b200/csrc/*.cu -> rainbow_b200/librainbow_b200.so for sm_100a."""
    if not force and not stale():
        return SO
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build librainbow_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO + ".tmp"] + SRCS
    subprocess.check_call(cmd)
    os.replace(SO + ".tmp", SO)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
