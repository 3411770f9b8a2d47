"""Build libdril_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdril_b200.so")
SOURCES = ["api.cu"]
HEADERS = ["common.cuh", "env.cuh", "mlp.cuh", "mma_tiles.cuh", "rollout.cuh", "rollout_tc.cuh", "gae.cuh", "update.cuh", "update_tc.cuh", "update_ft.cuh",
           os.path.join("..", "..", "include", "dril_b200.h")]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "--cudart", "static", "-Xptxas", "-v" if verbose else "-O3",
           "-o", LIB] + os.environ.get("DRIL_NVCC_EXTRA", "").split() + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libdril_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
