"""Build libevx_b200.so in-tree with nvcc for sm_100a:  python -m evoxels_b200.build"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libevx_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "0",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=True):
    if up_to_date() and not force:
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "lib64")
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", OUT, *sources(), "-lcufft",
           "-Xlinker", f"-rpath={cuda_lib}"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
