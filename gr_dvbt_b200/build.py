"""Builds gr_dvbt_b200/libdvbt_b200.so in-tree with nvcc for sm_100a (no JIT cache: the
built .so travels with the repo snapshot to the GPU box)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdvbt_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tmp = LIB + ".building"   # the finished library replaces the old one atomically (a GPU-box snapshot never sees half a file)
    extra = ["-DDVBT_B200_LEGACY_ACS"] if os.environ.get("DVBT_B200_BUILD_LEGACY_ACS") else []   # byte-SWAR + two-lane ACS kernels (A/B baselines)
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", tmp] + sources() + ["-lcufft"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libdvbt_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    build_native(force=True, verbose="-v" in sys.argv)
    print(LIB)
