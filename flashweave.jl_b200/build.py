"""Builds libfwgpu.so in-tree with nvcc for sm_100a (the only target)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libfwgpu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=default",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    hdr = os.path.join(os.path.dirname(HERE), "include", "fwgpu.h")
    return any(os.path.getmtime(s) > t for s in sources() + [hdr, os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("FW_NVCC_EXTRA", "").split()          # experiments only (e.g. -DFW_HITON_MINB=6)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", os.environ.get("FW_SO_OUT", SO), os.path.join(CSRC, "fwgpu.cu")]
    env = dict(os.environ)
    # the image exports CXX/CC wrappers that nvcc cannot drive; use the system host compiler
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libfwgpu.so")
    if verbose:
        print(r.stdout)
    return SO


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(SO)
