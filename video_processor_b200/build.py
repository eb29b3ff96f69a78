"""Builds libvpb200.so (hand-written sm_100a CUDA kernels + the C ABI of include/vpb200.h) in-tree.

    python -m video_processor_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  The shared object is git-ignored but travels to the
GPU box with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvpb200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-Xptxas", "-v",
    "-shared",
]


def _sources():
    srcs = [os.path.join(CSRC, "vpb200.cu")]
    host_dir = os.path.join(CSRC, "host")
    if os.path.isdir(host_dir):
        srcs += sorted(os.path.join(host_dir, f) for f in os.listdir(host_dir) if f.endswith((".cpp", ".cu")))
    return srcs


def _deps():
    deps = []
    for d, _, files in os.walk(CSRC):
        deps += [os.path.join(d, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".cpp", ".hpp"))]
    inc = os.path.join(ROOT, "include")
    for d, _, files in os.walk(inc):
        deps += [os.path.join(d, f) for f in files]
    deps.append(os.path.abspath(__file__))
    return deps


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False, variant=None, defines=()):
    """variant/defines: a development build `libvpb200_<variant>.so` with extra -D flags, for A/B timing on the GPU box
    (`VPB200_LIB=<path> python bench.py ...`); the product library is the plain build."""
    lib = LIB if not variant else os.path.join(HERE, f"libvpb200_{variant}.so")
    if not variant and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", lib] + _sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log" if not variant else f"build_{variant}.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libvpb200.so")
    if verbose:
        print(log)
    return lib


if __name__ == "__main__":
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=var, defines=defs))
