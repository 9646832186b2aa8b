"""Builds libposidonius_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libposidonius_b200.so")
SOURCES = ["pb200_api.cu", "host/case_io.cpp"]
CLI = os.path.join(HERE, "bin", "posidonius-b200")
HEADERS = ["host/json_min.hpp", "host/cli.cpp", "strict.cuh", "whfast_kernel.cuh", "dyn_effects.cuh", "forces_fast.cuh", "gr_variants.cuh", "strict_effects.cuh", "strict_gr_variants.cuh", "whfast_step.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--fmad=true", "-Xptxas", "-v",
]


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(CLI):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "posidonius_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    """out/defines: experimental variants (e.g. -DPB_MIN_BLOCKS=3) built next to the product library."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    target = out or LIB
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", target] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log" if out is None else os.path.basename(out) + ".log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building %s (see %s)" % (target, log))
    if out is None:
        build_cli()
    return target


def build_cli():
    """The `posidonius-b200 start|resume|ensemble` command line (host C++ over the C ABI)."""
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O2", "-std=c++17", "-o", CLI, os.path.join(CSRC, "host", "cli.cpp"), "-L" + HERE, "-lposidonius_b200",
           "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath," + HERE]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("building the CLI failed")
    return CLI


if __name__ == "__main__":
    build(force=True, verbose="-q" not in sys.argv)
    print(LIB)
