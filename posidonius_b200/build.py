"""Builds libposidonius_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

The step kernel is compiled once per geometry build and arithmetic mode (csrc/kernels_tu.cu with different -D flags);
the objects are built in parallel and linked with the C ABI (pb200_api.cu) and the host-side case I/O into one library.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libposidonius_b200.so")
CLI = os.path.join(HERE, "bin", "posidonius-b200")
HEADERS = ["host/json_min.hpp", "host/cli.cpp", "strict.cuh", "whfast_kernel.cuh", "cold_slots.cuh", "dyn_effects.cuh", "forces_fast.cuh",
           "gr_variants.cuh", "exact_effects.cuh", "strict_gr_variants.cuh", "whfast_step.cuh", "ensemble_host.hpp", "small_step.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--fmad=true", "-Xptxas", "-v",
]
# (object name, source, extra defines)
UNITS = [  # "k_s*": the lane = planet kernel of the 2- and 3-body configurations (kernels_small_tu.cu)
    ("api", "pb200_api.cu", []),
    ("case_io", "host/case_io.cpp", []),
    ("k_generic_fast", "kernels_tu.cu", ["PB_TU_GENERIC=0"]),
    ("k_generic_strict", "kernels_tu.cu", ["PB_TU_GENERIC=1"]),
    ("k_generic_hybrid", "kernels_tu.cu", ["PB_TU_GENERIC=2"]),
    ("k_n8", "kernels_tu.cu", ["PB_TU_FIXED=8"]),
    ("k_n8w", "kernels_tu.cu", ["PB_TU_FIXED=8", "PB_TU_WIDE=1"]),
    ("k_s2", "kernels_small_tu.cu", ["PB_TU_SMALL=2"]),
    ("k_s2t", "kernels_small_tu.cu", ["PB_TU_SMALL=20"]),
    ("k_s3", "kernels_small_tu.cu", ["PB_TU_SMALL=3"]),
    ("k_s3e", "kernels_small_tu.cu", ["PB_TU_SMALL=30"]),
    ("k_s3j", "kernels_small_tu.cu", ["PB_TU_SMALL=31"]),
    ("k_s3p", "kernels_small_tu.cu", ["PB_TU_SMALL=32"]),
    ("k_s2any", "kernels_small_tu.cu", ["PB_TU_SMALL=29"]),
    ("k_s3any", "kernels_small_tu.cu", ["PB_TU_SMALL=39"]),
    ("k_s3jany", "kernels_small_tu.cu", ["PB_TU_SMALL=38"]),
]


def _deps():
    deps = [os.path.join(CSRC, f) for f in HEADERS if os.path.exists(os.path.join(CSRC, f))]
    deps += [os.path.join(CSRC, src) for _, src, _ in UNITS]
    deps += [os.path.join(HERE, "..", "include", "posidonius_b200.h"), __file__]
    return deps


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(CLI):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(nvcc, name, src, defines, objdir, log_lines):
    obj = os.path.join(objdir, name + ".o")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-c", "-o", obj, os.path.join(CSRC, src)]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log_lines.append("### " + name + "\n" + " ".join(cmd) + "\n" + proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s" % (name, proc.stdout[-4000:]))
    return obj


def build(force=False, verbose=False, out=None, defines=(), only=None):
    """out/defines: experimental variants (e.g. -DPB_MIN_BLOCKS=3) built next to the product library.
    only: iterable of unit names to recompile (the other objects are reused) — development shortcut."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    target = out or LIB
    objdir = OBJ if out is None else OBJ + "_" + os.path.basename(out)
    os.makedirs(objdir, exist_ok=True)
    log_lines = []
    # a variant build (out != None) with `only`: the other units are the product objects
    reuse = out is not None and only is not None
    units = [u for u in UNITS if only is None or u[0] in only or (not reuse and not os.path.exists(os.path.join(objdir, u[0] + ".o")))]
    workers = int(os.environ.get("PB200_BUILD_JOBS", str(os.cpu_count() or 4)))
    with ThreadPoolExecutor(max_workers=workers) as pool:
        futures = [pool.submit(_compile, nvcc, name, src, list(defs) + list(defines), objdir, log_lines) for name, src, defs in units]
        errors = []
        for f in futures:
            try:
                f.result()
            except Exception as exc:  # collect every failing unit before raising
                errors.append(str(exc))
    log = os.path.join(HERE, "build.log" if out is None else os.path.basename(out) + ".log")
    objs = [os.path.join(OBJ if reuse and name not in only else objdir, name + ".o") for name, _, _ in UNITS]
    if not errors:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", target] + objs
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log_lines.append("### link\n" + " ".join(cmd) + "\n" + proc.stdout)
        if proc.returncode != 0:
            errors.append("link failed:\n" + proc.stdout[-4000:])
    with open(log, "w") as f:
        f.write("\n".join(log_lines))
    if verbose or errors:
        sys.stderr.write("\n".join(log_lines) if verbose else "\n".join(errors))
    if errors:
        raise RuntimeError("building %s failed (see %s)" % (target, log))
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
    only = None
    for a in sys.argv[1:]:
        if a.startswith("--only="):
            only = set(a[len("--only="):].split(","))
    build(force=True, verbose="-v" in sys.argv, only=only)
    print(LIB)
