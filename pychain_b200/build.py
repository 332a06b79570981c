"""Build libchain_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m pychain_b200.build [--force] [--verbose]

The .so lands in pychain_b200/lib/ (git-ignored, shipped to the GPU box by
gpurun).  No torch headers, no pybind: the library's interface is
include/chain_b200.h and nothing else.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib", "libchain_b200.so")
SOURCES = ["generic.cu", "den_pack.cu", "den_shared.cu", "den_shared_log.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _sources():
    return [os.path.join(CSRC, f) for f in SOURCES if os.path.isfile(os.path.join(CSRC, f))]


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = _sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "chain_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines):
    """An experimental variant of the library (tuning macros), e.g. build_variant("nf4", ["-DCHAIN_B200_NF2_G4=4"])
    -> pychain_b200/lib/libchain_b200_nf4.so; select it at run time with CHAIN_B200_LIB=<path>."""
    out = os.path.join(PKG, "lib", "libchain_b200_%s.so" % name)
    cmd = [os.environ.get("NVCC", "nvcc")] + NVCC_FLAGS + list(defines) + [
        "-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", out] + _sources()
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out


def build(force=False, verbose=False):
    """One `nvcc -c` per source (in parallel), then one link: libchain_b200.so."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    libdir = os.path.dirname(LIB)
    objdir = os.path.join(libdir, "obj")
    os.makedirs(objdir, exist_ok=True)
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    jobs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        cmd = [nvcc] + NVCC_FLAGS + inc + ["-c", src, "-o", obj]
        jobs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log, failed = "", False
    for cmd, obj, p in jobs:
        out = p.communicate()[0]
        log += " ".join(cmd) + "\n" + out
        failed |= p.returncode != 0
    if not failed:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + [j[1] for j in jobs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + r.stdout + r.stderr
        failed = r.returncode != 0
    with open(os.path.join(libdir, "build.log"), "w") as f:
        f.write(log)
    if verbose or failed:
        print(log)
    if failed:
        raise RuntimeError("nvcc failed building libchain_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv))
