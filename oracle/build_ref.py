#!/usr/bin/env python
"""Build the UNMODIFIED reference `pychain_C` extension into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package (`pychain_b200`); only tests/, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of bench.py may use it, and only as
the checker / the timed CPU baseline.

The sources are compiled *where they lie* under /root/reference (never copied
into this repo) with direct g++/nvcc commands -- the reference's own
setup.py / Makefile are not run.  Files compiled (all of
/root/reference/pytorch_binding/src):

    pychain.cc  base.cc  chain-computation.cc  chain-log-domain-computation.cc
    chain-kernels.cu  chain-log-domain-kernels.cu

The only external dependency is the PyTorch C++ API that ships inside this
image's torch wheel.  Output: oracle/_ref/pychain_C.so (git-ignored, NOT
gpurun-ignored: the prebuilt .so travels to the GPU box, where
/root/reference does not exist).

Flags mirror what the reference's stock setuptools build produces: -O2/-O3,
-DNDEBUG (asserts compiled out, see SURVEY.md Appx C-3), CUDA code for
sm_100 (plain `compute_100`, the reference has no arch-specific code).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("PYCHAIN_REFERENCE", "/root/reference") + "/pytorch_binding/src"
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "pychain_C.so")

CC_FILES = ["pychain.cc", "base.cc", "chain-computation.cc", "chain-log-domain-computation.cc"]
CU_FILES = ["chain-kernels.cu", "chain-log-domain-kernels.cu"]


def have_reference():
    return os.path.isdir(REF_SRC)


def is_built():
    return os.path.isfile(OUT_SO)


def build(force=False, verbose=False):
    """Compile the reference extension.  Returns path of the .so, or None when
    the reference tree is absent (GPU box) and no prebuilt .so exists."""
    stage_reference_loss()
    stage_reference_graph()
    if is_built() and not force:
        if not have_reference():
            return OUT_SO
        newest = max(os.path.getmtime(os.path.join(REF_SRC, f)) for f in CC_FILES + CU_FILES)
        if os.path.getmtime(OUT_SO) >= newest:
            return OUT_SO
    if not have_reference():
        return OUT_SO if is_built() else None

    import torch  # noqa: F401  (only for its include / lib dirs)
    from torch.utils.cpp_extension import include_paths, library_paths

    os.makedirs(os.path.join(OUT_DIR, "obj"), exist_ok=True)
    incs = [f"-I{p}" for p in include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}", f"-I{REF_SRC}"]
    defs = ["-DTORCH_EXTENSION_NAME=pychain_C", "-DTORCH_API_INCLUDE_EXTENSION_H", "-DNDEBUG"]
    cxx = ["g++", "-std=c++17", "-O2", "-fPIC", "-w"] + defs + incs
    nvcc = ["nvcc", "-std=c++17", "-O3", "-Xcompiler", "-fPIC", "-w",
            "-gencode", "arch=compute_100,code=sm_100",
            "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__"] + defs + incs

    procs, objs = [], []
    for f in CC_FILES:
        o = os.path.join(OUT_DIR, "obj", f + ".o")
        objs.append(o)
        procs.append(subprocess.Popen(cxx + ["-c", os.path.join(REF_SRC, f), "-o", o]))
    for f in CU_FILES:
        o = os.path.join(OUT_DIR, "obj", f + ".o")
        objs.append(o)
        procs.append(subprocess.Popen(nvcc + ["-c", os.path.join(REF_SRC, f), "-o", o]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference oracle: compile failed")

    libdirs = library_paths("cuda")
    link = ["g++", "-shared", "-o", OUT_SO] + objs
    for d in libdirs:
        link += [f"-L{d}", f"-Wl,-rpath,{d}"]
    link += ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    if verbose:
        print(" ".join(link))
    subprocess.check_call(link)
    return OUT_SO


REF_LOSS = os.environ.get("PYCHAIN_REFERENCE", "/root/reference") + "/pychain/loss.py"
STAGED_LOSS = os.path.join(OUT_DIR, "reference_pychain_loss.py")


REF_GRAPH = os.environ.get("PYCHAIN_REFERENCE", "/root/reference") + "/pychain/graph.py"
STAGED_GRAPH = os.path.join(OUT_DIR, "reference_pychain_graph.py")


def stage_reference_graph():
    """Same as stage_reference_loss for the reference's pychain/graph.py: tests/test_fst_golden.py runs it, as is,
    over the reference-compiled `simplefst` and compares the containers it builds with pychain_b200.graph's."""
    import shutil
    if os.path.isfile(REF_GRAPH):
        os.makedirs(OUT_DIR, exist_ok=True)
        shutil.copyfile(REF_GRAPH, STAGED_GRAPH)
    return STAGED_GRAPH if os.path.isfile(STAGED_GRAPH) else None


def stage_reference_loss():
    """Stage the reference's UNMODIFIED pychain/loss.py next to the reference .so (oracle/_ref/ is git-ignored and
    travels to the GPU box, where /root/reference does not exist).  tests/test_gpu_dropin.py runs that file, as is,
    on top of the drop-in modules.  Never committed; returns the staged path or None."""
    import shutil
    if os.path.isfile(REF_LOSS):
        os.makedirs(OUT_DIR, exist_ok=True)
        shutil.copyfile(REF_LOSS, STAGED_LOSS)
    return STAGED_LOSS if os.path.isfile(STAGED_LOSS) else None


FST_SRC = os.environ.get("PYCHAIN_REFERENCE", "/root/reference") + "/openfst_binding/src/fstext.cc"
FST_SO = os.path.join(OUT_DIR, "simplefst.so")
FST_SHIM = os.path.join(HERE, "fst_shim")


def build_simplefst(force=False):
    """Compile the reference's UNMODIFIED openfst_binding/src/fstext.cc (FstToTensor :19-117, SetLeakyProbs
    :120-171, the `simplefst` module :174-184) into oracle/_ref/simplefst.so.  OpenFst is not in this image, so
    the include path supplies oracle/fst_shim/fst/fstlib.h, an independent stand-in for the container classes the
    file touches; the algorithms that run are the reference's.  Returns the .so path or None."""
    if os.path.isfile(FST_SO) and not force:
        if not os.path.isfile(FST_SRC):
            return FST_SO
        newest = max(os.path.getmtime(FST_SRC), os.path.getmtime(os.path.join(FST_SHIM, "fst", "fstlib.h")))
        if os.path.getmtime(FST_SO) >= newest:
            return FST_SO
    if not os.path.isfile(FST_SRC):
        return FST_SO if os.path.isfile(FST_SO) else None
    import torch  # noqa: F401
    from torch.utils.cpp_extension import include_paths, library_paths
    os.makedirs(OUT_DIR, exist_ok=True)
    incs = [f"-I{p}" for p in include_paths()] + [f"-I{sysconfig.get_paths()['include']}", f"-I{FST_SHIM}"]
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-w", "-shared", "-DTORCH_EXTENSION_NAME=simplefst",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-DNDEBUG"] + incs + [FST_SRC, "-o", FST_SO]
    for d in library_paths():
        cmd += [f"-L{d}", f"-Wl,-rpath,{d}"]
    cmd += ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"]
    subprocess.check_call(cmd)
    return FST_SO


def load_simplefst():
    """Import oracle/_ref/simplefst.so under a private module name (the product probes for a real `simplefst`)."""
    import importlib.util
    import torch  # noqa: F401
    if not os.path.isfile(FST_SO):
        raise FileNotFoundError(FST_SO + " not built (run oracle/build_ref.py where /root/reference exists)")
    spec = importlib.util.spec_from_file_location("simplefst", FST_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Import oracle/_ref/pychain_C.so as a module object WITHOUT registering
    it in sys.modules under the name `pychain_C` (the product provides its own
    module of that name)."""
    import importlib.util
    import torch  # noqa: F401  must be imported first (libtorch symbols)
    if not is_built():
        raise FileNotFoundError(OUT_SO + " not built (run oracle/build_ref.py where /root/reference exists)")
    spec = importlib.util.spec_from_file_location("pychain_C", OUT_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    so = build(force="--force" in sys.argv, verbose=True)
    print("reference oracle:", so if so else "unavailable (no /root/reference, no prebuilt .so)")
    print("reference simplefst (over oracle/fst_shim):", build_simplefst(force="--force" in sys.argv))
