"""Compile integration/pychain_binding.cc (the C++ binding of INTEGRATION.md section 3) against
libchain_b200.so into integration/_build/pychain_C_b200.so (git-ignored; travels to the GPU box).

    python integration/build_binding.py

Direct g++ command (torch headers from the installed wheel); loaded by tests/test_gpu_dropin.py."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build", "pychain_C_b200.so")
SRC = os.path.join(HERE, "pychain_binding.cc")
LIBDIR = os.path.join(ROOT, "pychain_b200", "lib")


def build(force=False):
    lib = os.path.join(LIBDIR, "libchain_b200.so")
    if (not force and os.path.isfile(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC)
            and os.path.getmtime(OUT) >= os.path.getmtime(os.path.join(ROOT, "include", "chain_b200.h"))):
        return OUT
    from torch.utils.cpp_extension import include_paths, library_paths
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-DTORCH_EXTENSION_NAME=pychain_C_b200",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-I" + os.path.join(ROOT, "include"),
           "-I" + sysconfig.get_paths()["include"]]
    cmd += ["-I" + p for p in include_paths("cuda")]
    cmd += [SRC, "-o", OUT, "-L" + LIBDIR, "-lchain_b200", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,$ORIGIN/../../pychain_b200/lib"]
    for d in library_paths("cuda"):
        cmd += ["-L" + d, "-Wl,-rpath," + d]
    cmd += ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    subprocess.check_call(cmd)
    return OUT


def load():
    import importlib.util
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location("pychain_C_b200", build())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
