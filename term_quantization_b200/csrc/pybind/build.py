"""Ahead-of-time build of the pybind11 adapter (tr_cuda_pybind.cpp) against the installed torch.

    python term_quantization_b200/csrc/pybind/build.py

Produces term_quantization_b200/tr_cuda_pybind.so next to libtq_b200.so (found through an $ORIGIN rpath).  The
compile uses torch's own include / library paths and ABI flags; there is no GPU code in the adapter (the kernels
live in libtq_b200.so, sm_100a only), so no nvcc and no arch list are involved."""
import os
import subprocess
import sys
import sysconfig


def main():
    import torch
    from torch.utils import cpp_extension as ce
    here = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.dirname(os.path.dirname(here))
    root = os.path.dirname(pkg)
    out = os.path.join(pkg, "tr_cuda_pybind.so")
    src = os.path.join(here, "tr_cuda_pybind.cpp")
    lib = os.path.join(pkg, "libtq_b200.so")
    newest = max(os.path.getmtime(p) for p in (src, lib, os.path.join(root, "include", "tq_b200.h")))
    if os.path.exists(out) and os.path.getmtime(out) >= newest:
        return out
    cuda_home = ce.CUDA_HOME or "/usr/local/cuda"
    inc = ce.include_paths() + [os.path.join(cuda_home, "include"), sysconfig.get_paths()["include"],
                                os.path.join(root, "include")]
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=tr_cuda_pybind",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    cmd += [f"-I{p}" for p in inc] + [src, "-o", out]
    cmd += [f"-L{p}" for p in ce.library_paths()] + [f"-L{pkg}", "-l:libtq_b200.so", "-lc10", "-lc10_cuda", "-ltorch_cpu",
                                                     "-ltorch", "-ltorch_python", "-Wl,-rpath,$ORIGIN"]
    cmd += [f"-Wl,-rpath,{p}" for p in ce.library_paths()]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(main())
    sys.exit(0)
