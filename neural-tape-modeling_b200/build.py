"""In-tree build of libntm_b200.so (nvcc cross-compiles sm_100a without a GPU).

    python neural-tape-modeling_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
# NTM_B200_BUILD_SUFFIX: developer A/B builds (e.g. "_x" with NTM_EXTRA_NVCC_FLAGS=-DSOME_EXPERIMENT) next to the product
# library; selected at run time with NTM_B200_LIB (lib.py)
SUFFIX = os.environ.get("NTM_B200_BUILD_SUFFIX", "")
LIB = os.path.join(PKG, f"libntm_b200{SUFFIX}.so")
OBJ = os.path.join(PKG, f"build{SUFFIX}")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
EXTRA = os.environ.get("NTM_EXTRA_NVCC_FLAGS", "").split()
FLAGS = [*EXTRA, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(os.path.dirname(PKG), "include", "ntm_b200.h"))
    return max(os.path.getmtime(p) for p in paths)


def build_library(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link them into libntm_b200.so next to this file."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [NVCC, *ARCH, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [NVCC, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


TORCH_LIB = os.path.join(PKG, f"ntm_b200_torch{SUFFIX}.so")


def build_torch_extension(force=False):
    """Compile csrc/torch_binding.cpp -- the PyTorch C++ extension that registers torch.ops.ntm.* over the C ABI -- against
    the running torch (include / library paths from torch.utils.cpp_extension) and link it to the in-tree libntm_b200.so
    (found at run time through an $ORIGIN rpath).  In-tree, so that the .so travels with the source tree."""
    src = os.path.join(CSRC, "torch_binding.cpp")
    deps = max(os.path.getmtime(src), os.path.getmtime(os.path.join(os.path.dirname(PKG), "include", "ntm_b200.h")),
               os.path.getmtime(LIB))
    if not force and os.path.exists(TORCH_LIB) and os.path.getmtime(TORCH_LIB) >= deps:
        return TORCH_LIB
    import logging
    import torch
    logging.getLogger("torch.utils.cpp_extension").setLevel(logging.ERROR)
    from torch.utils import cpp_extension as ce
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    # the system g++ (the C++ runtime torch itself was built against); NOT $CXX: this image points it at another toolchain
    # whose libstdc++ the extension then mixes with torch's -- the first exception thrown across the boundary crashed
    cmd = [os.environ.get("NTM_CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", src, "-o", TORCH_LIB,
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-DTORCH_EXTENSION_NAME=ntm_b200_torch"]
    cmd += [f"-I{p}" for p in ce.include_paths()] + [f"-I{cuda_home}/include"]
    cmd += [f"-L{p}" for p in ce.library_paths()] + ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch"]
    cmd += [f"-L{PKG}", f"-l:{os.path.basename(LIB)}", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"torch extension build failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return TORCH_LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_torch_extension(force="--force" in sys.argv))
