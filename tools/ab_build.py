#!/usr/bin/env python3
"""Dev tool (build container): A/B builds of ONE translation unit.  For every `name:flags` argument compile csrc/<unit>.cu with
the extra flags, link it with the product build's other objects into libntm_b200_<name>.so and link a matching torch extension
(ntm_b200_torch_<name>.so).  Select at run time with NTM_B200_LIB / NTM_B200_TORCH_LIB (tools/ab_run.sh does).

    python tools/ab_build.py gru_mma "k5:-DNTM_MMA_KSPLIT=5" "bs:-DNTM_MMA_BLEND=1 -DNTM_MMA_SHARE=1"
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neural-tape-modeling_b200")
NVCC = "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def one(unit, spec):
    name, _, flags = spec.partition(":")
    bdir = os.path.join(PKG, f"build_{name}")
    os.makedirs(bdir, exist_ok=True)
    obj = os.path.join(bdir, unit + ".o")
    cmd = [NVCC, *ARCH, *flags.split(), "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v",
           "-c", os.path.join(PKG, "csrc", unit + ".cu"), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    open(obj + ".log", "w").write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError(r.stderr[-3000:])
    others = [os.path.join(PKG, "build", f) for f in sorted(os.listdir(os.path.join(PKG, "build")))
              if f.endswith(".o") and f != unit + ".o"]
    lib = os.path.join(PKG, f"libntm_b200_{name}.so")
    subprocess.run([NVCC, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", lib, *others, obj], check=True)
    env = dict(os.environ, NTM_B200_BUILD_SUFFIX=f"_{name}")
    subprocess.run([sys.executable, "-c",
                    "import importlib.util,sys; s=importlib.util.spec_from_file_location('b', sys.argv[1]); m=importlib.util.module_from_spec(s); "
                    "s.loader.exec_module(m); m.build_torch_extension(force=True)", os.path.join(PKG, "build.py")], env=env, check=True)
    return name


if __name__ == "__main__":
    unit, specs = sys.argv[1], sys.argv[2:]
    with ThreadPoolExecutor(max_workers=8) as ex:
        print("built:", list(ex.map(lambda s: one(unit, s), specs)))
