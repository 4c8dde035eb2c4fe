#!/bin/bash
# Dev tool (GPU box): A/B every build of the library found next to the package (libntm_b200*.so, selected through
# NTM_B200_LIB) on the mma.sync kernel with tools/ab_libs.py, then run the GPU tests that exercise that kernel.
# Variants are built by hand, e.g.:
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -DSOME_EXPERIMENT \
#        -c csrc/gru_mma.cu -o build/x/gru_mma.o && nvcc ... -shared -o libntm_b200_x.so <other objects> build/x/gru_mma.o
P=neural-tape-modeling_b200
for l in $(cd $P; ls libntm_b200*.so); do
  NTM_B200_LIB=$PWD/$P/$l python tools/ab_libs.py f16 2>&1 | tail -12
done > gpurun_out/ab_order.txt 2>&1
cat gpurun_out/ab_order.txt
python -m pytest tests/test_tc_gpu.py tests/test_rt_gpu.py -m gpu -x -q 2>&1 | tail -3
