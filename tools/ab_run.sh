#!/bin/bash
# A/B builds of the library (NTM_B200_LIB) on the mma.sync kernel, then the GPU tests that exercise it
P=neural-tape-modeling_b200
for l in $(cd $P; ls libntm_b200*.so); do
  NTM_B200_LIB=$PWD/$P/$l python tools/ab_libs.py f16 2>&1 | tail -12
done > gpurun_out/ab_order.txt 2>&1
cat gpurun_out/ab_order.txt
python tools/cfg3_time.py 2>&1 | tail -3
python -m pytest tests/test_tc_gpu.py tests/test_rt_gpu.py tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -3
