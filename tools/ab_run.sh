#!/bin/bash
# A/B builds of the library (NTM_B200_LIB), then GPU tests
P=neural-tape-modeling_b200
for l in $(cd $P; ls libntm_b200*.so); do
  NTM_B200_LIB=$PWD/$P/$l python tools/ab_libs.py f16 2>&1 | tail -12
done > gpurun_out/ab_order.txt 2>&1
cat gpurun_out/ab_order.txt
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k best 2>&1 | grep -E "^E|assert|passed|failed" | head -30
python -m pytest tests/test_tc_gpu.py tests/test_rt_gpu.py -m gpu -x -q 2>&1 | tail -5
