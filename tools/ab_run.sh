#!/bin/bash
# A/B builds of the library (NTM_B200_LIB), then the tensor-core GPU tests
P=neural-tape-modeling_b200
for l in libntm_b200.so libntm_b200_o1.so; do
  NTM_B200_LIB=$PWD/$P/$l python tools/ab_libs.py f16 2>&1 | tail -12
done > gpurun_out/ab_order.txt 2>&1
cat gpurun_out/ab_order.txt
python -m pytest tests/test_tc_gpu.py -m gpu -x -q 2>&1 | tail -5
