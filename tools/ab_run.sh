#!/bin/bash
# A/B builds of the library (NTM_B200_LIB) on the mma.sync kernel
P=neural-tape-modeling_b200
for l in $(cd $P; ls libntm_b200*.so); do
  NTM_B200_LIB=$PWD/$P/$l python tools/ab_libs.py f16 2>&1 | grep "B= 1024" | grep "(4, 3)"
done > gpurun_out/ab_order.txt 2>&1
cat gpurun_out/ab_order.txt
