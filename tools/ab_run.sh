#!/bin/bash
# Dev tool (GPU box): A/B every build of the library found next to the package (libntm_b200*.so + the matching torch
# extension, selected through NTM_B200_LIB / NTM_B200_TORCH_LIB; built by tools/ab_build.py) with tools/ab_libs.py.
# usage: tools/ab_run.sh [out-tag] [mode]
P=$PWD/neural-tape-modeling_b200
TAG=${1:-ab}
MODE=${2:-f16}
for l in $(cd $P; ls libntm_b200*.so); do
  sfx=${l#libntm_b200}; sfx=${sfx%.so}
  NTM_B200_LIB=$P/$l NTM_B200_TORCH_LIB=$P/ntm_b200_torch$sfx.so python tools/ab_libs.py $MODE 2>&1 | tail -14
done > gpurun_out/${TAG}.txt 2>&1
cat gpurun_out/${TAG}.txt
