#!/bin/bash
# runs tools/tc_probe.bin over the layout hypotheses; one process per configuration
P=tools/tc_probe.bin
for fmt in tf32 bf16; do
  for cfg in "0 16" "0 192" "1 192" "1 16" "0 64"; do
    for swap in 0 1; do
      timeout 20 $P $fmt $cfg $swap 1 1 || echo "  -> rc=$? ($fmt $cfg swap=$swap)"
    done
  done
done
echo "---- chain latency (iters=200)"
for fmt in tf32 bf16; do
  for cfg in "0 16 0 200 1" "0 16 0 200 3" "0 16 0 200 6" "1 192 0 200 1" "1 192 0 200 2" "0 192 0 200 1" "0 64 0 200 1" "1 64 0 200 1" "0 256 0 200 1"; do
    timeout 20 $P $fmt $cfg || echo "  -> rc=$?"
  done
done
echo "---- N=8 at M=128 (may be illegal)"
timeout 20 $P tf32 0 8 0 1 1 || echo "  -> rc=$?"
timeout 20 $P tf32 0 24 0 1 1 || echo "  -> rc=$?"
