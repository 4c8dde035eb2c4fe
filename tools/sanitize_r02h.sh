#!/bin/bash
# compute-sanitizer memcheck over the kernels changed last in round 2: the stand-alone delay line (tap reuse, 32 groups per thread with
# prefetched delays, peeled row heads: aligned and misaligned rows, long and tiny rows), the swizzled DCPreESR ring, the zero-padded hidden
# sizes through the lean kernel.
run() { echo "== $*"; timeout 280 compute-sanitizer --tool "$1" --error-exitcode 9 "${@:2}" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|checksum|passed|failed|Error" | head -6; }
run memcheck python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "delay_line"
run memcheck python tools/delay_once.py 37 300001 365
run memcheck python -m pytest tests/test_loss_gpu.py -m gpu -x -q
run racecheck python tools/esr_once.py 8 200000
