#!/bin/bash
# compute-sanitizer passes over round 2's new / changed kernels (small shapes): strict forms of the mma.sync kernel (pair-column
# 4-stream layout, 8-stream layout), the late-blend 4-stream form, the stream-major tcgen05 kernel in every operand format (f16 with
# the four-unit n reciprocal, strict f16x3, tf32), the binary16 transport's conversion passes, racecheck on the strict 4-stream form.
run() { echo "== $*"; timeout 280 compute-sanitizer --tool "$1" --error-exitcode 9 "${@:2}" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|checksum|passed|failed|Error" | head -6; }
run memcheck python tools/run_once.py 5 700 f16x3 4 3
run memcheck python tools/run_once.py 9 700 f16x3 8 3
run memcheck python tools/run_once.py 7 700 f16 4 3
run racecheck python tools/run_once.py 5 300 f16x3 4 3
run memcheck python tools/run_once.py 300 200 f16 6 4
run memcheck python tools/run_once.py 300 200 f16x3 5 4
run memcheck python tools/run_once.py 300 200 tf32 5 4
run memcheck python -m pytest tests/test_host_f16_gpu.py -m gpu -x -q -k "33-9000"
