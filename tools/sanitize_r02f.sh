#!/bin/bash
# compute-sanitizer passes over the lean 4-stream mma.sync kernel (csrc/gru_mma4.cu, automatic dispatch at these widths): GRU in f16 /
# bf16 / strict f16x3 with ragged widths and lengths that are no multiple of the 4-step unroll or the 128-step chunk, DiffDelGRU with the
# on-chip pre_d ring (cfg 3's D = 365) and racecheck on both; the zero-padded hidden sizes through the Python classes.
run() { echo "== $*"; timeout 280 compute-sanitizer --tool "$1" --error-exitcode 9 "${@:2}" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|checksum|passed|failed|Error" | head -6; }
run memcheck python tools/run_once.py 5 701 f16
run memcheck python tools/run_once.py 7 389 bf16
run memcheck python tools/run_once.py 6 643 f16x3
run memcheck python tools/run_once.py 5 701 f16 0 0 diffdel
run memcheck python tools/run_once.py 3 515 f16x3 0 0 diffdel
run racecheck python tools/run_once.py 5 300 f16
run racecheck python tools/run_once.py 5 300 f16 0 0 diffdel
run memcheck python -m pytest tests/test_hidden_sizes.py -m gpu -x -q -k "carried or default or (predict and fp32 and (0 or 5))"
