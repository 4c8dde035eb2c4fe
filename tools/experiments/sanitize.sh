#!/bin/bash
# compute-sanitizer passes over the round's new / changed kernels (small shapes; memcheck + racecheck for the shared-memory ring)
for args in "5 700 f16 4 3" "9 700 f16 8 3" "1200 300 f16 0 0"; do
  echo "== memcheck run_once $args"
  timeout 250 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/run_once.py $args 2>&1 | grep -E "ERROR SUMMARY|Invalid|checksum|Error" | head -5
done
echo "== memcheck esr_once 9 70000"
timeout 250 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/esr_once.py 9 70000 2>&1 | grep -E "ERROR SUMMARY|Invalid|^[0-9]" | head -5
echo "== racecheck esr_once 3 40001"
timeout 250 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/esr_once.py 3 40001 2>&1 | grep -E "RACECHECK SUMMARY|hazard|^[0-9]" | head -5
echo "== memcheck delay_once 7 50003 365 (unaligned rows) / 8 50000 6000"
timeout 250 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/delay_once.py 7 50003 365 2>&1 | grep -E "ERROR SUMMARY|Invalid|^-?[0-9]" | head -5
timeout 250 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/delay_once.py 8 50000 6000 2>&1 | grep -E "ERROR SUMMARY|Invalid|^-?[0-9]" | head -5
echo "== memcheck cfg3-style DiffDelGRU (deferred-head form + on-chip ring)"
timeout 250 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_tc_gpu.py -m gpu -x -q -k "delay_read_paths and 364" 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed" | head -5
