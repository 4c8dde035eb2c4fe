#!/bin/bash
# one GPU-box pass: tests, smoke, parity report, bench (both arms), launch list and an ncu capture of the dominant kernels
TAG=${1:-r02}
rm -f gpurun_out/parity_10s.json
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_gpu.txt; tail -5 gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python tools/parity_report.py > gpurun_out/${TAG}_parity.json 2> gpurun_out/${TAG}_parity.err; tail -c 400 gpurun_out/${TAG}_parity.json
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_reference.json gpurun_out/${TAG}_bench.json
