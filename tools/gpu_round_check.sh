#!/bin/bash
# one GPU-box pass: tests, smoke, bench (both arms), launch list and an ncu capture of the dominant kernel
TAG=${1:-r01}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.txt; tail -3 gpurun_out/${TAG}_pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --seconds 2 --no-aux --no-cpu > gpurun_out/${TAG}_bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru_mma -s 1 -c 1 -f -o gpurun_out/${TAG}_mma_b1024 python tools/run_once.py 1024 96000 f16 > gpurun_out/${TAG}_ncu_full.log 2>&1
cat gpurun_out/${TAG}_bench.json
