#!/bin/bash
# one GPU-box pass: tests, smoke, parity report, bench (both arms), launch list of the bench command and ncu --set full captures of
# the dominant kernels (the headline mma.sync kernel at cfg 2's width, the strict form, the tcgen05 kernel)
TAG=${1:-r02}
rm -f gpurun_out/parity_10s.json
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_gpu.txt; tail -5 gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python tools/parity_report.py > gpurun_out/${TAG}_parity.json 2> gpurun_out/${TAG}_parity.err; tail -c 400 gpurun_out/${TAG}_parity.json
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_reference.json gpurun_out/${TAG}_bench.json
if [ "$2" == "profile" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --seconds 2 --no-aux --no-cpu --no-cfg4 > gpurun_out/${TAG}_bench_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_mma -s 1 -c 1 -f -o gpurun_out/${TAG}_mma_b1024 python tools/run_once.py 1024 96000 f16 > gpurun_out/${TAG}_ncu_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_mma -s 1 -c 1 -f -o gpurun_out/${TAG}_mma_strict_b1024 python tools/run_once.py 1024 48000 f16x3 >> gpurun_out/${TAG}_ncu_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_tcs -c 1 -f -o gpurun_out/${TAG}_tcs_b37888 python tools/run_once.py 37888 1500 f16 >> gpurun_out/${TAG}_ncu_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:delay_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_delay_b1024 python tools/delay_once.py 1024 1440000 365 >> gpurun_out/${TAG}_ncu_full.log 2>&1
  tail -3 gpurun_out/${TAG}_ncu_full.log
fi
