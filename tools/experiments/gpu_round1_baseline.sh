set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01_pytest_gpu.txt
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 1 --warmup 3 --seconds 2 --no-aux --no-cpu > gpurun_out/r01_bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru_fp32 -c 1 -f -o gpurun_out/r01_fp32 python tools/run_once.py 1024 4800 > gpurun_out/r01_ncu_full.log 2>&1
tail -3 gpurun_out/r01_pytest_gpu.txt; cat gpurun_out/r01_bench.json
