python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01_pytest_gpu2.txt; tail -5 gpurun_out/r01_pytest_gpu2.txt
ncu --set full --clock-control none --import-source on -k regex:gru_tc -s 1 -c 1 -f -o gpurun_out/r01_tc_b1024 python tools/run_once.py 1024 4800 f16 > gpurun_out/ncu_tc1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru_tc -s 1 -c 1 -f -o gpurun_out/r01_tc_b16384 python tools/run_once.py 16384 2400 f16 64 2 > gpurun_out/ncu_tc2.log 2>&1
tail -2 gpurun_out/ncu_tc1.log gpurun_out/ncu_tc2.log
