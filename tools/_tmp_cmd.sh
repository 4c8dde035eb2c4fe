timeout 240 python tools/experiments/delay_time.py 2>&1 | tail -6
timeout 600 python -m pytest tests -m gpu -q -x -k "delay or driver or diffdel or caller or hidden" 2>&1 | tail -2
