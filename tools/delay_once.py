#!/usr/bin/env python3
"""Dev tool (GPU box): one stand-alone delay-line pass over B x T samples (for ncu captures).  usage: delay_once.py B T D"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ntm_b200
from ntm_b200 import signals

B, T, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = torch.randn(B, 1, T, device="cuda:0")
d = signals.delay_trajectory_device(B, T, "cuda:0").reshape(B, 1, T) * (D / 365.0)
dl = ntm_b200.TimeVaryingDelayLine(max_delay=D)
dl.check_delay = False
with torch.inference_mode():
    for _ in range(2):
        dl.init_buffer(B)
        y = dl(x, d)
print(float(y.double().sum()))
