#!/usr/bin/env python3
"""Dev tool (GPU box): throughput of the stand-alone delay line (ntm_delay_forward) against its HBM roofline
(12 algorithmic bytes per sample: x and d read, y written)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import ntm_b200
from ntm_b200 import signals

dev = "cuda:0"
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for B, T, D in ((1, 480000, 365), (256, 1440000, 365), (1024, 1440000, 365), (1024, 1440000, 2000), (1024, 1440000, 6000), (1023, 1440001, 365)):
    x = torch.randn(B, 1, T, device=dev)
    d = signals.delay_trajectory_device(B, T, dev).reshape(B, 1, T) * (D / 365.0)
    dl = ntm_b200.TimeVaryingDelayLine(max_delay=D)
    dl.check_delay = False
    with torch.inference_mode():
        dl.init_buffer(B)
        dl(x, d)
        best = 1e9
        for _ in range(3):
            dl.init_buffer(B)
            e0.record(); y = dl(x, d); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    print(f"B={B:5d} T={T:8d} D={D:5d}: {best:8.3f} ms  {B*T/best/1e6:8.2f} Gsamples/s  {12*B*T/best/1e6:8.1f} GB/s  "
          f"checksum {float(y.double().sum()):.6f}", flush=True)
    del x, d, y
