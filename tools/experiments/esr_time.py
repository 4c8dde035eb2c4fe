#!/usr/bin/env python3
"""Dev tool (GPU box): throughput of the on-device ESR / DCPreESR pass against its HBM roofline (8 B per sample)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import ntm_b200

dev = "cuda:0"
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for B, T in ((1, 480000), (1024, 480000), (1024, 1440000)):
    t = 0.2 * torch.randn(B, 1, T, device=dev)
    o = t + 0.02 * torch.randn(B, 1, T, device=dev)
    for name, L in (("ESR", ntm_b200.ESRLoss()), ("DCPreESR", ntm_b200.DCPreESR())):
        L(o, t)
        best = 1e9
        for _ in range(3):
            e0.record(); v = L(o, t); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"B={B:5d} T={T:8d} {name:9s}: {best:8.3f} ms  {B*T/best/1e6:8.2f} Gsamples/s  {8*B*T/best/1e6:8.1f} GB/s  loss {float(v):.6f}", flush=True)
    del t, o
