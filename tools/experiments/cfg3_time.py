#!/usr/bin/env python3
"""Dev tool (GPU box): BASELINE cfg 3 (DiffDelGRU, 256 streams x 30 s, predict() semantics) per arithmetic mode."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import ntm_b200
from ntm_b200 import signals
from conftest import load_ckpt

dev = "cuda:0"
B, T = 256, 30 * 48000
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.inference_mode():
    md = ntm_b200.DiffDelRNN(1, 64, 1, False, max_delay=signals.DELAY_MAX).to(dev)
    md.load_state_dict(load_ckpt("cfg3"))
    md.diffdel.check_delay = False
    x = signals.stream_batch_device(B, T, dev, dur=30.0).reshape(B, 1, T)
    d = signals.delay_trajectory_device(B, T, dev).reshape(B, 1, T)
    for mode in ("f16", "fp32"):
        md.mode = mode
        md.predict(x[:, :, :4800], d[:, :, :4800])
        torch.cuda.synchronize()
        e0.record(); y, p = md.predict(x, d); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"cfg3 {mode}: {ms:.1f} ms, {B*T/ms/1e6:.3f} Gsamples/s, {ms*1e6/T:.1f} ns/step", flush=True)
