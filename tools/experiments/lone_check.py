#!/usr/bin/env python3
"""Dev tool (GPU box): the latency-regime widths of the mma.sync kernel (one CTA per SM: batch 1, cfg 3's 256 streams, 592) and cfg 2,
f16 and strict, plain GRU and DiffDelGRU, ns per step + the real-time block path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import ntm_b200
from ntm_b200 import lib, signals
from conftest import load_ckpt

dev = "cuda:0"
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.inference_mode():
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg2"))
    for mode in ("f16", "bf16", "f16x3"):
        m.mode = mode
        m.initialize_hidden(); m.warm_start()
        hw = m.hidden.clone()
        row = []
        for B, T in ((1, 200000), (256, 48000), (592, 48000), (1024, 48000)):
            x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
            m.hidden = hw.expand(1, B, 64).contiguous(); m(x[:, :, :1000])
            best = 1e9
            for _ in range(3):
                m.hidden = hw.expand(1, B, 64).contiguous()
                e0.record(); y = m(x); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            row.append(f"B={B}: {best*1e6/T:6.1f} ns/step")
        print(mode, " | ".join(row), flush=True)
    md = ntm_b200.DiffDelRNN(1, 64, 1, False, max_delay=signals.DELAY_MAX).to(dev)
    md.load_state_dict(load_ckpt("cfg3"))
    md.diffdel.check_delay = False
    B, T = 256, 30 * 48000
    xd = signals.stream_batch_device(B, T, dev, dur=30.0).reshape(B, 1, T)
    dd = signals.delay_trajectory_device(B, T, dev).reshape(B, 1, T)
    for mode in ("f16", "f16x3"):
        md.mode = mode
        md.predict(xd[:, :, :4800], dd[:, :, :4800]); torch.cuda.synchronize()
        e0.record(); md.predict(xd, dd); e1.record(); torch.cuda.synchronize()
        print(f"cfg3 DiffDelGRU 256 x 30 s {mode}: {e0.elapsed_time(e1)*1e6/T:6.1f} ns/step {B*T/e0.elapsed_time(e1)/1e6:.3f} Gs/s", flush=True)
