#!/usr/bin/env python3
"""Dev tool (GPU box): 4-streams-per-CTA (HALF) variant of the mma.sync kernel vs the 8-stream one: equality, timing."""
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
L = lib.load()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.inference_mode():
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg2"))
    for mode in ("f16", "tf32"):
        m.mode = mode
        L.ntm_set_tuning(8, 3)
        m.initialize_hidden(); m.warm_start()
        hw = m.hidden.clone()
        for B, T in [(1, 200000), (256, 48000), (592, 48000), (1024, 24000), (1184, 24000), (1536, 24000), (2048, 24000), (4096, 12000)] if len(sys.argv) > 1 else ((1, 200000), (5, 100000), (256, 48000), (512, 48000), (592, 48000), (1024, 24000)):
            x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
            res = {}
            for tune in ((8, 3), (4, 3), (0, 0)):
                L.ntm_set_tuning(*tune)
                m.hidden = hw.expand(1, B, 64).contiguous(); m(x[:, :, :1000])
                best = 1e9
                for _ in range(2):
                    m.hidden = hw.expand(1, B, 64).contiguous()
                    e0.record(); y = m(x); e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                res[tune] = (y, m.hidden.clone(), best)
            eq = torch.equal(res[(8, 3)][0], res[(4, 3)][0]) and torch.equal(res[(8, 3)][1], res[(4, 3)][1])
            print(f"{mode} B={B:5d} T={T}: 8/CTA {res[(8,3)][2]*1e6/T:7.1f} ns/step | 4/CTA {res[(4,3)][2]*1e6/T:7.1f} ns/step | auto {res[(0,0)][2]*1e6/T:7.1f} ns/step "
                  f"({B*T/res[(0,0)][2]/1e6:6.3f} Gs/s) | identical {eq}", flush=True)
L.ntm_set_tuning(0, 0)
