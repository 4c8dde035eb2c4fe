#!/usr/bin/env python3
"""Dev tool (GPU box): dynamic (job-queue) vs static schedule of the stream-major kernel -- identical results, timing.
tuning value = tiles + 4 * (var + 1); var 31 = default variant, var 63 = default variant with the static schedule."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import ntm_b200
from ntm_b200 import lib, signals
from conftest import load_ckpt

dev = "cuda:0"
L = lib.load()
DYN = lambda tiles: (tiles + 4 * 32, 4)
STA = lambda tiles: (tiles + 4 * 64, 4)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.inference_mode():
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg2"))
    m.mode = "f16"
    m.initialize_hidden(); m.warm_start()
    hw = m.hidden.clone()

    def run(xb, tune):
        L.ntm_set_tuning(*tune)
        m.hidden = hw.expand(1, xb.shape[0], 64).contiguous()
        y = m(xb)
        return y, m.hidden.clone()

    for B, T in ((40000, 1000), (65536 + 77, 777), (19100, 1500)):
        x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
        for tiles in (2, 1):
            ys, hs = run(x, STA(tiles))
            yd, hd = run(x, DYN(tiles))
            print(f"B={B} T={T} tiles={tiles}: dynamic == static: y {bool(torch.equal(ys, yd))} h {bool(torch.equal(hs, hd))} "
                  f"max|dy| {float((ys - yd).abs().max()):.1e}", flush=True)
        del x
    for B, T in ((20000, 3000), (30000, 3000), (37888, 3000), (40000, 3000), (50000, 3000), (65536, 3000)):
        x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
        row = [f"B={B:6d} T={T}"]
        for name, tune in (("mma", (8, 3)), ("static2", STA(2)), ("dyn2", DYN(2)), ("static1", STA(1)), ("dyn1", DYN(1)), ("auto", (0, 0))):
            run(x[:, :, :300], tune)
            best = 1e9
            for _ in range(2):
                L.ntm_set_tuning(*tune)
                m.hidden = hw.expand(1, B, 64).contiguous()
                e0.record(); m(x); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            row.append(f"{name} {B*T/best/1e6:6.2f} (k{lib.query(lib.Q_LAST_KERNEL)})")
        print(" | ".join(row), flush=True)
        del x
L.ntm_set_tuning(0, 0)
