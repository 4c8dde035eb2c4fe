#!/usr/bin/env python3
"""Dev tool (GPU box): the two-group pipelined mma.sync kernel (tuning (8, 5) / (16, 5)) against the established forms
((4, 3) four streams per CTA, (8, 3) eight): agreement of outputs / final state and ns per step over the batch widths.
usage: mma2_check.py [mode]"""
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
mode = sys.argv[1] if len(sys.argv) > 1 else "f16"
TUNES = ((4, 3), (8, 3), (0, 0))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.inference_mode():
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg2"))
    m.mode = "fp32"
    m.initialize_hidden(); m.warm_start()
    hw = m.hidden.clone()
    for B, T in ((1, 100000), (5, 100000), (8, 100000), (13, 30011), (592, 48000), (1024, 48000), (1184, 48000), (2048, 24000),
                 (2368, 24000), (4096, 12000), (8192, 12000), (16000, 6000)):
        x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
        m.mode = "fp32"
        Tr = min(T, 6000)
        m.hidden = hw.expand(1, B, 64).contiguous()
        yref = m(x[:, :, :Tr]).double()
        m.mode = mode
        row = []
        for tune in TUNES:
            L.ntm_set_tuning(*tune)
            m.hidden = hw.expand(1, B, 64).contiguous(); m(x[:, :, :1000])
            best = 1e9
            for _ in range(2):
                m.hidden = hw.expand(1, B, 64).contiguous()
                e0.record(); y = m(x); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            kern = L.ntm_query(lib.Q_LAST_KERNEL)
            yd = y[:, :, :Tr].double()
            esr = float(((yd - yref) ** 2).sum() / ((yref ** 2).sum() + 1e-12))
            # chunk-boundary / tail handling: the same call in two pieces must give the same samples
            m.hidden = hw.expand(1, B, 64).contiguous()
            cut = 1000 + 37
            y2 = torch.cat([m(x[:, :, :cut]), m(x[:, :, cut:Tr])], 2)
            same = bool(torch.equal(y2, y[:, :, :Tr]))
            row.append(f"{tune}k{kern} {best*1e6/T:6.1f}ns esr={esr:.1e}{'' if same else ' SPLIT-MISMATCH'}{'' if bool(torch.isfinite(y).all()) else ' NONFINITE'}")
        print(f"{mode} B={B:5d} T={T:6d}: " + " | ".join(row), flush=True)
L.ntm_set_tuning(0, 0)
