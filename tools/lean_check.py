#!/usr/bin/env python3
"""Dev tool (GPU box): the lean 4-stream mma.sync kernel (gru_mma4.cu, tuning (4, 6)) against the general kernel's 4-stream form
((4, 3)): bit-identity of outputs and final state (ragged widths, odd lengths, skip connection, unaligned rows), ns per step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
    for mode in ("f16", "bf16", "f16x3"):
        for skip in (False, True):
            m = ntm_b200.RNN(1, 64, 1, skip).to(dev)
            m.load_state_dict(load_ckpt("cfg2"))
            m.mode = mode
            m.initialize_hidden(); m.warm_start()
            hw = m.hidden.clone()
            for B, T, off in ((1, 1000, 0), (5, 1027, 0), (7, 513, 1), (600, 2051, 0), (1024, 4097, 3), (1184, 4096, 0)):
                big = signals.stream_batch_device(B, T + 8, dev, dur=10.0)
                x = big[:, off:off + T].reshape(B, 1, T)             # off != 0: rows not 16-byte aligned
                outs = []
                for tune in ((4, 3), (4, 6)):
                    L.ntm_set_tuning(*tune)
                    m.hidden = hw.expand(1, B, 64).contiguous()
                    y1 = m(x[:, :, :T // 3])
                    y2 = m(x[:, :, T // 3:])
                    outs.append((torch.cat([y1, y2], 2), m.hidden.clone()))
                same = torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
                print(f"{mode} skip={skip} B={B} T={T} off={off}: identical={same} finite={bool(torch.isfinite(outs[1][0]).all())}", flush=True)
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg2"))
    m.initialize_hidden(); m.warm_start()
    hw = m.hidden.clone()
    for smode in ("f16", "f16x3"):
        m.mode = smode
        for B, T in ((1, 100000), (592, 48000), (597, 48000), (1024, 48000), (1184, 48000)):
            x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
            row = []
            for tune in ((4, 3), (4, 6), (0, 0)):
                L.ntm_set_tuning(*tune)
                m.hidden = hw.expand(1, B, 64).contiguous(); m(x[:, :, :1000])
                best = 1e9
                for _ in range(3):
                    m.hidden = hw.expand(1, B, 64).contiguous()
                    e0.record(); y = m(x); e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                row.append(f"{tune} {best*1e6/T:6.1f} ns/step")
            print(f"{smode} B={B}: " + " | ".join(row), flush=True)
L.ntm_set_tuning(0, 0)
