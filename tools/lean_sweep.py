#!/usr/bin/env python3
"""Dev tool (GPU box): one line of ns/step per library build (NTM_B200_LIB, tools/ab_build.py) at the widths the lean 4-stream
mma.sync kernel serves: GRU f16 at 1 / 592 / 1024 streams, strict f16x3 at 592 / 1024, DiffDelGRU (cfg 3) f16 at 256."""
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def best_of(fn, n=3):
    fn()
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


row = []
with torch.inference_mode():
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg2"))
    for mode, widths in (("f16", ((1, 200000), (592, 96000), (1024, 96000))), ("f16x3", ((592, 48000), (1024, 48000)))):
        m.mode = mode
        for B, T in widths:
            x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
            def run():
                m.initialize_hidden()
                return m(x)
            ms = best_of(run)
            row.append(f"{mode} B={B} {ms*1e6/T:6.1f}")
    md = ntm_b200.DiffDelRNN(1, 64, 1, False, max_delay=signals.DELAY_MAX).to(dev)
    md.load_state_dict(load_ckpt("cfg3"))
    md.diffdel.check_delay = False
    md.mode = "f16"
    B, T = 256, 96000
    x = signals.stream_batch_device(B, T, dev, dur=30.0).reshape(B, 1, T)
    d = signals.delay_trajectory_device(B, T, dev).reshape(B, 1, T)
    ms = best_of(lambda: md.predict(x, d))
    row.append(f"cfg3 f16 {ms*1e6/(T+1024):6.1f}")
print(os.path.basename(lib.LIB_PATH).replace("libntm_b200", "").replace(".so", "") or "product", "|", " | ".join(row), flush=True)

# accuracy of the same build on golden signals (automatic dispatch = the lean kernel at batch 1): ESR against the reference's fp32 output
from conftest import load_golden
from oracle import c_oracle
import numpy as np
acc = []
with torch.inference_mode():
    for tag in ("cfg1", "cfg2"):
        g = load_golden(f"golden_{tag}")
        m = ntm_b200.RNN(1, 64, 1, False).to(dev)
        m.load_state_dict(load_ckpt(tag))
        for mode in ("f16", "f16x3"):
            m.mode = mode
            for sig in ("sweepnoise", "noise", "sine"):
                y = m.predict(torch.from_numpy(g[f"x_{sig}"]).to(dev).reshape(1, 1, -1)).cpu().numpy().reshape(-1)
                acc.append(f"{tag} {mode} {sig} esr={c_oracle.esr(y, g[f'y_{sig}']):.2e} max={np.max(np.abs(y - g[f'y_{sig}'])):.1e}")
    g = load_golden("golden_cfg3")
    md.mode = "f16"
    for sig in ("sweepnoise", "pulse", "sine"):
        y, pre = md.predict(torch.from_numpy(g[f"x_{sig}"]).to(dev).reshape(1, 1, -1), torch.from_numpy(g[f"d_{sig}"]).to(dev).reshape(1, 1, -1))
        acc.append(f"cfg3 f16 {sig} esr={c_oracle.esr(y.cpu().numpy().reshape(-1), g[f'y_{sig}']):.2e}")
print("   " + " | ".join(acc), flush=True)
