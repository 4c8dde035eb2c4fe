#!/usr/bin/env python3
"""Dev tool (GPU box): variants of the stream-major tcgen05 kernel -- ESR against the golden reference, timing.
   usage: tcs_variants.py var [var ...]    (tuning value = tiles + 4 * (var + 1), ksplit 4)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import ntm_b200
from ntm_b200 import lib, signals
from conftest import load_ckpt, load_golden

dev = "cuda:0"
L = lib.load()
variants = [int(v) for v in sys.argv[1:]] or [0]
BT = ((37888, 3000),)


def esr(y, t):
    return float(np.sum((y - t) ** 2) / (np.sum(t ** 2) + 1e-5))


e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.inference_mode():
    models = {}
    for tag in ("cfg1", "cfg2"):
        m = ntm_b200.RNN(1, 64, 1, False).to(dev)
        m.load_state_dict(load_ckpt(tag))
        m.mode = "f16"
        models[tag] = m
    xs = {bt: signals.stream_batch_device(bt[0], bt[1], dev, dur=10.0).reshape(bt[0], 1, bt[1]) for bt in BT}
    for var in variants:
        row = [f"var={var:2d}"]
        for tiles in (2, 1):
            L.ntm_set_tuning(tiles + 4 * (var + 1), 4)
            if tiles == 2:
                for tag in ("cfg1", "cfg2"):
                    g = load_golden(f"golden_{tag}")
                    for sig in ("sweepnoise_lo", "noise", "sine1k"):
                        y = models[tag].predict(torch.from_numpy(g[f"x_{sig}"]).to(dev).reshape(1, 1, -1)).cpu().numpy().reshape(-1)
                        row.append(f"{tag}/{sig}: {esr(y, g[f'y_{sig}']):.1e}")
            m = models["cfg2"]
            for (B, T), x in xs.items():
                m.initialize_hidden(); m(x[:, :, :200])
                best = 1e9
                for _ in range(3):
                    m.initialize_hidden()
                    e0.record(); m(x); e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                row.append(f"tiles={tiles} B={B}: {B*T/best/1e6:7.3f} Gs/s {best*1e-3*1.965e9/T/(B/148):5.1f} clk/ss (k{lib.query(lib.Q_LAST_KERNEL)})")
        print(" | ".join(row), flush=True)
L.ntm_set_tuning(0, 0)
