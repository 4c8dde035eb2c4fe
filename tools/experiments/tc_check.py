#!/usr/bin/env python3
"""Dev tool (GPU box): ESR / max-abs of the tensor-core modes against the golden reference outputs, plus timing.
   usage: tc_check.py [modes...]   (default: f16 tf32 bf16)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import ntm_b200
from ntm_b200 import lib, signals
from conftest import SIGNALS, load_ckpt, load_golden

dev = "cuda:0"
modes = sys.argv[1:] or ["f16", "tf32", "bf16"]


def esr(y, t):
    return float(np.sum((y - t) ** 2) / (np.sum(t ** 2) + 1e-5))


with torch.inference_mode():
    for tag in ("cfg1", "cfg2"):
        g = load_golden(f"golden_{tag}")
        for mode in ["fp32"] + modes + [mm + "/mma" for mm in modes]:
            lib.load().ntm_set_tuning(8, 3) if mode.endswith("/mma") else lib.load().ntm_set_tuning(0 if mode == "fp32" else 32, 0 if mode == "fp32" else 2)
            mode = mode.split("/")[0]
            m = ntm_b200.RNN(1, 64, 1, False).to(dev)
            m.load_state_dict(load_ckpt(tag))
            m.mode = mode
            m.initialize_hidden(); m.warm_start()
            hw = float(np.max(np.abs(m.hidden.cpu().numpy().reshape(-1) - g["h_warm"])))
            row = [f"{tag} {mode:5s} h_warm_err={hw:.1e}"]
            for sig in SIGNALS:
                y = m.predict(torch.from_numpy(g[f"x_{sig}"]).to(dev).reshape(1, 1, -1)).cpu().numpy().reshape(-1)
                row.append(f"{sig}: esr={esr(y, g[f'y_{sig}']):.1e} max={np.max(np.abs(y - g[f'y_{sig}'])):.1e} (floor {float(g[f'floor_{sig}']):.0e})")
            print(" | ".join(row), flush=True)
    # batch consistency + timing
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg2"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L = lib.load()
    for B, T in ((1024, 24000), (4096, 12000), (16384, 6000), (65536, 3000)):
        x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
        m.mode = "fp32"
        L.ntm_set_tuning(0, 0)
        yref = m.predict(x[:, :, :4800])
        for mode in modes:
            m.mode = mode
            for (n, g_) in ((0, 0), (32, 1), (64, 1), (32, 2), (64, 2), (8, 3), (16, 3)):
                if B // max(n * g_, 1) > 20000:
                    continue
                L.ntm_set_tuning(n, g_)
                y = m.predict(x[:, :, :4800])
                err = float((y - yref).abs().max())
                es = float(((y - yref) ** 2).sum() / ((yref ** 2).sum() + 1e-5))
                m.initialize_hidden()
                e0.record(); m(x); e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                print(f"B={B:6d} T={T:6d} {mode:5s} n={n:2d} g={g_}: {ms:9.3f} ms {B*T/ms/1e6:9.3f} Gsamples/s {ms*1e6/T:8.1f} ns/step"
                      f"   vs fp32 kernel: max={err:.1e} esr={es:.1e}", flush=True)
        L.ntm_set_tuning(0, 0)
