#!/usr/bin/env python3
"""Dev tool (GPU box): max-abs error vs the golden reference outputs for the accurate and the MUFU-approx
activation flavours of the fp32 kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import ntm_b200
from ntm_b200 import lib
from conftest import SIGNALS, load_ckpt, load_golden

dev = "cuda:0"
with torch.inference_mode():
    for tag in ("cfg1", "cfg2"):
        m = ntm_b200.RNN(1, 64, 1, False).to(dev)
        m.load_state_dict(load_ckpt(tag))
        g = load_golden(f"golden_{tag}")
        for sig in SIGNALS:
            row = []
            for ks in (4, 0x104):
                lib.load().ntm_set_tuning(1, ks)
                y = m.predict(torch.from_numpy(g[f"x_{sig}"]).to(dev).reshape(1, 1, -1)).cpu().numpy().reshape(-1)
                row.append((np.max(np.abs(y - g[f"y_{sig}"])), np.max(np.abs(y - g[f"y64_{sig}"]))))
            print(f"{tag} {sig:14s} floor={float(g[f'floor_{sig}']):.2e}  acc: vs_ref={row[0][0]:.2e} vs_f64={row[0][1]:.2e}"
                  f"   fast: vs_ref={row[1][0]:.2e} vs_f64={row[1][1]:.2e}")
