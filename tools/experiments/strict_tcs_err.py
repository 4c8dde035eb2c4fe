#!/usr/bin/env python3
"""Dev tool (GPU box): strict mode on the stream-major tcgen05 kernel against the reference goldens (max-abs per signal)."""
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

L = lib.load()
with torch.inference_mode():
    for tag in ("cfg1", "cfg2"):
        g = load_golden(f"golden_{tag}")
        m = ntm_b200.RNN(1, 64, 1, False).to("cuda:0")
        m.load_state_dict(load_ckpt(tag))
        for mode, tune in (("fp32", (0, 0)), ("f16x3", (8, 3)), ("f16x3", (1, 4)), ("f16x3", (2, 4))):
            m.mode = mode
            L.ntm_set_tuning(*tune)
            row = []
            for sig in SIGNALS:
                y = m.predict(torch.from_numpy(g[f"x_{sig}"]).to("cuda:0").reshape(1, 1, -1)).cpu().numpy().reshape(-1)
                row.append(f"{sig} {np.max(np.abs(y - g[f'y_{sig}'])):.1e}/{np.max(np.abs(y - g[f'y64_{sig}'])):.1e}")
            print(os.path.basename(lib.LIB_PATH), tag, mode, tune, " | ".join(row), flush=True)
L.ntm_set_tuning(0, 0)
