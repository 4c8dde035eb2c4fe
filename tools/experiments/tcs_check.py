#!/usr/bin/env python3
"""Dev tool (GPU box): the stream-major tcgen05 kernel (csrc/gru_tcs.cu) -- ESR / max-abs against the golden reference
outputs and the fp32 kernel, then timing against the mma.sync kernel.   usage: tcs_check.py [quick]"""
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
L = lib.load()


def esr(y, t):
    return float(np.sum((y - t) ** 2) / (np.sum(t ** 2) + 1e-5))


with torch.inference_mode():
    for tag in ("cfg1", "cfg2"):
        g = load_golden(f"golden_{tag}")
        for mode, tune in (("f16", (8, 3)), ("f16", (1, 4)), ("f16", (2, 4)), ("bf16", (2, 4))):
            L.ntm_set_tuning(*tune)
            m = ntm_b200.RNN(1, 64, 1, False).to(dev)
            m.load_state_dict(load_ckpt(tag))
            m.mode = mode
            m.initialize_hidden(); m.warm_start()
            hw = float(np.max(np.abs(m.hidden.cpu().numpy().reshape(-1) - g["h_warm"])))
            row = [f"{tag} {mode:5s} tune={tune} kernel={lib.query(lib.Q_LAST_KERNEL)} h_warm_err={hw:.1e}"]
            for sig in SIGNALS:
                y = m.predict(torch.from_numpy(g[f"x_{sig}"]).to(dev).reshape(1, 1, -1)).cpu().numpy().reshape(-1)
                row.append(f"{sig}: esr={esr(y, g[f'y_{sig}']):.1e} max={np.max(np.abs(y - g[f'y_{sig}'])):.1e}")
            print(" | ".join(row), flush=True)
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg2"))
    # ragged multi-tile batch vs the fp32 kernel, segmentation invariance, skip
    B, T = 300, 4000
    x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
    m.mode = "fp32"; L.ntm_set_tuning(0, 0)
    yref = m.predict(x)
    m.mode = "f16"
    for tune in ((1, 4), (2, 4)):
        L.ntm_set_tuning(*tune)
        y = m.predict(x)
        per = (((y - yref) ** 2).sum(2) / ((yref ** 2).sum(2) + 1e-5)).max()
        m.initialize_hidden(); m.warm_start()
        m.hidden = m.hidden.expand(1, B, 64).contiguous()
        parts = [m(x[:, :, s0:s1]) for s0, s1 in ((0, 1), (1, 2), (2, 67), (67, 2000), (2000, 4000))]
        seg = float((torch.cat(parts, 2) - y).abs().max())
        one = float((m.predict(x[17:18]) - y[17:18]).abs().max())
        print(f"ragged B={B} tune={tune} kernel={lib.query(lib.Q_LAST_KERNEL)}: max per-stream ESR vs fp32 kernel {float(per):.2e}, "
              f"max|err| {float((y - yref).abs().max()):.2e}, segmentation diff {seg:.1e}, single-stream diff {one:.1e}", flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sizes = ((18944, 3000), (37888, 3000), (65536, 3000)) if len(sys.argv) < 2 else ((37888, 2000),)
    for B, T in sizes:
        x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
        for tune in ((8, 3), (1, 4), (2, 4), (0, 0)):
            L.ntm_set_tuning(*tune)
            m.initialize_hidden(); m(x[:, :, :200])
            best = 1e9
            for _ in range(2):
                m.initialize_hidden()
                e0.record(); m(x); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"B={B:6d} T={T:5d} f16 tune={tune} kernel={lib.query(lib.Q_LAST_KERNEL)}: {best:9.3f} ms "
                  f"{B*T/best/1e6:8.3f} Gsamples/s  {best*1e6/T:8.1f} ns/step  {best*1e-3*1.965e9/T/(B/148):6.1f} clk/stream-step/SM", flush=True)
    L.ntm_set_tuning(0, 0)
