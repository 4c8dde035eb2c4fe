#!/usr/bin/env python3
"""Dev tool (GPU box): the strict tensor-core mode (f16x3) against the reference goldens -- achieved max-abs error next to
the exact-fp32 CUDA-core kernel's -- and its speed per width.  usage: strict_check.py [quick]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import ntm_b200
from ntm_b200 import lib, signals
from conftest import SIGNALS, load_ckpt, load_golden

dev = "cuda:0"
L = lib.load()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def d(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


with torch.inference_mode():
    for tag in ("cfg1", "cfg2"):
        g = load_golden(f"golden_{tag}")
        m = ntm_b200.RNN(1, 64, 1, False).to(dev)
        m.load_state_dict(load_ckpt(tag))
        for mode, tune in (("fp32", (0, 0)), ("f16x3", (4, 3)), ("f16x3", (8, 3)), ("f16x3", (0, 0))):
            m.mode = mode
            L.ntm_set_tuning(*tune)
            m.initialize_hidden(); m.warm_start()
            hw = float(np.max(np.abs(m.hidden.cpu().numpy().reshape(-1) - g["h_warm"])))
            row = [f"h_warm {hw:.1e}"]
            for sig in SIGNALS:
                y = m.predict(d(g[f"x_{sig}"]).reshape(1, 1, -1)).cpu().numpy().reshape(-1)
                row.append(f"{sig} {np.max(np.abs(y - g[f'y_{sig}'])):.1e}/{np.max(np.abs(y - g[f'y64_{sig}'])):.1e} (floor {float(g[f'floor_{sig}']):.1e})")
            print(tag, mode, tune, lib.KERNEL_NAMES.get(L.ntm_query(lib.Q_LAST_KERNEL)), "|", " | ".join(row), flush=True)
    L.ntm_set_tuning(0, 0)
    # all 12 shipped checkpoints (plain GRU ones and DiffDel)
    g = load_golden("golden_best12")
    worst = {}
    for i in range(int(g["n"])):
        pre = f"w{i}_"
        sd = {k[len(pre):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre)}
        kind = str(g[f"kind{i}"])
        if kind == "GRU":
            m = ntm_b200.RNN(1, 64, 1, False).to(dev)
        else:
            m = ntm_b200.DiffDelRNN(1, 64, 1, False, max_delay=int(g["max_delay"])).to(dev)
        m.load_state_dict(sd)
        for mode in ("fp32", "f16x3"):
            m.mode = mode
            for sig in g["signals"]:
                x = d(g[f"x_{sig}"]).reshape(1, 1, -1)
                if kind == "GRU":
                    y = m.predict(x).cpu().numpy().reshape(-1)
                else:
                    y = m.predict(x, d(g[f"d_{sig}"]).reshape(1, 1, -1))[0].cpu().numpy().reshape(-1)
                err = float(np.max(np.abs(y - g[f"y{i}_{sig}"])))
                print(f"best12 #{i} {kind} {mode} {sig}: max-abs {err:.2e} (floor {float(g[f'floor{i}_{sig}']):.1e})", flush=True)
    # speed
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg2"))
    m.initialize_hidden(); m.warm_start()
    hw = m.hidden.clone()
    widths = ((1, 100000), (256, 48000), (1024, 48000), (2048, 24000), (8192, 6000)) if len(sys.argv) < 2 else ((1, 100000), (1024, 24000))
    for B, T in widths:
        x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
        for mode, tune in (("fp32", (0, 0)), ("f16x3", (4, 3)), ("f16x3", (8, 3)), ("f16", (0, 0))):
            m.mode = mode
            L.ntm_set_tuning(*tune)
            m.hidden = hw.expand(1, B, 64).contiguous(); m(x[:, :, :1000])
            best = 1e9
            for _ in range(2):
                m.hidden = hw.expand(1, B, 64).contiguous()
                e0.record(); y = m(x); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"speed {mode} tune={tune} B={B} T={T}: {best*1e6/T:7.1f} ns/step {B*T/best/1e6:7.3f} Gs/s", flush=True)
    L.ntm_set_tuning(0, 0)
