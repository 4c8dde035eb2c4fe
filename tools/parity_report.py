#!/usr/bin/env python3
"""GPU box: achieved parity numbers, one record per (checkpoint, signal, mode, kernel): max-abs against the reference's fp32
output and against the float64 ground truth, ESR against the reference output, next to the reference's own fp32-vs-fp64 floor.
Covers the three BASELINE checkpoints on the seven golden signals and all 12 shipped _BEST checkpoints.
    python tools/parity_report.py > gpurun_out/r02_parity.json        (committed as profiles/r02_parity.json)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import ntm_b200
from ntm_b200 import lib
from oracle import c_oracle
from conftest import SIGNALS, load_ckpt, load_golden

DEV = "cuda:0"
L = lib.load()
# (mode, kernel selector, label)
CASES = [("fp32", (0, 0)), ("f16x3", (4, 3)), ("f16x3", (4, 6)), ("f16x3", (8, 3)), ("f16x3", (1, 4)), ("f16", (4, 3)), ("f16", (4, 6)),
         ("f16", (8, 3)), ("f16", (1, 4)),
         ("tf32", (8, 3)), ("tf32", (1, 4)), ("bf16", (1, 4))]
rows = []


def d(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def rec(ckpt, sig, mode, kern, y, ref, truth, floor, what="y"):
    rows.append({"checkpoint": ckpt, "signal": sig, "output": what, "mode": mode, "kernel": kern,
                 "max_abs_vs_ref_fp32": float(np.max(np.abs(y - ref))),
                 "max_abs_vs_fp64": None if truth is None else float(np.max(np.abs(y - truth))),
                 "esr_vs_ref_fp32": float(c_oracle.esr(y, ref)), "reference_floor": float(floor)})


with torch.inference_mode():
    for tag in ("cfg1", "cfg2"):
        g = load_golden(f"golden_{tag}")
        m = ntm_b200.RNN(1, 64, 1, False).to(DEV)
        m.load_state_dict(load_ckpt(tag))
        for mode, tune in CASES:
            m.mode = mode
            L.ntm_set_tuning(*tune)
            for sig in SIGNALS:
                y = m.predict(d(g[f"x_{sig}"]).reshape(1, 1, -1)).cpu().numpy().reshape(-1)
                rec(tag, sig, mode, lib.KERNEL_NAMES[L.ntm_query(lib.Q_LAST_KERNEL)] + f" {tune}", y, g[f"y_{sig}"], g[f"y64_{sig}"],
                    g[f"floor_{sig}"])
    g = load_golden("golden_cfg3")
    m = ntm_b200.DiffDelRNN(1, 64, 1, False, max_delay=int(g["max_delay"])).to(DEV)
    m.load_state_dict(load_ckpt("cfg3"))
    for mode, tune in CASES:
        m.mode = mode
        L.ntm_set_tuning(*tune)
        for sig in SIGNALS:
            y, pre = m.predict(d(g[f"x_{sig}"]).reshape(1, 1, -1), d(g[f"d_{sig}"]).reshape(1, 1, -1))
            k = lib.KERNEL_NAMES[L.ntm_query(lib.Q_LAST_KERNEL)] + f" {tune}"
            rec("cfg3", sig, mode, k, pre.cpu().numpy().reshape(-1), g[f"pre_{sig}"], g[f"pre64_{sig}"], g[f"floor_{sig}"], "pre_d")
            rec("cfg3", sig, mode, k, y.cpu().numpy().reshape(-1), g[f"y_{sig}"], None, g[f"floor_{sig}"], "y")
    L.ntm_set_tuning(0, 0)
    g = load_golden("golden_best12")
    for i in range(int(g["n"])):
        pre = f"w{i}_"
        sd = {k[len(pre):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre)}
        kind = str(g[f"kind{i}"])
        if kind == "GRU":
            m = ntm_b200.RNN(1, 64, 1, False).to(DEV)
        else:
            m = ntm_b200.DiffDelRNN(1, 64, 1, False, max_delay=int(g["max_delay"])).to(DEV)
        m.load_state_dict(sd)
        for mode in ("fp32", "f16x3", "f16"):
            m.mode = mode
            for sig in g["signals"]:
                x = d(g[f"x_{sig}"]).reshape(1, 1, -1)
                if kind == "GRU":
                    y = m.predict(x).cpu().numpy().reshape(-1)
                else:
                    y = m.predict(x, d(g[f"d_{sig}"]).reshape(1, 1, -1))[0].cpu().numpy().reshape(-1)
                rec(f"_BEST #{i} ({kind})", str(sig), mode, lib.KERNEL_NAMES[L.ntm_query(lib.Q_LAST_KERNEL)] + " auto", y,
                    g[f"y{i}_{sig}"], None, g[f"floor{i}_{sig}"])
ten = os.path.join(ROOT, "gpurun_out", "parity_10s.json")
out = {"records": rows, "ten_seconds": json.load(open(ten)) if os.path.exists(ten) else None}
worst = {}
for r in rows:
    if r["reference_floor"] < 3e-6 and r["signal"] != "silence":
        key = r["mode"]
        worst[key] = max(worst.get(key, 0.0), r["max_abs_vs_ref_fp32"] if r["mode"] in ("fp32", "f16x3") else r["esr_vs_ref_fp32"])
out["worst_on_stable_signals"] = {k: {"max_abs_vs_ref_fp32" if k in ("fp32", "f16x3") else "esr_vs_ref_fp32": v} for k, v in worst.items()}
print(json.dumps(out, indent=1))
