#!/usr/bin/env python3
"""Dev tool (GPU box): the stream-major tcgen05 kernel per build (NTM_B200_LIB: threads per stream, tools/ab_build.py gru_tcs
"uw4:-DNTM_TCS_UW=4") over the mid-range and large batch widths: one tile per CTA (t1), two (t2), automatic dispatch, against the
mma.sync kernel; agreement with the mma.sync output on 256 sampled streams.  usage: tcs_uw_sweep.py [mode]"""
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
widths = [int(w) for w in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8192, 12000, 16384, 18944, 24000, 28416, 37888, 65536]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
print(os.path.basename(lib.LIB_PATH), flush=True)
with torch.inference_mode():
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg2"))
    m.mode = mode
    m.initialize_hidden(); m.warm_start()
    hw = m.hidden.clone()
    for B in widths:
        T = 6000 if B <= 16384 else 3000
        x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
        row, ref = [], None
        for name, tune in (("mma8", (8, 3)), ("t1", (1 + 4 * 16, 4)), ("t2", (2 + 4 * 16, 4)), ("auto", (0, 0))):
            if L.ntm_set_tuning(*tune) != 0:
                continue
            try:
                m.hidden = hw.expand(1, B, 64).contiguous(); m(x[:, :, :300])
            except RuntimeError as e:
                row.append(f"{name} n/a")
                continue
            best = 1e9
            for _ in range(2):
                m.hidden = hw.expand(1, B, 64).contiguous()
                e0.record(); y = m(x); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            ys = y[:: max(1, B // 256)].double()
            if ref is None:
                ref = ys
            esr = float(((ys - ref) ** 2).sum() / (ref ** 2).sum())
            row.append(f"{name} {B*T/best/1e6:6.2f}{'' if esr < 1e-5 else ' ESR=%.1e' % esr}")
        print(f"{mode} B={B:6d}: " + " | ".join(row) + "  Gs/s", flush=True)
L.ntm_set_tuning(0, 0)
