#!/usr/bin/env python3
"""Dev tool (GPU box): the stream-major tcgen05 kernel with partially filled tiles (32 / 64 / 96 / 128 streams per tile, two staggered
tiles per CTA) against the mma.sync kernel and the one-full-tile-per-SM schedule over the mid-range batch widths; agreement with the
mma.sync output on 256 sampled streams.  usage: tcs_fill.py [mode]"""
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
VAR = 15


def tcs(tiles, fill):
    return (tiles + 4 * ((VAR | ((fill // 32) % 4) << 6) + 1), 4)


e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.inference_mode():
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg2"))
    m.mode = mode
    m.initialize_hidden(); m.warm_start()
    hw = m.hidden.clone()
    for B, T in ((4096, 6000), (6144, 6000), (8192, 6000), (9472, 6000), (12000, 6000), (16384, 6000), (18944, 4000), (24000, 4000),
                 (28416, 3000), (32768, 3000), (37888, 3000), (48000, 3000), (65536, 3000)):
        x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
        row, ref = [], None
        for name, tune in (("mma8", (8, 3)), ("t1f128", tcs(1, 128)), ("t2f32", tcs(2, 32)), ("t2f64", tcs(2, 64)), ("t2f96", tcs(2, 96)),
                           ("t2f128", tcs(2, 128)), ("auto", (0, 0))):
            L.ntm_set_tuning(*tune)
            m.hidden = hw.expand(1, B, 64).contiguous(); m(x[:, :, :300])
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
