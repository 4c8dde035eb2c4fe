#!/usr/bin/env python3
"""Dev tool (GPU box): time a few (batch, kernel-variant) points.  usage: quick_time.py mode "n,g" ["n,g" ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import ntm_b200
from ntm_b200 import lib, signals

dev = torch.device("cuda:0")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
z = np.load(os.path.join(ROOT, "tests/golden/ckpt_cfg2.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files if k not in ("model_type", "weights_dir")}
m = ntm_b200.RNN(1, 64, 1, False).to(dev)
m.load_state_dict(sd)
m.mode = sys.argv[1]
variants = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [(0, 0)]
L = lib.load()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.inference_mode():
    for B, T in ((256, 48000), (1024, 24000), (8192, 12000), (65536, 3000)):
        x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
        for n, g in variants:
            L.ntm_set_tuning(n, g)
            m.initialize_hidden()
            m(x[:, :, :600])
            best = 1e9
            for _ in range(3):
                m.initialize_hidden()
                e0.record(); m(x); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"B={B:6d} T={T:6d} {m.mode:5s} n={n:2d} g={g}: {best:9.3f} ms {B*T/best/1e6:9.3f} Gsamples/s {best*1e6/T:8.1f} ns/step", flush=True)
L.ntm_set_tuning(0, 0)
