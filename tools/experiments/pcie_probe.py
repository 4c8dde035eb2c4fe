#!/usr/bin/env python3
"""Dev tool (GPU box): what the host link gives, and where predict_host's time goes.
   1-D and 2-D (time-chunk shaped) pinned copies each way and both ways at once, then RNN.predict_host over a few
   chunk lengths.  usage: pcie_probe.py [B] [seconds]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ntm_b200
from ntm_b200 import signals

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
T = int(sec * 48000)
dev = torch.device("cuda:0")
print("cpus", os.cpu_count(), "B", B, "T", T, flush=True)

t0 = time.perf_counter()
xh = torch.empty((B, 1, T), dtype=torch.float32, pin_memory=True)
yh = torch.empty((B, 1, T), dtype=torch.float32, pin_memory=True)
print(f"pin 2 x {B*T*4/1e9:.2f} GB: {time.perf_counter()-t0:.2f} s", flush=True)
xh.normal_(0, 0.1)
xd = torch.empty((B, 1, T), dtype=torch.float32, device=dev)
yd = torch.zeros((B, 1, T), dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
gb = B * T * 4 / 1e9


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n


t = timed(lambda: xd.copy_(xh, non_blocking=True)); print(f"H2D 1-D  {gb/t:6.1f} GB/s", flush=True)
t = timed(lambda: yh.copy_(yd, non_blocking=True)); print(f"D2H 1-D  {gb/t:6.1f} GB/s", flush=True)


def both():
    with torch.cuda.stream(s1):
        xd.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2):
        yh.copy_(yd, non_blocking=True)


t = timed(both); print(f"H2D+D2H concurrent 1-D  {2*gb/t:6.1f} GB/s total", flush=True)

for C in (4096, 16384, 65536, 262144):
    if C > T:
        continue
    n = T // C
    stage = torch.empty((B, C), dtype=torch.float32, device=dev)

    def h2d_2d():
        for c in range(n):
            stage.copy_(xh[:, 0, c * C:(c + 1) * C], non_blocking=True)

    def d2h_2d():
        for c in range(n):
            yh[:, 0, c * C:(c + 1) * C].copy_(stage, non_blocking=True)
    t = timed(h2d_2d, 2); print(f"H2D 2-D chunks of {C:7d}: {B*n*C*4/1e9/t:6.1f} GB/s", flush=True)
    t = timed(d2h_2d, 2); print(f"D2H 2-D chunks of {C:7d}: {B*n*C*4/1e9/t:6.1f} GB/s", flush=True)
del xd, yd

z = np.load(os.path.join(ROOT, "tests/golden/ckpt_cfg2.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files if k not in ("model_type", "weights_dir")}
m = ntm_b200.RNN(1, 64, 1, False).to(dev)
m.load_state_dict(sd)
m.mode = "f16"
xh.copy_(signals.stream_batch_device(B, T, dev, dur=sec).reshape(B, 1, T))
with torch.inference_mode():
    xd = xh.to(dev)
    t = timed(lambda: m.predict(xd), 2)
    print(f"device-resident predict: {B*T/t/1e9:.3f} Gsamples/s", flush=True)
    del xd
    for chunk in (0, 4096, 16384, 65536, 262144):
        t = timed(lambda: m.predict_host(xh, chunk=chunk, out=yh), 2)
        print(f"predict_host chunk={chunk:7d}: {B*T/t/1e9:.3f} Gsamples/s  ({2*gb/t:.1f} GB/s both ways)", flush=True)
