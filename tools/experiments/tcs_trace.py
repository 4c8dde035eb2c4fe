#!/usr/bin/env python3
"""Dev tool (GPU box, library built with NTM_EXTRA_NVCC_FLAGS=-DNTM_TCS_TRACE): per-step timeline of the two tiles of
CTA 0 of the stream-major kernel.   usage: tcs_trace.py tuning_value [B] [T]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ntm_b200
from ntm_b200 import lib, signals

s = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 37888
T = int(sys.argv[3]) if len(sys.argv) > 3 else 600
dev = torch.device("cuda:0")
z = np.load(os.path.join(ROOT, "tests/golden/ckpt_cfg2.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files if k not in ("model_type", "weights_dir")}
L = lib.load()
L.ntm_set_tuning(s, 4)
with torch.inference_mode():
    x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(sd)
    m.mode = "f16"
    m.initialize_hidden()
    m(x)
    torch.cuda.synchronize()
buf = np.zeros(2 * 256 * 3, dtype=np.int64)
rc = L.ntm_debug_tcs_trace(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
tr = buf.reshape(2, 256, 3)
t0 = tr[0, 0, 0]
E = tr[:, :, 2] - tr[:, :, 0]
M = tr[:, 1:, 0] - tr[:, :-1, 2]
H1 = tr[:, :, 1] - tr[:, :, 0]
cyc = tr[:, 1:, 0] - tr[:, :-1, 0]
lag = tr[1, :, 0] - tr[0, :, 0]
print(f"tuning {s}: B={B} T={T}")
for t in range(2):
    print(f" tile {t}: cycle {cyc[t].mean():7.0f}  E {E[t].mean():7.0f} (first half {H1[t].mean():6.0f})  M+handoff {M[t].mean():6.0f} (min {M[t].min()}, max {M[t].max()})")
print(f" lag tile1-tile0 start: mean {lag.mean():7.0f} min {lag.min()} max {lag.max()}")
print(" first steps (start0, end0, start1, end1 relative):")
for k in range(6):
    print("  ", tr[0, k, 0] - t0, tr[0, k, 2] - t0, tr[1, k, 0] - t0, tr[1, k, 2] - t0)
