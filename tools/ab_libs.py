#!/usr/bin/env python3
"""Dev tool (GPU box): A/B builds of libntm_b200 (NTM_B200_LIB) on the mma.sync kernel: timing per width, output hash.
usage: NTM_B200_LIB=path ab_libs.py [mode]"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import ntm_b200
from ntm_b200 import lib, signals
from conftest import load_ckpt

dev = "cuda:0"
L = lib.load()
mode = sys.argv[1] if len(sys.argv) > 1 else "f16"
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.inference_mode():
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg2"))
    m.mode = mode
    m.initialize_hidden(); m.warm_start()
    hw = m.hidden.clone()
    for B, T in ((1, 200000), (256, 48000), (592, 48000), (1024, 48000), (2048, 24000), (8192, 12000)):
        x = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
        for tune in ((8, 3), (4, 3)):
            L.ntm_set_tuning(*tune)
            m.hidden = hw.expand(1, B, 64).contiguous(); m(x[:, :, :1000])
            best = 1e9
            for _ in range(3):
                m.hidden = hw.expand(1, B, 64).contiguous()
                e0.record(); y = m(x); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            hsh = hashlib.sha1(y.cpu().numpy().tobytes()).hexdigest()[:10]
            print(f"{os.path.basename(lib.LIB_PATH)} {mode} B={B:5d} T={T} tune={tune}: {best*1e6/T:7.1f} ns/step "
                  f"({B*T/best/1e6:6.3f} Gs/s) sha={hsh}", flush=True)
L.ntm_set_tuning(0, 0)
