#!/usr/bin/env python3
"""Dev tool (GPU box): the resident real-time server (ntm_rt_*) -- equality with one long forward call, latency."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import ntm_b200
from ntm_b200 import signals
from conftest import load_ckpt

dev = "cuda:0"
with torch.inference_mode():
    m = ntm_b200.RNN(1, 64, 1, False).to(dev)
    m.load_state_dict(load_ckpt("cfg1"))
    m.mode = "f16"
    for B, T, nblk in ((1, 64, 400), (3, 100, 50), (4, 256, 20), (1, 1, 300)):
        xh = torch.from_numpy(signals.stream_batch(B, T * nblk)).contiguous()
        m.initialize_hidden(); m.warm_start()
        hw = m.hidden.expand(1, B, 64).contiguous()
        m.hidden = hw.clone()
        yref = m(xh.to(dev).reshape(B, 1, -1)).cpu().reshape(B, -1)
        href = m.hidden.cpu()
        torch.cuda.synchronize()
        m.hidden = hw.clone()
        rt = m.realtime_stream(B, T)
        blocks = [xh[:, k * T:(k + 1) * T].contiguous() for k in range(nblk)]
        out, lat = [], []
        for k in range(nblk):
            t0 = time.perf_counter()
            y = rt.process(blocks[k])
            lat.append(time.perf_counter() - t0)
            out.append(y.clone())
        lat2 = []
        for k in range(min(nblk, 200)):                      # step(): fixed input block, pre-resolved pointers
            t0 = time.perf_counter()
            rt.step()
            lat2.append(time.perf_counter() - t0)
        lat2.sort()
        m2 = ntm_b200.RNN(1, 64, 1, False).to(dev); m2.load_state_dict(load_ckpt("cfg1")); m2.mode = "f16"
        yrt = torch.cat(out, 1)
        print(f"   step(): median {lat2[len(lat2)//2]*1e6:.1f} us", flush=True)
        rt.x_in.zero_()
        h = None
        lat = sorted(lat[5:])
        rt2 = None
        rt.close()
        # state check on a fresh stream (the step() calls above advanced the first one)
        m.hidden = hw.clone()
        rt2 = m.realtime_stream(B, T)
        for k in range(nblk):
            rt2.process(blocks[k])
        h = rt2.close().cpu()
        print(f"B={B} T={T}: y identical {bool(torch.equal(yrt, yref))} (max|d| {float((yrt - yref).abs().max()):.1e}), state identical "
              f"{bool(torch.equal(h, href))}; block latency median {lat[len(lat)//2]*1e6:.1f} us, p99 {lat[int(len(lat)*0.99)]*1e6:.1f} us, "
              f"{lat[len(lat)//2]*1e9/T:.0f} ns/sample", flush=True)
    # idle timeout: the server leaves by itself, process() reports it
    m.initialize_hidden()
    rt = m.realtime_stream(1, 64, idle_timeout_ms=300)
    rt.process(torch.zeros(1, 64))
    time.sleep(1.0)
    try:
        rt.process(torch.zeros(1, 64))
        print("idle timeout: NOT detected")
    except RuntimeError as e:
        print("idle timeout detected:", e)
    rt.close()
    torch.cuda.synchronize()
    print("done")
