#!/bin/bash
# the 8-warp mma.sync form (tuning (8, 5)) against the dispatcher's choice: timing, ESR between them, oracle check
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, ntm_b200
from ntm_b200 import lib, signals
from conftest import load_ckpt
from oracle import c_oracle
dev = "cuda:0"; L = lib.load()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.inference_mode():
    m = ntm_b200.RNN(1, 64, 1, False).to(dev); m.load_state_dict(load_ckpt("cfg2")); m.mode = "f16"
    # correctness: 13 streams x 3000 samples vs the C oracle and vs the dispatcher's kernel
    x = signals.stream_batch(13, 3000)
    xd = torch.from_numpy(x).to(dev).reshape(13, 1, 3000)
    yo, _ = c_oracle.rnn_predict(c_oracle.GruWeights.from_state_dict(load_ckpt("cfg2")), x)
    L.ntm_set_tuning(0, 0); ya = m.predict(xd).cpu().numpy().reshape(13, 3000); ha = m.hidden.clone()
    L.ntm_set_tuning(8, 5); y8 = m.predict(xd).cpu().numpy().reshape(13, 3000); h8 = m.hidden.clone()
    print("kernel", lib.query(lib.Q_LAST_KERNEL), "ESR mma8 vs oracle", c_oracle.esr(y8, yo), "auto vs oracle", c_oracle.esr(ya, yo),
          "max|h8-ha|", float((h8 - ha).abs().max()), flush=True)
    # segmentation: two calls == one call
    m.initialize_hidden(); m.warm_start(); m.hidden = m.hidden.expand(1, 13, 64).contiguous(); h0 = m.hidden.clone()
    y1 = m(xd); m.hidden = h0.clone(); y2 = torch.cat([m(xd[:, :, :1234]), m(xd[:, :, 1234:])], 2)
    print("segmentation exact", bool(torch.equal(y1, y2)), flush=True)
    m.initialize_hidden(); m.warm_start(); hw = m.hidden.clone()
    for B, T in ((592, 48000), (1024, 48000), (1184, 48000)):
        xb = signals.stream_batch_device(B, T, dev, dur=10.0).reshape(B, 1, T)
        for tune in ((0, 0), (8, 5)):
            L.ntm_set_tuning(*tune)
            m.hidden = hw.expand(1, B, 64).contiguous(); m(xb[:, :, :1000])
            best = 1e9
            for _ in range(3):
                m.hidden = hw.expand(1, B, 64).contiguous()
                e0.record(); y = m(xb); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"B={B} tune={tune}: {best*1e6/T:7.1f} ns/step ({B*T/best/1e6:6.3f} Gs/s) kernel {lib.query(lib.Q_LAST_KERNEL)}", flush=True)
    L.ntm_set_tuning(0, 0)
PY
