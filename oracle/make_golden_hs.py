#!/usr/bin/env python3
"""Golden vectors for hidden sizes below 64, from the REAL reference (imports /root/reference/code/model.py).

The reference builds `RNN(hidden_size=...)` at any width: its default is 8 (code/model.py:22), code/train.py:50 defaults to 16,
scripts/sbatch-train.sh:15 trains 32.  No such checkpoint ships, so the weights here are the reference classes' own
initialisation under a fixed seed (scaled up so that the gates leave their linear range), run through the reference's
`predict` on CPU in fp32.  Run in the build container only:
    python oracle/make_golden_hs.py           -> tests/golden/golden_hs.npz
Per case i: `kind{i}`, `H{i}`, `skip{i}`, the state_dict tensors `w{i}_*`, and per signal the input, the reference output(s), the
reference's final `hidden` and its own fp32-vs-fp64 floor.  Asserts oracle/ref_torch.py bit-identical to the reference and
oracle/ntm_oracle.c equal to round-off at these widths too.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
for name in ("soundfile", "librosa", "librosa.filters"):     # absent here; only plotting/IO helpers use them
    m = types.ModuleType(name)
    if name == "librosa.filters":
        m.mel = None
    sys.modules.setdefault(name, m)
sys.path.insert(0, os.path.join(REF, "code"))
import model as refmodel  # noqa: E402  (the reference)

import ntm_b200  # noqa: E402,F401
from ntm_b200 import signals  # noqa: E402
from oracle import c_oracle, ref_torch  # noqa: E402

torch.set_num_threads(1)
GOLD = os.path.join(ROOT, "tests", "golden")
T = 6000
SIGNALS = ("sweepnoise", "pulse", "sine1k")
MAX_DELAY = 40
# (kind, hidden size, skip)
CASES = [("GRU", 8, False), ("GRU", 16, False), ("GRU", 32, True), ("GRU", 1, False), ("GRU", 63, False),
         ("DiffDelGRU", 16, False), ("DiffDelGRU", 32, False)]


def main():
    out = {"n": len(CASES), "signals": np.array(SIGNALS), "max_delay": MAX_DELAY, "T": T}
    for i, (kind, H, skip) in enumerate(CASES):
        torch.manual_seed(100 + i)
        if kind == "GRU":
            m = refmodel.RNN(1, H, 1, skip)
        else:
            m = refmodel.DiffDelRNN(1, H, 1, skip, max_delay=MAX_DELAY)
        with torch.no_grad():                       # livelier than the +-1/sqrt(H) initialisation, still stable
            m.GRU.weight_hh_l0.mul_(2.5)
            m.GRU.weight_ih_l0.mul_(3.0)
            m.output.weight.mul_(1.5)
        m.eval()
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        out[f"kind{i}"], out[f"H{i}"], out[f"skip{i}"] = kind, H, int(skip)
        for k, v in sd.items():
            out[f"w{i}_{k}"] = v.numpy()
        w = c_oracle.GruWeights.from_state_dict(sd)
        net, net64 = ref_torch.RefNet(sd), ref_torch.RefNet(sd, torch.float64)
        with torch.inference_mode():
            for j, sig in enumerate(SIGNALS):
                x = signals.signal(sig, T, seed=j)
                xt = torch.from_numpy(x).reshape(1, 1, -1)
                if kind == "GRU":
                    y = m.predict(xt).numpy().reshape(-1)
                    yp, _ = net.predict(xt, skip=skip)
                    assert np.array_equal(y, yp.numpy().reshape(-1)), "ref_torch != reference"
                    y64 = net64.predict(xt, skip=skip)[0].numpy().reshape(-1)
                    yc, _ = c_oracle.rnn_predict(w, x.reshape(1, -1), skip=skip)
                    floor = float(np.max(np.abs(y - y64)))
                    print(f"GRU H={H:2d} {sig:10s} floor={floor:.2e} C-f32 vs ref={np.max(np.abs(yc.reshape(-1) - y)):.2e}")
                    assert np.max(np.abs(yc.reshape(-1) - y)) <= 5e-6 + 4 * floor
                    out[f"y{i}_{sig}"] = y
                else:
                    d = (signals.delay_trajectory(1, T, first_stream=j)[0] * (MAX_DELAY / signals.DELAY_MAX) * 0.9).astype(np.float32)
                    dtt = torch.from_numpy(d).reshape(1, 1, -1)
                    y, pre = m.predict(xt, dtt)
                    y, pre = y.numpy().reshape(-1), pre.numpy().reshape(-1)
                    yp, prep, _, _ = ref_torch.diffdel_predict(net, xt, dtt, MAX_DELAY)
                    assert np.array_equal(y, yp.numpy().reshape(-1)) and np.array_equal(pre, prep.numpy().reshape(-1))
                    pre64 = net64.predict(xt)[0].numpy().reshape(-1)
                    yc, prec, _, _ = c_oracle.diffdel_predict(w, x.reshape(1, -1), d.reshape(1, -1), MAX_DELAY)
                    floor = float(np.max(np.abs(pre - pre64)))
                    print(f"DiffDelGRU H={H:2d} {sig:10s} floor={floor:.2e} C-f32 vs ref: pre_d "
                          f"{np.max(np.abs(prec.reshape(-1) - pre)):.2e} y {np.max(np.abs(yc.reshape(-1) - y)):.2e}")
                    assert np.max(np.abs(prec.reshape(-1) - pre)) <= 5e-6 + 4 * floor
                    out[f"d_{sig}"] = d
                    out[f"y{i}_{sig}"] = y
                    out[f"pre{i}_{sig}"] = pre
                    out[f"hist{i}_{sig}"] = m.diffdel.buffer.numpy().reshape(-1).copy()
                out[f"x_{sig}"] = x
                out[f"h{i}_{sig}"] = m.hidden.numpy().reshape(-1).copy()
                out[f"floor{i}_{sig}"] = floor
    np.savez_compressed(os.path.join(GOLD, "golden_hs.npz"), **out)
    print("written", os.path.join(GOLD, "golden_hs.npz"))


if __name__ == "__main__":
    main()
