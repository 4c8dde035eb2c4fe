#!/usr/bin/env python3
"""Generate tests/golden/golden_loss.npz from the REAL reference loss classes (build container only):
    code/GreyBoxDRC/loss_funcs.py ESRLoss(dc_pre=True)      ("DCPreESR" of code/test-model.py:252)
    code/Automated_GuitarAmpModelling/CoreAudioML/training.py ESRLoss  ("ESR" of code/test-model.py:251)
evaluated with torch on CPU (fp32) on pairs (output, target) built from the golden GRU fixtures, and asserts that the
C restatement (oracle/ntm_oracle.c: ntm_oracle_dcpre_esr) agrees with them to the reference's own fp32 round-off."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "code"))
from GreyBoxDRC.loss_funcs import ESRLoss as DCPreESR  # noqa: E402  (the reference)
from Automated_GuitarAmpModelling.CoreAudioML.training import ESRLoss  # noqa: E402  (the reference)

from oracle import c_oracle  # noqa: E402

torch.set_num_threads(1)
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    rng = np.random.default_rng(7)
    g2 = np.load(os.path.join(GOLD, "golden_cfg2.npz"))
    g1 = np.load(os.path.join(GOLD, "golden_cfg1.npz"))
    cases = {}
    # (output, target): model output vs a perturbed "tape" target, a DC-offset pair, a ragged multi-stream batch,
    # a signal shorter than the filter, silence
    y = g2["y_sweepnoise"].astype(np.float32)
    cases["sweep_vs_perturbed"] = (y[None], (0.9 * y + 0.01 * rng.standard_normal(y.shape)).astype(np.float32)[None])
    y = g1["y_noise"].astype(np.float32)
    cases["dc_offset"] = ((y + 0.05)[None].astype(np.float32), (0.8 * y - 0.02)[None].astype(np.float32))
    B, T = 5, 6000
    o = (0.3 * rng.standard_normal((B, T))).astype(np.float32)
    t = (o + 0.05 * rng.standard_normal((B, T)) + 0.01).astype(np.float32)
    cases["batch5"] = (o, t)
    cases["short"] = ((0.2 * rng.standard_normal((2, 700))).astype(np.float32),
                      (0.2 * rng.standard_normal((2, 700))).astype(np.float32))
    cases["silence_target"] = ((1e-3 * rng.standard_normal((1, 4096))).astype(np.float32), np.zeros((1, 4096), np.float32))
    out = {"taps": c_oracle.dcpre_taps()}
    ref_dc, ref_pl = DCPreESR(dc_pre=True), ESRLoss()
    assert np.array_equal(out["taps"], torch.flipud(ref_dc.dc_pre.pars.reshape(-1)).numpy()), "taps differ from the reference"
    for name, (o, t) in cases.items():
        ot, tt = torch.from_numpy(o).unsqueeze(1), torch.from_numpy(t).unsqueeze(1)       # (B, 1, T) as test-model.py
        l_dc = float(ref_dc(ot, tt))
        l_pl = float(ref_pl(ot, tt))
        c_dc, _, _ = c_oracle.dcpre_esr(o, t, True)
        c_pl, _, _ = c_oracle.dcpre_esr(o, t, False)
        print(f"{name:20s} DCPreESR ref {l_dc:.8e} oracle {c_dc:.8e} | ESR ref {l_pl:.8e} oracle {c_pl:.8e}")
        assert abs(c_dc - l_dc) <= 2e-5 * abs(l_dc) + 1e-9, name
        assert abs(c_pl - l_pl) <= 2e-5 * abs(l_pl) + 1e-9, name
        out[f"o_{name}"], out[f"t_{name}"] = o, t
        out[f"dcpre_{name}"], out[f"esr_{name}"] = np.float64(l_dc), np.float64(l_pl)
    np.savez_compressed(os.path.join(GOLD, "golden_loss.npz"), **out)
    print("wrote tests/golden/golden_loss.npz")


if __name__ == "__main__":
    main()
