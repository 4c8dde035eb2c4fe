"""On-device evaluation losses (csrc/esr.cu) through the C ABI: ESR and DCPreESR against the golden values produced by the
reference's own loss classes (oracle/make_golden_loss.py) and against the C restatement on fresh inputs.
Tolerance: floating point -- the reference evaluates the 2000-tap filter as an fp32 conv1d (relative round-off ~1e-6 of
the loss); the engine carries the filter in double.  Bound: |loss - ref| <= 2e-5 * |ref| + 1e-9."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from ntm_b200 import DCPreESR, ESRLoss, lib
from oracle import c_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = ("sweep_vs_perturbed", "dc_offset", "batch5", "short", "silence_target")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def close(a, b):
    return abs(a - b) <= 2e-5 * abs(b) + 1e-9


@pytest.mark.parametrize("case", CASES)
def test_losses_vs_reference_golden(case):
    g = load_golden("golden_loss")
    o, t = dev(g[f"o_{case}"]).unsqueeze(1), dev(g[f"t_{case}"]).unsqueeze(1)       # (B, 1, T) as code/test-model.py
    n0 = lib.query(lib.Q_KERNEL_LAUNCHES)
    l_dc = float(DCPreESR(dc_pre=True)(o, t))
    l_pl = float(ESRLoss()(o, t))
    assert lib.query(lib.Q_KERNEL_LAUNCHES) == n0 + 2
    assert close(l_dc, float(g[f"dcpre_{case}"])), (case, l_dc, float(g[f"dcpre_{case}"]))
    assert close(l_pl, float(g[f"esr_{case}"])), (case, l_pl, float(g[f"esr_{case}"]))
    assert close(float(DCPreESR(dc_pre=False)(o, t)), float(g[f"esr_{case}"]))


def test_losses_vs_oracle_ragged_and_strided():
    """Chunk boundaries (T not a multiple of the 128-sample tile or of the chunk), many streams, row strides, the
    INIT_LEN cut of code/test-model.py:367-370 (a sliced view), (B, T) layout."""
    rng = np.random.default_rng(3)
    for B, T in ((3, 40001), (64, 5000), (1, 1999), (1, 2000), (2, 2001), (7, 129), (1, 1)):
        t = (0.2 * rng.standard_normal((B, T)) + 0.05).astype(np.float32)
        o = (t + 0.03 * rng.standard_normal((B, T))).astype(np.float32)
        for dc in (True, False):
            ref, _, _ = c_oracle.dcpre_esr(o, t, dc)
            got = float(ESRLoss(dc_pre=dc)(dev(o), dev(t)))
            assert close(got, ref), (B, T, dc, got, ref)
    B, T, cut = 4, 30000, 1024
    t = (0.2 * rng.standard_normal((B, T))).astype(np.float32)
    o = (t + 0.03 * rng.standard_normal((B, T))).astype(np.float32)
    ref, _, _ = c_oracle.dcpre_esr(o[:, cut:], t[:, cut:], True)
    od, td = dev(o).unsqueeze(1), dev(t).unsqueeze(1)
    assert close(float(DCPreESR()(od[:, :, cut:], td[:, :, cut:])), ref)


def test_loss_properties_at_full_width():
    """cfg-2 width (1024 streams x 1 s): size-independent properties -- loss(t, t) = 0, loss(0, t) = 1 up to epsilon,
    scale invariance, and the DC filter removes a constant offset that plain ESR sees."""
    B, T = 1024, 48000
    g = torch.Generator(device=DEV).manual_seed(1)
    t = 0.2 * torch.randn(B, 1, T, device=DEV, generator=g)
    o = t + 0.02 * torch.randn(B, 1, T, device=DEV, generator=g)
    for L in (ESRLoss(), DCPreESR()):
        assert float(L(t, t)) == 0.0
        assert abs(float(L(torch.zeros_like(t), t)) - 1.0) < 1e-3
        base = float(L(o, t))
        assert abs(base - 0.01) < 1e-3
        assert abs(float(L(2 * o, 2 * t)) - base) < 1e-3 * base
    # a constant offset: plain ESR sees it in full (0.1^2 / 0.2^2 = 0.25); the pre-emphasis filter only lets the
    # switch-on transient through: sum_n (0.1 R^n)^2 ~ 1.0 per stream against an error energy of 0.02^2 * T = 19.2 (+5 %)
    off = o + 0.1
    assert float(ESRLoss()(off, t)) > 20 * float(ESRLoss()(o, t))
    rel = (float(DCPreESR()(off, t)) - float(DCPreESR()(o, t))) / float(DCPreESR()(o, t))
    assert 0.03 < rel < 0.08, rel


def test_loss_rejects_bad_arguments():
    t = torch.zeros(2, 1, 16, device=DEV)
    with pytest.raises(RuntimeError, match="CPU"):
        ESRLoss()(t.cpu(), t.cpu())
    with pytest.raises(RuntimeError, match="differ"):
        ESRLoss()(t, t[:, :, :8])
    with pytest.raises(RuntimeError, match=r"\(B, 1, T\)"):
        ESRLoss()(torch.zeros(2, 2, 16, device=DEV), torch.zeros(2, 2, 16, device=DEV))
