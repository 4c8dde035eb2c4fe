"""The oracle against the committed golden vectors (outputs of the real reference, oracle/make_golden.py).
CPU only.  Tolerances: the C restatement sums in a different order than MKL, so it is compared at float32
round-off (measured 3e-7 .. 2e-6 on these signals); the delay line and the torch restatement are bit-exact."""
import numpy as np
import pytest
import torch

from conftest import SIGNALS, load_ckpt, load_golden
from oracle import c_oracle, ref_torch


@pytest.mark.parametrize("tag", ["cfg1", "cfg2"])
def test_c_oracle_rnn_vs_reference(tag):
    w = c_oracle.GruWeights.from_state_dict(load_ckpt(tag))
    g = load_golden(f"golden_{tag}")
    _, h1 = c_oracle.gru_forward(w, np.zeros((1, 1024), np.float32))
    assert np.max(np.abs(h1.reshape(-1) - g["h_warm"])) < 2e-6
    for sig in SIGNALS:
        y, _ = c_oracle.rnn_predict(w, g[f"x_{sig}"].reshape(1, -1))
        assert np.max(np.abs(y.reshape(-1) - g[f"y_{sig}"])) < 5e-6, sig
        y64, _ = c_oracle.rnn_predict(w, g[f"x_{sig}"].reshape(1, -1), f64=True)
        assert np.max(np.abs(y64.reshape(-1) - g[f"y64_{sig}"])) < 1e-9, sig


def test_c_oracle_skip():
    w = c_oracle.GruWeights.from_state_dict(load_ckpt("cfg1"))
    g = load_golden("golden_cfg1")
    y, _ = c_oracle.rnn_predict(w, g["x_noise"].reshape(1, -1), skip=True)
    assert np.max(np.abs(y.reshape(-1) - g["y_skip_noise"])) < 5e-6


def test_c_oracle_diffdel_vs_reference():
    w = c_oracle.GruWeights.from_state_dict(load_ckpt("cfg3"))
    g = load_golden("golden_cfg3")
    assert w.b_out is None
    for sig in SIGNALS:
        y, pre, _, hist = c_oracle.diffdel_predict(w, g[f"x_{sig}"].reshape(1, -1), g[f"d_{sig}"].reshape(1, -1),
                                                   int(g["max_delay"]))
        assert np.max(np.abs(pre.reshape(-1) - g[f"pre_{sig}"])) < 5e-6, sig
        assert np.max(np.abs(y.reshape(-1) - g[f"y_{sig}"])) < 5e-6, sig
        assert np.max(np.abs(hist.reshape(-1) - g[f"hist_{sig}"])) < 5e-6, sig
        # the delay stage alone is bit-exact given the reference's pre_d
        yd, hd = c_oracle.delay_forward(g[f"pre_{sig}"].reshape(1, -1), g[f"d_{sig}"].reshape(1, -1),
                                        g["hist_warm"].reshape(1, -1))
        assert np.array_equal(yd.reshape(-1), g[f"y_{sig}"]), sig
        assert np.array_equal(hd.reshape(-1), g[f"hist_{sig}"]), sig


@pytest.mark.parametrize("case", ["frac", "integer", "short_T", "edge", "negfrac"])
def test_c_oracle_delay_cases_bit_exact(case):
    g = load_golden("golden_delay")
    x, d, h0 = g[f"{case}_x"][:, 0], g[f"{case}_d"][:, 0], g[f"{case}_hist0"][:, 0]
    for window in (False, True):
        y, h1 = c_oracle.delay_forward(x, d, h0, window=window)
        assert np.array_equal(y, g[f"{case}_y"][:, 0])
        assert np.array_equal(h1, g[f"{case}_hist1"][:, 0])
    yw, h2 = c_oracle.delay_forward(x, d, g[f"{case}_hist1"][:, 0], warmup=True)
    assert np.array_equal(yw, g[f"{case}_ywarm"][:, 0]) and np.array_equal(h2, g[f"{case}_hist2"][:, 0])


def test_c_oracle_delay_assert():
    with pytest.raises(AssertionError):
        c_oracle.delay_forward(np.zeros((1, 4), np.float32), np.full((1, 4), 9.0, np.float32),
                               np.zeros((1, 8), np.float32))


def test_torch_restatement_vs_reference():
    """oracle/ref_torch.py is the same torch arithmetic the reference calls; bit-exact on the machine that made
    the fixtures, float32 round-off elsewhere (MKL code path depends on the CPU)."""
    torch.set_num_threads(1)
    for tag in ("cfg1", "cfg2"):
        net = ref_torch.RefNet(load_ckpt(tag))
        g = load_golden(f"golden_{tag}")
        for sig in ("sweepnoise", "pulse"):
            y, _ = net.predict(torch.from_numpy(g[f"x_{sig}"]).reshape(1, 1, -1))
            assert np.max(np.abs(y.numpy().reshape(-1) - g[f"y_{sig}"])) < 5e-6
    net = ref_torch.RefNet(load_ckpt("cfg3"))
    g = load_golden("golden_cfg3")
    y, pre, _, buf = ref_torch.diffdel_predict(net, torch.from_numpy(g["x_noise"]).reshape(1, 1, -1),
                                               torch.from_numpy(g["d_noise"]).reshape(1, 1, -1), int(g["max_delay"]))
    assert np.max(np.abs(y.numpy().reshape(-1) - g["y_noise"])) < 5e-6
    assert np.max(np.abs(pre.numpy().reshape(-1) - g["pre_noise"])) < 5e-6


def test_oracle_segmentation_invariance():
    """Reference behaviour 9.3#1: predict == one long forward == 64-sample blocks, bit-identical."""
    w = c_oracle.GruWeights.from_state_dict(load_ckpt("cfg1"))
    x = load_golden("golden_cfg1")["x_noise"][:1000].reshape(1, -1)
    y_all, h_all = c_oracle.gru_forward(w, x)
    h = None
    ys = []
    for s in range(0, 1000, 64):
        y, h = c_oracle.gru_forward(w, x[:, s:s + 64], h)
        ys.append(y)
    assert np.array_equal(np.concatenate(ys, 1), y_all) and np.array_equal(h, h_all)


def test_esr():
    a = np.random.default_rng(0).standard_normal(1000).astype(np.float32)
    assert c_oracle.esr(a, a) == 0.0
    e = c_oracle.esr(a * 0.9, a)
    ref = np.mean((a - 0.9 * a).astype(np.float64) ** 2) / (np.mean(a.astype(np.float64) ** 2) + 1e-5)
    assert abs(e - ref) < 1e-12


def test_known_answers_baseline_md():
    """BASELINE.md section 2 known answers: hidden state after warm_start()."""
    for tag, s, n in (("cfg1", -0.149917979, 2.213100609), ("cfg2", -4.187483256, 2.532229933),
                      ("cfg3", -0.782204804, 2.725984538)):
        h = load_golden(f"golden_{tag}")["h_warm"].astype(np.float64)
        assert abs(h.sum() - s) < 1e-5 and abs(np.linalg.norm(h) - n) < 1e-5


def test_dcpre_esr_oracle_vs_reference_golden():
    """The C restatement of ESR / DCPreESR (ntm_oracle_dcpre_esr) against values the reference's own loss classes
    produced (oracle/make_golden_loss.py: GreyBoxDRC.loss_funcs.ESRLoss(dc_pre=True), CoreAudioML ESRLoss)."""
    import numpy as np
    g = load_golden("golden_loss")
    assert np.array_equal(c_oracle.dcpre_taps(), g["taps"])
    assert g["taps"][0] == 1.0 and abs(g["taps"][1] + 0.005) < 1e-9 and len(g["taps"]) == 2000
    for case in ("sweep_vs_perturbed", "dc_offset", "batch5", "short", "silence_target"):
        dc, num, den = c_oracle.dcpre_esr(g[f"o_{case}"], g[f"t_{case}"], True)
        pl, _, _ = c_oracle.dcpre_esr(g[f"o_{case}"], g[f"t_{case}"], False)
        assert abs(dc - float(g[f"dcpre_{case}"])) <= 2e-5 * abs(float(g[f"dcpre_{case}"])) + 1e-9, case
        assert abs(pl - float(g[f"esr_{case}"])) <= 2e-5 * abs(float(g[f"esr_{case}"])) + 1e-9, case
        n = g[f"o_{case}"].size
        assert abs(dc - (num / n) / (den / n + 1e-5)) < 1e-12


def _best12():
    g = load_golden("golden_best12")
    for i in range(int(g["n"])):
        pre = f"w{i}_"
        sd = {k[len(pre):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre)}
        yield i, str(g[f"kind{i}"]), sd, g


def test_c_oracle_on_all_12_shipped_best_checkpoints():
    """Every `weights/*_BEST/best.pth` (6 GRU, 6 DiffDelGRU): warm-start known answer and predict() on two signals vs the
    reference's own outputs (oracle/make_golden_best.py).  Tolerance: float32 round-off, or 4x the reference's own
    fp32-vs-fp64 floor where the checkpoint is chaotic on the signal."""
    n = 0
    for i, kind, sd, g in _best12():
        w = c_oracle.GruWeights.from_state_dict(sd)
        assert (w.b_out is None) == (kind == "DiffDelGRU")
        _, h1 = c_oracle.gru_forward(w, np.zeros((1, 1024), np.float32))
        assert np.max(np.abs(h1.reshape(-1) - g[f"h_warm{i}"])) < 5e-6, i
        for sig in g["signals"]:
            x, tol = g[f"x_{sig}"].reshape(1, -1), max(5e-6, 4.0 * float(g[f"floor{i}_{sig}"]))
            if kind == "GRU":
                y, _ = c_oracle.rnn_predict(w, x)
            else:
                y, pre, _, _ = c_oracle.diffdel_predict(w, x, g[f"d_{sig}"].reshape(1, -1), int(g["max_delay"]))
                assert np.max(np.abs(pre.reshape(-1) - g[f"pre{i}_{sig}"])) < tol, (i, sig)
            assert np.max(np.abs(y.reshape(-1) - g[f"y{i}_{sig}"])) < tol, (i, sig)
        n += 1
    assert n == 12
