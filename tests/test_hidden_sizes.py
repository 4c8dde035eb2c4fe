"""Hidden sizes below 64 (the reference builds RNN / DiffDelRNN at any width: default 8, code/model.py:22; code/train.py:50
defaults to 16; scripts/sbatch-train.sh:15 trains 32).  The engine embeds them, zero-padded, in its 64-unit kernels.
Goldens: outputs of the reference's own classes on CPU (oracle/make_golden_hs.py).

Tolerances (BASELINE.json north_star): fp32 and the strict tensor-core mode max-abs <= 1e-5 against the reference's fp32 output;
f16 operands ESR <= 1e-4."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import c_oracle

DEV = "cuda:0"
G = load_golden("golden_hs")
CASES = list(range(int(G["n"])))
SIGS = [str(s) for s in G["signals"]]


def _sd(i):
    pre = f"w{i}_"
    return {k[len(pre):]: torch.from_numpy(G[k]) for k in G.files if k.startswith(pre)}


@pytest.mark.parametrize("i", CASES)
def test_c_oracle_small_hidden_vs_reference(i):
    """The oracle itself at these widths (CPU)."""
    w = c_oracle.GruWeights.from_state_dict(_sd(i))
    assert w.H == int(G[f"H{i}"])
    for sig in SIGS:
        x = G[f"x_{sig}"].reshape(1, -1)
        if str(G[f"kind{i}"]) == "GRU":
            y, h = c_oracle.rnn_predict(w, x, skip=bool(G[f"skip{i}"]))
        else:
            y, pre, h, hist = c_oracle.diffdel_predict(w, x, G[f"d_{sig}"].reshape(1, -1), int(G["max_delay"]))
            assert np.max(np.abs(pre.reshape(-1) - G[f"pre{i}_{sig}"])) < 5e-6
            assert np.max(np.abs(hist.reshape(-1) - G[f"hist{i}_{sig}"])) < 5e-6
        assert np.max(np.abs(y.reshape(-1) - G[f"y{i}_{sig}"])) < 5e-6, sig
        assert np.max(np.abs(h.reshape(-1) - G[f"h{i}_{sig}"])) < 5e-6, sig


def _model(i, mode):
    from ntm_b200 import DiffDelRNN, RNN
    H, skip = int(G[f"H{i}"]), bool(G[f"skip{i}"])
    if str(G[f"kind{i}"]) == "GRU":
        m = RNN(input_size=1, hidden_size=H, output_size=1, skip=skip)
    else:
        m = DiffDelRNN(input_size=1, hidden_size=H, output_size=1, skip=skip, max_delay=int(G["max_delay"]))
    m.load_state_dict(_sd(i))                  # strict: same keys and shapes as the reference's state_dict
    m = m.to(DEV)
    m.mode = mode
    return m


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV).reshape(1, 1, -1)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "f16x3", "f16"])
@pytest.mark.parametrize("i", CASES)
def test_small_hidden_predict_vs_reference(i, mode):
    m = _model(i, mode)
    H = int(G[f"H{i}"])
    gru = str(G[f"kind{i}"]) == "GRU"
    with torch.inference_mode():
        for sig in SIGS:
            x = _dev(G[f"x_{sig}"])
            if gru:
                outs = {"y": m.predict(x)}
            else:
                y, pre = m.predict(x, _dev(G[f"d_{sig}"]))
                outs = {"y": y, "pre": pre}
            assert tuple(m.hidden.shape) == (1, 1, H)             # the reference's state shape, not the engine's 64
            for name, got in outs.items():
                got = got.cpu().numpy().reshape(-1)
                ref = G[f"{name}{i}_{sig}"]
                if mode == "f16":
                    assert c_oracle.esr(got, ref) <= 1e-4, (sig, name, c_oracle.esr(got, ref))
                else:
                    assert np.max(np.abs(got - ref)) <= 1e-5, (sig, name, float(np.max(np.abs(got - ref))))
            if mode != "f16":
                assert np.max(np.abs(m.hidden.cpu().numpy().reshape(-1) - G[f"h{i}_{sig}"])) <= 1e-5
                if not gru:
                    assert np.max(np.abs(m.diffdel.buffer.cpu().numpy().reshape(-1) - G[f"hist{i}_{sig}"])) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "f16"])
def test_small_hidden_carried_state_and_batches(mode):
    """forward() in ragged segments with self.hidden carried == one call, bit for bit, for a batch; the block-stream and host
    pipelines keep the reference's state shape; every stream of a batch equals the same stream run alone."""
    from ntm_b200 import signals
    i = 1                                             # GRU-HS[16]
    m = _model(i, mode)
    H, B, T = int(G[f"H{i}"]), 5, 3000
    x = torch.from_numpy(signals.stream_batch(B, T)).to(DEV).reshape(B, 1, T)
    with torch.inference_mode():
        y_all = m.predict(x)
        h_all = m.hidden.clone()
        assert tuple(h_all.shape) == (1, B, H)
        m.predict(x[:, :, :0])
        parts, s = [], 0
        for n in (64, 1, 33, 2048, 854):
            parts.append(m(x[:, :, s:s + n]))
            s += n
        assert s == T and torch.equal(torch.cat(parts, 2), y_all) and torch.equal(m.hidden, h_all)
        assert torch.equal(m.predict(x[2:3]), y_all[2:3])
        # block stream: state handed back in the reference's shape
        m.predict(x[:, :, :0])
        bs = m.block_stream(B, 64)
        yb = torch.cat([bs.process(x[:, :, k:k + 64]).clone() for k in range(0, 640, 64)], 2)
        assert tuple(bs.close().shape) == (1, B, H) and torch.equal(yb, y_all[:, :, :640])
        # host pipeline
        yh = m.predict_host(x.cpu())
        assert tuple(m.hidden.shape) == (1, B, H) and torch.equal(yh, y_all.cpu()) and torch.equal(m.hidden, h_all)


@pytest.mark.gpu
def test_hidden_size_above_64_is_rejected():
    from ntm_b200 import RNN
    m = RNN(1, 65, 1, False).to(DEV)
    with pytest.raises(RuntimeError, match="hidden_size"):
        m(torch.zeros(1, 1, 8, device=DEV))


@pytest.mark.gpu
def test_default_constructed_models_run():
    """`RNN()` / `DiffDelRNN()` exactly as the reference's signature defaults build them (hidden_size 8)."""
    from ntm_b200 import DiffDelRNN, RNN
    with torch.inference_mode():
        m = RNN().to(DEV)
        y = m.predict(torch.randn(2, 1, 500, device=DEV) * 0.1)
        assert tuple(y.shape) == (2, 1, 500) and tuple(m.hidden.shape) == (1, 2, 8) and bool(torch.isfinite(y).all())
        md = DiffDelRNN(max_delay=32).to(DEV)
        y, pre = md.predict(torch.randn(1, 1, 500, device=DEV) * 0.1, torch.full((1, 1, 500), 7.5, device=DEV))
        assert tuple(md.hidden.shape) == (1, 1, 8) and bool(torch.isfinite(y).all())
        # delayed by 7.5 samples: the mean of two neighbours of pre_d
        assert torch.allclose(y[0, 0, 100:], 0.5 * (pre[0, 0, 93:-7] + pre[0, 0, 92:-8]), atol=1e-6)
