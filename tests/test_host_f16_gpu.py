"""The opt-in 16-bit host transport (ntm_gru_predict_host_f16, include/ntm_b200.h): binary16 samples over the host link, the same
fp32 kernels in between.  Checked against the float32 transport of the same engine mode and against the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_ckpt
from ntm_b200 import RNN, signals
from oracle import c_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def esr(y, t):
    y, t = y.astype(np.float64), t.astype(np.float64)
    return float(((y - t) ** 2).mean() / ((t ** 2).mean() + 1e-5))


@pytest.mark.parametrize("mode", ["fp32", "f16"])
@pytest.mark.parametrize("B,T,chunk", [(5, 3001, 0), (33, 9000, 2048), (3, 48000, 4096)])
def test_f16_transport_matches_f32_transport(mode, B, T, chunk):
    sd = load_ckpt("cfg2")
    m = RNN(1, 64, 1, False).to(DEV)
    m.load_state_dict(sd)
    m.mode = mode
    x = signals.stream_batch(B, T)
    x16 = torch.from_numpy(x).to(torch.float16).reshape(B, 1, T)
    with torch.inference_mode():
        # the float32 transport fed with the SAME (binary16-representable) samples: what the engine computes must be identical,
        # only the rounding of the output to binary16 differs
        y32 = m.predict_host(x16.float().pin_memory(), chunk=chunk).numpy().reshape(B, T)
        h32 = m.hidden.cpu().numpy().copy()
        y16 = m.predict_host(x16.pin_memory(), chunk=chunk)
        h16 = m.hidden.cpu().numpy()
    assert y16.dtype == torch.float16 and tuple(y16.shape) == (B, 1, T)
    assert np.array_equal(y16.numpy().reshape(B, T), y32.astype(np.float16)), "only the output rounding may differ"
    assert np.array_equal(h16, h32)
    # against the oracle on the original float32 input: input + output quantisation to 11 significant bits
    yo, _ = c_oracle.rnn_predict(c_oracle.GruWeights.from_state_dict(sd), x)
    assert esr(y16.numpy().reshape(B, T).astype(np.float32), yo) <= (1e-6 if mode == "fp32" else 1e-4)


def test_f16_transport_rejects_wrong_out_buffer():
    m = RNN(1, 64, 1, False).to(DEV)
    m.load_state_dict(load_ckpt("cfg2"))
    x16 = torch.zeros((2, 1, 256), dtype=torch.float16)
    with pytest.raises(RuntimeError):
        m.predict_host(x16, out=torch.empty((2, 1, 256), dtype=torch.float32))
