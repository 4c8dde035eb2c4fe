"""torch.ops.ntm.* on the GPU: same results as the raw C ABI (ctypes), PyTorch's current stream is honoured, forward() is
capturable in a CUDA Graph and reusable output buffers work; plus the lifetime / validation fixes of round 2 (handles under
live streams, NaN delays, per-example losses in one launch, per-handle kernel selection, two host threads)."""
import ctypes
import threading

import numpy as np
import pytest
import torch

from conftest import load_ckpt
import ntm_b200
from ntm_b200 import DCPreESR, DiffDelRNN, ESRLoss, RNN, TimeVaryingDelayLine, lib, signals

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_rnn(tag="cfg2", mode="f16"):
    m = RNN(1, 64, 1, False).to(DEV)
    m.load_state_dict(load_ckpt(tag))
    m.mode = mode
    return m


@pytest.mark.parametrize("mode", ["fp32", "f16", "f16x3"])
def test_op_equals_c_abi_and_honours_current_stream(mode):
    m = make_rnn(mode=mode)
    B, T = 5, 700
    x = torch.from_numpy(signals.stream_batch(B, T)).to(DEV).reshape(B, 1, T)
    h0 = 0.1 * torch.randn(1, B, 64, device=DEV)
    with torch.inference_mode():
        handle = m._handle(torch.device(DEV))
        y, h = lib.ops().gru_forward(handle, lib.MODES[mode], x, h0, False)
        y2, h2 = torch.empty_like(y), torch.empty_like(h)
        rc = lib.load().ntm_gru_forward(handle, lib.MODES[mode], x.data_ptr(), T, y2.data_ptr(), T, h0.data_ptr(), h2.data_ptr(),
                                        B, T, 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        assert torch.equal(y, y2) and torch.equal(h, h2)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                     # the op launches on the stream that is current HERE
            big = torch.randn(64, 1 << 20, device=DEV).sin_().sum()      # keeps `side` busy in front of the launch
            xs = x * 1.0
            y3, h3 = lib.ops().gru_forward(handle, lib.MODES[mode], xs, h0, False)
        side.synchronize()
        assert torch.equal(y3, y) and torch.equal(h3, h) and bool(torch.isfinite(big))
        with pytest.raises(RuntimeError, match="unit stride"):
            lib.ops().gru_forward(handle, lib.MODES[mode], x.expand(B, 2, T)[:, :1, ::2], None, False)
        with pytest.raises(RuntimeError, match="h_in"):
            lib.ops().gru_forward(handle, lib.MODES[mode], x, h0[:, :2], False)


@pytest.mark.parametrize("mode", ["fp32", "f16"])
def test_forward_in_cuda_graph_equals_one_long_call(mode):
    """BASELINE cfg 5 / SURVEY 8d: consecutive 64-sample forward() calls with carried hidden state, captured ONCE in a CUDA
    Graph and replayed -- bit-identical to a single long call (no allocation, no synchronisation inside the op)."""
    m = make_rnn("cfg1", mode)
    nblk, L = 40, 64
    x = torch.from_numpy(signals.signal("sweepnoise", nblk * L, seed=2)).to(DEV).reshape(1, 1, -1)
    with torch.inference_mode():
        m.initialize_hidden(); m.warm_start()
        h0 = m.hidden.clone()
        y_all = m(x)
        h_all = m.hidden.clone()
        m.static_io = True
        xs = torch.zeros(1, 1, L, device=DEV)
        m.hidden = None
        m(xs)                                       # allocates the module's static buffers; hidden now IS the static state buffer
        m.hidden.copy_(h0)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ys = m(xs)
        m.hidden.copy_(h0)                          # (the capture did not run the kernel; reset anyway)
        outs = []
        for k in range(nblk):
            xs.copy_(x[:, :, k * L:(k + 1) * L])
            g.replay()
            outs.append(ys.clone())
        assert torch.equal(torch.cat(outs, 2), y_all) and torch.equal(m.hidden, h_all)
        # static_io without a graph: same buffers handed back every call
        m.hidden = h0.clone()
        a = m(x[:, :, :L])
        b = m(x[:, :, L:2 * L])
        assert a.data_ptr() == b.data_ptr() and torch.equal(b, y_all[:, :, L:2 * L])


def test_block_and_realtime_streams_survive_a_parameter_reload():
    """Round-1 advisor finding: streams cached the raw handle and dereferenced it after load_state_dict() freed it."""
    m = make_rnn("cfg1", "f16")
    x = torch.from_numpy(signals.signal("noise", 640, seed=1)).to(DEV).reshape(1, 1, -1)
    with torch.inference_mode():
        m.initialize_hidden(); m.warm_start()
        y_all = m(x)
        m.initialize_hidden(); m.warm_start()
        bs = m.block_stream(1, 64)
        out = [bs.process(x[:, :, :64]).clone()]
        m.load_state_dict(load_ckpt("cfg1"))                     # releases the packed handle
        out += [bs.process(x[:, :, 64 * k:64 * k + 64]).clone() for k in range(1, 10)]
        assert torch.equal(torch.cat(out, 2), y_all)
        m.initialize_hidden(); m.warm_start()
        torch.cuda.synchronize()
        rt = m.realtime_stream(1, 64, idle_timeout_ms=3000)
        xh = x.cpu().reshape(1, -1)
        first = rt.process(xh[:, :64].contiguous()).clone()
        m.load_state_dict(load_ckpt("cfg1"))                     # the stream keeps its own reference to the blob
        second = rt.process(xh[:, 64:128].contiguous()).clone()
        rt.close()
        assert torch.equal(torch.cat([first, second], 1), y_all.cpu().reshape(1, -1)[:, :128])


def test_delay_check_rejects_nan_and_host_pipeline_checks_too():
    dl = TimeVaryingDelayLine(max_delay=16)
    dl.init_buffer(2)
    x = torch.randn(2, 1, 64, device=DEV)
    d = torch.rand(2, 1, 64, device=DEV) * 16
    d[1, 0, 7] = float("nan")
    with torch.inference_mode():
        with pytest.raises(AssertionError):
            dl(x, d)
        assert lib.ops().delay_check(torch.zeros(2, 1, 64, device=DEV), 16)
        md = DiffDelRNN(1, 64, 1, False, max_delay=32).to(DEV)
        md.load_state_dict(load_ckpt("cfg3"))
        xh = torch.zeros(2, 1, 256).pin_memory()
        with pytest.raises(AssertionError):                      # code/model.py:283 on the host pipeline as well
            md.predict_host(xh, torch.full((2, 1, 256), 40.0).pin_memory())
        y, pre = md.predict_host(xh, torch.full((2, 1, 256), 3.5).pin_memory())
        assert bool(torch.isfinite(y).all())


@pytest.mark.parametrize("loss", [ESRLoss(), DCPreESR()])
def test_per_example_losses_equal_the_one_by_one_loop(loss):
    B, T, cut = 7, 9000, 300
    g = torch.Generator().manual_seed(3)
    t = (0.2 * torch.randn(B, 1, T, generator=g)).to(DEV)
    o = t + (0.02 * torch.randn(B, 1, T, generator=g)).to(DEV)
    lens = torch.tensor([9000, 8999, 5000, 2049, 301, 300, 7], dtype=torch.int64)
    with torch.inference_mode():
        got = loss.per_example(o, t, torch.full((B,), cut, dtype=torch.int64), (lens - cut).clamp(min=0)).cpu()
        for b in range(B):
            n = int(lens[b])
            if n > cut:
                want = float(loss(o[b:b + 1, :, cut:n], t[b:b + 1, :, cut:n]))
                assert abs(float(got[b]) - want) <= 2e-6 * max(1.0, abs(want)), (b, float(got[b]), want)
            else:
                assert float(got[b]) == 0.0
        whole = loss.per_example(o, t).cpu()
        assert abs(float(whole[0]) - float(loss(o[:1], t[:1]))) <= 2e-6


def test_per_handle_kernel_selection_and_threads():
    """Kernel selection is per handle (no process-global mutable state on the launch path): two host threads drive two
    models with different forced kernels concurrently; each gets its own kernel and the single-threaded result."""
    L = lib.load()
    ms = [make_rnn("cfg2", "f16"), make_rnn("cfg1", "f16")]
    dev = torch.device(DEV)
    B, T = 24, 4000
    x = torch.from_numpy(signals.stream_batch(B, T)).to(DEV).reshape(B, 1, T)
    want, kernels = [], [(8, 3), (1, 4)]
    with torch.inference_mode():
        for m, k in zip(ms, kernels):
            assert L.ntm_handle_set_tuning(m._handle(dev), *k) == 0
            want.append(m.predict(x).clone())
            assert L.ntm_handle_last_kernel(m._handle(dev)) == (1 if k[1] == 3 else 3)
        got = [None, None]

        def work(i):
            with torch.inference_mode():
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    for _ in range(5):
                        got[i] = ms[i].predict(x)
                s.synchronize()

        threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
        [t.start() for t in threads]
        [t.join() for t in threads]
        for i in range(2):
            assert torch.equal(got[i], want[i])
            assert L.ntm_handle_last_kernel(ms[i]._handle(dev)) == (1 if kernels[i][1] == 3 else 3)
            assert L.ntm_handle_set_tuning(ms[i]._handle(dev), -1, 0) == 0
