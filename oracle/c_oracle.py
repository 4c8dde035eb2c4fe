"""ctypes front-end of oracle/ntm_oracle.c (numpy in, numpy out).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Every function names the reference lines
its C counterpart restates.
"""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libntm_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64


def build(force=False):
    """Compile libntm_oracle.so with gcc (seconds)."""
    src = os.path.join(_DIR, "ntm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _DIR, "libntm_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        gru_args32 = [_f32p] * 6 + [ctypes.c_int, _f32p, _f32p, _f32p, _i64, _i64, _i64, _i64, ctypes.c_int]
        gru_args64 = [_f32p] * 6 + [ctypes.c_int, _f32p, _f64p, _f64p, _i64, _i64, _i64, _i64, ctypes.c_int]
        _lib.ntm_oracle_gru_f32.argtypes = gru_args32
        _lib.ntm_oracle_gru_f64.argtypes = gru_args64
        dl = [_f32p] * 5 + [_i64] * 6 + [ctypes.c_int]
        _lib.ntm_oracle_delay_f32.argtypes = dl
        _lib.ntm_oracle_delay_window_f32.argtypes = dl
        _lib.ntm_oracle_esr.argtypes = [_f32p, _f32p, _i64]
        _lib.ntm_oracle_esr.restype = ctypes.c_double
        _lib.ntm_oracle_dcpre_esr.argtypes = [_f32p, _f32p, _i64, _i64, _i64, _i64, _f32p, _i64, _f64p]
        _lib.ntm_oracle_dcpre_esr.restype = ctypes.c_double
    return _lib


def _p32(a):
    return a.ctypes.data_as(_f32p) if a is not None else None


def _c32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class GruWeights:
    """The six tensors of a best.pth state_dict as float32 numpy arrays."""

    def __init__(self, w_ih, w_hh, b_ih, b_hh, w_out, b_out=None):
        self.w_ih = _c32(w_ih).reshape(-1)
        self.H = self.w_ih.shape[0] // 3
        self.w_hh = _c32(w_hh).reshape(3 * self.H, self.H)
        self.b_ih = _c32(b_ih).reshape(-1)
        self.b_hh = _c32(b_hh).reshape(-1)
        self.w_out = _c32(w_out).reshape(-1)
        self.b_out = None if b_out is None else _c32(b_out).reshape(-1)

    @classmethod
    def from_state_dict(cls, sd):
        g = lambda k: np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k])
        return cls(g("GRU.weight_ih_l0"), g("GRU.weight_hh_l0"), g("GRU.bias_ih_l0"), g("GRU.bias_hh_l0"),
                   g("output.weight"), g("output.bias") if "output.bias" in sd else None)

    def _ptrs(self):
        return [_p32(self.w_ih), _p32(self.w_hh), _p32(self.b_ih), _p32(self.b_hh), _p32(self.w_out),
                _p32(self.b_out)]


def gru_forward(w, x, h=None, skip=False, f64=False):
    """RNN.forward arithmetic (code/model.py:67-88; gates torch rnn.py:1221-1224).

    x: (B, T) float32; h: (B, H) or None (zeros, code/model.py:50-52).  Returns (y, h_out);
    float32 unless f64 (double state and output)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, T = x.shape
    if f64:
        hh = np.zeros((B, w.H), np.float64) if h is None else np.array(h, dtype=np.float64, order="C").reshape(B, w.H)
        y = np.empty((B, T), np.float64)
        rc = lib().ntm_oracle_gru_f64(*w._ptrs(), w.H, _p32(x), y.ctypes.data_as(_f64p),
                                      hh.ctypes.data_as(_f64p), B, T, T, T, int(skip))
    else:
        hh = np.zeros((B, w.H), np.float32) if h is None else np.array(h, dtype=np.float32, order="C").reshape(B, w.H)
        y = np.empty((B, T), np.float32)
        rc = lib().ntm_oracle_gru_f32(*w._ptrs(), w.H, _p32(x), _p32(y), _p32(hh), B, T, T, T, int(skip))
    if rc != 0:
        raise RuntimeError(f"ntm_oracle_gru failed: {rc}")
    return y, hh


def delay_forward(x, d, hist, warmup=False, window=False):
    """TimeVaryingDelayLine.forward (code/model.py:269-320).  x, d: (B, T); hist: (B, D).
    Returns (y, new_hist).  window=True runs the literal O(T*D) form."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    d = np.ascontiguousarray(d, dtype=np.float32)
    hist = np.ascontiguousarray(hist, dtype=np.float32)
    B, T = x.shape
    D = hist.shape[1]
    y = np.empty((B, T), np.float32)
    ho = np.empty((B, D), np.float32)
    fn = lib().ntm_oracle_delay_window_f32 if window else lib().ntm_oracle_delay_f32
    rc = fn(_p32(x), _p32(d), _p32(y), _p32(hist), _p32(ho), B, T, D, T, T, T, int(warmup))
    if rc == -3:
        raise AssertionError("max_delay >= max(dt) violated (code/model.py:283)")
    if rc != 0:
        raise RuntimeError(f"ntm_oracle_delay failed: {rc}")
    return y, ho


def esr(out, target):
    """ESRLoss, CoreAudioML/training.py:10-16 (epsilon 1e-5)."""
    o = np.ascontiguousarray(out, dtype=np.float32).reshape(-1)
    t = np.ascontiguousarray(target, dtype=np.float32).reshape(-1)
    return float(lib().ntm_oracle_esr(_p32(o), _p32(t), o.shape[0]))


WARM_LEN = 1024   # code/model.py:60, :386
SEGMENT = 2048    # code/model.py:222, :622


def rnn_predict(w, x, skip=False, f64=False):
    """RNN.predict (code/model.py:218-246) generalised to B streams by broadcasting the batch-1 warm state
    (SURVEY.md section 8b quirk 2): zero state -> 1024 zero samples -> the input."""
    B, T = x.shape
    _, h1 = gru_forward(w, np.zeros((1, WARM_LEN), np.float32), None, skip, f64)
    h = np.repeat(h1, B, axis=0)
    return gru_forward(w, x, h, skip, f64)


def diffdel_predict(w, x, d, max_delay, skip=False):
    """DiffDelRNN.predict (code/model.py:618-653): D = int(max_delay)+1 (code/model.py:375); warm start =
    GRU+delay on 1024 zeros with zero delay (code/model.py:382-391); then forward (code/model.py:393-424).
    Returns (y, pre_d, h, hist)."""
    B, T = x.shape
    D = int(max_delay) + 1
    z = np.zeros((1, WARM_LEN), np.float32)
    pre_w, h1 = gru_forward(w, z, None, skip)
    _, hist1 = delay_forward(pre_w, z, np.zeros((1, D), np.float32))
    h = np.repeat(h1, B, axis=0)
    hist = np.repeat(hist1, B, axis=0)
    pre_d, h = gru_forward(w, x, h, skip)
    y, hist = delay_forward(pre_d, d, hist)
    return y, pre_d, h, hist


def dcpre_taps(R=0.995, n=2000):
    """The reference's pre-emphasis impulse response (code/GreyBoxDRC/loss_funcs.py:10-11), float32:
    scipy.signal.dimpulse of (1 - z^-1)/(1 - R z^-1), n samples."""
    import scipy.signal as signal
    _, ir = signal.dimpulse(signal.dlti([1, -1], [1, -R]), n=n)
    return np.ascontiguousarray(ir[0][:, 0].astype(np.float32))


def dcpre_esr(out, target, dc_pre=True):
    """ESRLoss(dc_pre) of code/GreyBoxDRC/loss_funcs.py:32-52 (dc_pre=False: CoreAudioML/training.py:10-16) on
    (B, T) or (B, 1, T) arrays; returns (loss, sum_num, sum_den)."""
    o = _c32(out)
    t = _c32(target)
    o = o.reshape(-1, o.shape[-1])
    t = t.reshape(-1, t.shape[-1])
    taps = dcpre_taps() if dc_pre else None
    sums = np.zeros(2, dtype=np.float64)
    loss = lib().ntm_oracle_dcpre_esr(_p32(o), _p32(t), o.shape[0], o.shape[1], o.shape[1], t.shape[1],
                                      _p32(taps), 0 if taps is None else taps.shape[0], sums.ctypes.data_as(_f64p))
    return float(loss), float(sums[0]), float(sums[1])
