"""Seeded synthetic audio in the family of the reference's input generator
(code/generate-inputs.py:116-130: sine, unit pulse train, logarithmic chirp) plus Gaussian noise and a
wow/flutter delay trajectory.  numpy on the host (bit-reproducible from the seed),
torch on a device for the bench-sized workloads.
"""
import math

import numpy as np
import torch

FS = 48000


def logchirp(T, fs=FS, f0=20.0, f1=20000.0, dur=None):
    """sin-phase logarithmic sweep f0 -> f1 over `dur` seconds (scipy.signal.chirp(method='logarithmic',
    phi=-90), the call in code/generate-inputs.py:125-130)."""
    dur = T / fs if dur is None else dur
    t = np.arange(T, dtype=np.float64) / fs
    beta = dur / math.log(f1 / f0)
    return np.sin(2.0 * math.pi * beta * f0 * (np.power(f1 / f0, t / dur) - 1.0))


def pulse_train(T, fs=FS, rate=100.0):
    """Unit pulses every fs/rate samples (code/generate-inputs.py:118-123)."""
    x = np.zeros(T, dtype=np.float64)
    x[::int(round(fs / rate))] = 1.0
    return x


def sine(T, freq, fs=FS):
    return np.sin(2.0 * math.pi * freq * np.arange(T, dtype=np.float64) / fs)


def noise(T, seed, sigma=1.0):
    return np.random.default_rng(int(seed)).standard_normal(T) * sigma


def signal(kind, T, seed=0, fs=FS, dur=None):
    """Named test signals, all inside the checkpoints' numerically stable regime."""
    if kind == "sweepnoise":        # cfg 1: 0.25*sweep + N(0, 0.05^2)
        x = 0.25 * logchirp(T, fs, dur=dur) + noise(T, seed, 0.05)
    elif kind == "sweepnoise_lo":   # cfg 2 mix 0: 0.1*sweep + N(0, 0.05^2)
        x = 0.1 * logchirp(T, fs, dur=dur) + noise(T, seed, 0.05)
    elif kind == "noise":           # N(0, 0.1^2)
        x = noise(T, seed, 0.1)
    elif kind == "pulse":           # 0.5 * 100 Hz unit pulse train
        x = 0.5 * pulse_train(T, fs)
    elif kind == "sine":            # 0.5 * sine, log-uniform 50 Hz .. 2 kHz from the seed
        f = 50.0 * (2000.0 / 50.0) ** np.random.default_rng(int(seed) + 7919).random()
        x = 0.5 * sine(T, f, fs)
    elif kind == "sine1k":
        x = 0.25 * sine(T, 1000.0, fs)
    elif kind == "silence":
        x = np.zeros(T)
    else:
        raise ValueError(kind)
    return x.astype(np.float32)


MIX = ("sweepnoise_lo", "noise", "pulse", "sine")


def stream_batch(B, T, first_stream=0, fs=FS, dur=None):
    """cfg 2/3/4 mix: stream s uses seed s and signal kind MIX[s % 4].  Returns (B, T) float32."""
    out = np.empty((B, T), dtype=np.float32)
    for i in range(B):
        s = first_stream + i
        out[i] = signal(MIX[s % 4], T, seed=s, fs=fs, dur=dur)
    return out


def delay_trajectory(B, T, first_stream=0, fs=FS):
    """cfg 3 wow/flutter trajectory in SAMPLES: 240 + 48 sin(2 pi 0.5 t + phi_s) + 4 sin(2 pi 12 t + psi_s)."""
    t = np.arange(T, dtype=np.float64) / fs
    out = np.empty((B, T), dtype=np.float32)
    for i in range(B):
        rng = np.random.default_rng(100003 + first_stream + i)
        phi, psi = rng.random(2) * 2.0 * math.pi
        out[i] = 240.0 + 48.0 * np.sin(2 * math.pi * 0.5 * t + phi) + 4.0 * np.sin(2 * math.pi * 12.0 * t + psi)
    return out


DELAY_MAX = int(1.25 * 292.0)   # max_delay ctor argument for the trajectory above (code/test-model.py:223)


# ---------------------------------------------------------------------------------------------
# Device-side generation for bench-sized batches (same family; NOT bit-identical to the numpy path).

@torch.no_grad()
def stream_batch_device(B, T, device, first_stream=0, fs=FS, dur=None, out=None, chunk=1 << 22):
    """(B, T) float32 on `device`; stream s: kind MIX[s % 4], seeded by s.  Generated in time-chunks so the
    float64 temporaries stay small."""
    x = torch.empty(B, T, dtype=torch.float32, device=device) if out is None else out
    dur = T / fs if dur is None else dur
    beta = dur / math.log(20000.0 / 20.0)
    sid = torch.arange(first_stream, first_stream + B, device=device)
    kind = sid % 4
    g = torch.Generator(device=device)
    g.manual_seed(1234567 + first_stream)
    u = torch.rand(B, generator=g, device=device, dtype=torch.float64)
    freq = 50.0 * (2000.0 / 50.0) ** u
    period = int(round(fs / 100.0))
    for t0 in range(0, T, chunk):
        t1 = min(T, t0 + chunk)
        n = torch.arange(t0, t1, device=device, dtype=torch.float64)
        t = n / fs
        sweep = torch.sin(2.0 * math.pi * beta * 20.0 * (torch.pow(torch.tensor(1000.0, dtype=torch.float64,
                                                                                    device=device), t / dur) - 1.0))
        pulses = ((torch.arange(t0, t1, device=device) % period) == 0).to(torch.float32) * 0.5
        for k in range(4):
            rows = (kind == k).nonzero().flatten()
            if rows.numel() == 0:
                continue
            if k == 0:
                blk = torch.randn(rows.numel(), t1 - t0, generator=g, device=device) * 0.05
                blk += (0.1 * sweep).to(torch.float32)
            elif k == 1:
                blk = torch.randn(rows.numel(), t1 - t0, generator=g, device=device) * 0.1
            elif k == 2:
                blk = pulses.expand(rows.numel(), -1)
            else:
                blk = (0.5 * torch.sin(2.0 * math.pi * freq[rows, None] * t[None, :])).to(torch.float32)
            x[rows, t0:t1] = blk
    return x


@torch.no_grad()
def delay_trajectory_device(B, T, device, first_stream=0, fs=FS, chunk=1 << 22):
    d = torch.empty(B, T, dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(7654321 + first_stream)
    ph = torch.rand(B, 2, generator=g, device=device, dtype=torch.float64) * 2.0 * math.pi
    for t0 in range(0, T, chunk):
        t1 = min(T, t0 + chunk)
        t = torch.arange(t0, t1, device=device, dtype=torch.float64) / fs
        d[:, t0:t1] = (240.0 + 48.0 * torch.sin(2 * math.pi * 0.5 * t[None, :] + ph[:, 0:1])
                       + 4.0 * torch.sin(2 * math.pi * 12.0 * t[None, :] + ph[:, 1:2])).to(torch.float32)
    return d
