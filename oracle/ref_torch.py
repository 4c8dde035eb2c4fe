"""torch restatement of the reference path: the SAME third-party arithmetic the reference calls
(torch.nn.GRU + torch.nn.Linear on CPU, torch 2.11), driven the way code/model.py drives it.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  This is the "port" timed as cpu_baseline and
by `bench.py --impl reference` (the reference's own classes cannot travel to the GPU box; their
arithmetic is torch's and torch is on the box).  oracle/make_golden.py proves this file bit-identical
to the imported reference classes in the build container.
"""
import math

import torch

WARM_LEN = 2 ** 10   # code/model.py:60, :386
SEGMENT = 2 ** 11    # code/model.py:222, :622


class RefNet:
    """GRU(1->H, batch_first) + Linear(H->1) pair built from a best.pth state_dict
    (code/model.py:44-45 / :364-365; keys listed in SURVEY.md section 2.1 #16)."""

    def __init__(self, state_dict, dtype=torch.float32):
        H = state_dict["GRU.weight_hh_l0"].shape[1]
        self.H = H
        self.has_bias = "output.bias" in state_dict
        self.gru = torch.nn.GRU(1, H, batch_first=True)
        self.out = torch.nn.Linear(H, 1, bias=self.has_bias)
        self.gru.load_state_dict({k[4:]: v for k, v in state_dict.items() if k.startswith("GRU.")})
        self.out.load_state_dict({k[7:]: v for k, v in state_dict.items() if k.startswith("output.")})
        self.gru.to(dtype).eval()
        self.out.to(dtype).eval()
        self.dtype = dtype

    @torch.no_grad()
    def forward(self, x, hidden=None, skip=False):
        """code/model.py:75-88: (B,1,T) -> reshape (B,T,1) -> GRU -> Linear (+skip) -> reshape back."""
        x = x.to(self.dtype) if (x.dtype == torch.float64 or self.dtype == torch.float64) else x
        xs = x.reshape(x.shape[0], x.shape[2], x.shape[1])
        hs, hidden = self.gru(xs, hidden)
        y = self.out(hs)
        if skip:
            y += xs
        return y.reshape(y.shape[0], y.shape[2], y.shape[1]), hidden

    @torch.no_grad()
    def warm_hidden(self, B=1):
        """code/model.py:58-65: batch-1 forward over 1024 zeros from zero state, broadcast to B streams."""
        _, h = self.forward(torch.zeros(1, 1, WARM_LEN, dtype=self.dtype), None)
        return h.expand(1, B, self.H).contiguous()

    @torch.no_grad()
    def predict(self, x, skip=False, segment=SEGMENT):
        """code/model.py:218-246 with the warm state broadcast over the batch."""
        h = self.warm_hidden(x.shape[0])
        out = torch.empty(x.shape, dtype=self.dtype)
        for s in range(int(math.ceil(x.shape[-1] / segment))):
            sl = slice(s * segment, (s + 1) * segment)
            out[:, :, sl], h = self.forward(x[:, :, sl], h, skip)
        return out, h


@torch.no_grad()
def delay_forward(x, dt, buffer, warmup=False):
    """code/model.py:283-318 -- unfold / triangular-weight form; returns (y, new_buffer)."""
    D = buffer.shape[2]
    assert D >= torch.max(dt)
    padded = torch.cat((buffer, x), dim=2)
    new_buffer = torch.cat((buffer[:, :, x.shape[2]:], x[:, :, -D:]), dim=2)
    if warmup:
        return x, new_buffer
    taps = torch.linspace(D, 0, D + 1)
    wts = torch.relu(1 - torch.abs(taps - dt.unsqueeze(3)))
    y = torch.sum(wts * padded.unfold(2, D + 1, 1), 3)
    return y, new_buffer


@torch.no_grad()
def diffdel_predict(net, x, d_traj, max_delay, skip=False, segment=SEGMENT):
    """code/model.py:618-653 (+ :372-391 warm start), batch generalised by broadcast.
    Returns (y, pre_d, hidden, buffer)."""
    B = x.shape[0]
    D = int(max_delay) + 1                       # code/model.py:375
    z = torch.zeros(1, 1, WARM_LEN)
    pre, h = net.forward(z, None, skip)
    _, buf = delay_forward(pre, z, torch.zeros(1, 1, D))
    h = h.expand(1, B, net.H).contiguous()
    buf = buf.expand(B, 1, D).contiguous()
    y = torch.empty(x.shape)
    pre_d = torch.empty(x.shape)
    for s in range(int(math.ceil(x.shape[-1] / segment))):
        sl = slice(s * segment, (s + 1) * segment)
        pre_d[:, :, sl], h = net.forward(x[:, :, sl], h, skip)
        y[:, :, sl], buf = delay_forward(pre_d[:, :, sl], d_traj[:, :, sl], buf)
    return y, pre_d, h, buf
