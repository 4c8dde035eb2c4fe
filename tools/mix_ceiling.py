#!/usr/bin/env python3
"""Dev tool (GPU box): what a plain elementwise pass achieves on this GPU at the delay line's traffic mix (two floats read, one
written per sample: torch.add(a, b, out=c)) and at a copy's (one read, one written), same size as the bench's delay pass
(1024 x 1 440 000 floats).  The delay line's 4.7 TB/s is to be read against the first number, not against the copy peak."""
import torch

dev = "cuda:0"
B, T = 1024, 1440000
a = torch.randn(B, T, device=dev)
b = torch.randn(B, T, device=dev)
c = torch.empty_like(a)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def best(fn, n=10):
    fn(); torch.cuda.synchronize()
    t = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t = min(t, e0.elapsed_time(e1))
    return t


t_add = best(lambda: torch.add(a, b, out=c))
t_cpy = best(lambda: c.copy_(a))
t_sum = best(lambda: a.sum())
print(f"torch.add (2 reads + 1 write): {12 * B * T / t_add / 1e9:.2f} TB/s   copy (1 + 1): {8 * B * T / t_cpy / 1e9:.2f} TB/s   "
      f"sum (read only): {4 * B * T / t_sum / 1e9:.2f} TB/s")
