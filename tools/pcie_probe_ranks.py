#!/usr/bin/env python3
"""Dev tool (GPU box, torchrun): what ONE HOST gives k ranks that stream to / from their GPUs at the same time -- the ceiling
of the bench's end-to-end number at N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/pcie_probe_ranks.py [GB per direction and rank, default 2]

For k = 1, 2, 4, .. world: ranks 0 .. k-1 copy pinned host memory to the device, the device to pinned host memory, and both at
once (the shape of RNN.predict_host: a 2-D time-chunk copy each way on its own stream); the other ranks idle at the barrier.
Rank 0 prints one line per k: per-rank min / mean GB/s and the aggregate, per direction and for the concurrent case.
Timed on the host around a stream synchronize, max over the active ranks (gloo all_reduce)."""
import os
import sys
import time

import torch
import torch.distributed as dist

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("gloo")

B = 1024
T = int(gb * 1e9 / 4 / B) // 16384 * 16384
C = 16384                                   # predict_host's default chunk at 1024 streams (64 MiB per staged array)
xh = torch.empty((B, T), dtype=torch.float32, pin_memory=True).normal_(0, 0.1)
yh = torch.empty((B, T), dtype=torch.float32, pin_memory=True).zero_()
stage_in = [torch.empty((B, C), dtype=torch.float32, device=dev) for _ in range(2)]
stage_out = [torch.zeros((B, C), dtype=torch.float32, device=dev) for _ in range(2)]
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
nbytes = B * T * 4


def h2d():
    with torch.cuda.stream(s_in):
        for c in range(T // C):
            stage_in[c & 1].copy_(xh[:, c * C:(c + 1) * C], non_blocking=True)


def d2h():
    with torch.cuda.stream(s_out):
        for c in range(T // C):
            yh[:, c * C:(c + 1) * C].copy_(stage_out[c & 1], non_blocking=True)


def both():
    h2d()
    d2h()


def timed(fn, active):
    dist.barrier()
    if not active:
        dist.barrier()
        return 0.0
    fn()
    torch.cuda.synchronize(dev)
    dist.barrier()
    t = time.perf_counter()
    fn()
    torch.cuda.synchronize(dev)
    return time.perf_counter() - t


if rank == 0:
    print(f"# {world} ranks available, {nbytes/1e9:.2f} GB per direction and rank, 2-D chunks of {C} samples x {B} rows, host cpus {os.cpu_count()}")
k = 1
while k <= world:
    active = rank < k
    row = []
    for name, fn, mult in (("H2D", h2d, 1), ("D2H", d2h, 1), ("H2D+D2H", both, 2)):
        dist.barrier()
        # warm-up pass and the timed pass are both inside timed(); the two barriers keep idle ranks in step
        dt = timed(fn, active)
        rate = torch.tensor([mult * nbytes / dt / 1e9 if active else 0.0], dtype=torch.float64)
        rates = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(rates, rate)
        r = [float(v) for v in rates[:k]]
        row.append(f"{name}: per rank min {min(r):5.1f} mean {sum(r)/k:5.1f} GB/s, aggregate {sum(r):6.1f}")
    if rank == 0:
        print(f"{k} active rank(s) | " + " | ".join(row), flush=True)
    k *= 2
dist.destroy_process_group()
