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

import ctypes
import glob

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


# cudaMemcpy2DAsync straight from the CUDA runtime torch already loaded (torch's own copy_ of a strided host view goes through a
# CPU-side repack, which is not what the engine's pipeline does: csrc/capi.cu predict_host)
_cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")) + ["libcudart.so.12", "libcudart.so"]
for _c in _cands:
    try:
        cudart = ctypes.CDLL(_c)
        break
    except OSError:
        cudart = None
assert cudart is not None, "libcudart not found"
cudart.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                     ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
cudart.cudaMemcpy2DAsync.restype = ctypes.c_int


def copy2d(dst, dpitch, src, spitch, width, height, kind, stream):
    rc = cudart.cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, stream.cuda_stream)
    assert rc == 0, f"cudaMemcpy2DAsync -> {rc}"


def h2d():
    for c in range(T // C):
        copy2d(stage_in[c & 1].data_ptr(), C * 4, xh.data_ptr() + c * C * 4, T * 4, C * 4, B, 1, s_in)


def d2h():
    for c in range(T // C):
        copy2d(yh.data_ptr() + c * C * 4, T * 4, stage_out[c & 1].data_ptr(), C * 4, C * 4, B, 2, s_out)


def both():
    h2d()
    d2h()


big_in = torch.empty((B, T), dtype=torch.float32, device=dev)
big_out = torch.zeros((B, T), dtype=torch.float32, device=dev)


def h2d_1d():                               # one contiguous cudaMemcpyAsync of the whole buffer
    with torch.cuda.stream(s_in):
        big_in.copy_(xh, non_blocking=True)


def d2h_1d():
    with torch.cuda.stream(s_out):
        yh.copy_(big_out, non_blocking=True)


def both_1d():
    h2d_1d()
    d2h_1d()


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
    print(f"# {world} ranks available, {nbytes/1e9:.2f} GB per direction and rank, host cpus {os.cpu_count()}; GB/s per active rank (min, mean) and "
          f"summed over the active ranks; H2D / D2H: cudaMemcpy2DAsync in chunks of {C} samples x {B} rows (the engine's pipeline), 1-D: one "
          f"contiguous cudaMemcpyAsync; H2D+D2H: both directions at once, bytes of both counted")
k = 1
while k <= world:
    active = rank < k
    row = []
    for name, fn, mult in (("H2D", h2d, 1), ("D2H", d2h, 1), ("H2D+D2H", both, 2), ("1-D H2D", h2d_1d, 1), ("1-D D2H", d2h_1d, 1),
                           ("1-D H2D+D2H", both_1d, 2)):
        dist.barrier()
        # warm-up pass and the timed pass are both inside timed(); the two barriers keep idle ranks in step
        dt = timed(fn, active)
        rate = torch.tensor([mult * nbytes / dt / 1e9 if active else 0.0], dtype=torch.float64)
        rates = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(rates, rate)
        r = [float(v) for v in rates[:k]]
        row.append(f"{name}: min {min(r):5.1f} mean {sum(r)/k:5.1f} sum {sum(r):6.1f}")
    if rank == 0:
        print(f"{k} active rank(s) | " + " | ".join(row), flush=True)
    k *= 2
dist.destroy_process_group()
