"""Multi-GPU host logic on CPU: world size 2, gloo backend (the N > 1 path of bench.py minus the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ntm_b200 import sharding, signals


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, T = 6, 400
        lo, hi = sharding.weak_scaling_range(B, rank)
        x = signals.stream_batch(B, T, first_stream=lo, dur=1.0)          # this rank's own streams, nothing exchanged
        dist.barrier()
        slowest = sharding.max_over_ranks(0.25 * (rank + 1), dist=dist)    # rank r "took" 0.25 (r + 1) s
        rate = sharding.aggregate_rate(B * T, world, slowest)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=x, lo=lo, hi=hi, slowest=slowest, rate=rate)
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_streams_without_exchange(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    B, T = 6, 400
    whole = signals.stream_batch(world * B, T, dur=1.0)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert (int(z["lo"]), int(z["hi"])) == (r * B, (r + 1) * B)
        assert np.array_equal(z["x"], whole[r * B:(r + 1) * B])           # global stream numbering is rank-independent
        assert float(z["slowest"]) == 0.5                                  # max over ranks, same on every rank
        assert float(z["rate"]) == world * B * T / 0.5


@pytest.mark.parametrize("total,world", [(1024, 1), (1024, 8), (65536, 8), (10, 4), (3, 8), (0, 2)])
def test_shard_range_partitions(total, world):
    ranges = [sharding.shard_range(total, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == total
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a1 >= a0
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(total, world, world)


def test_max_over_ranks_without_group():
    assert sharding.max_over_ranks(1.5) == 1.5
