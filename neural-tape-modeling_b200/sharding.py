"""Stream sharding across GPUs: independent streams, contiguous block split, NO collective on
the data path.  torch.distributed is used only for the timing barrier and the max-over-ranks reduction of a duration.

The reference has nothing to mirror here (single process, single device, code/model.py:61); this is the host logic of
`bench.py --gpus N` and of any caller that drives several devices."""
import torch


def shard_range(total_streams, rank, world):
    """[begin, end) of the streams rank `rank` owns when `total_streams` are split into `world` contiguous blocks whose
    sizes differ by at most one (the first `total_streams % world` ranks get the extra stream)."""
    if world <= 0 or not 0 <= rank < world or total_streams < 0:
        raise ValueError(f"bad shard request: total={total_streams} rank={rank} world={world}")
    base, extra = divmod(total_streams, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def weak_scaling_range(streams_per_gpu, rank):
    """bench.py's weak-scaling layout: every rank owns `streams_per_gpu` streams, numbered globally."""
    return rank * streams_per_gpu, (rank + 1) * streams_per_gpu


def max_over_ranks(value, device="cpu", dist=None):
    """MAX of a python float over all ranks (identity without an initialised process group)."""
    if dist is None or not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_rate(units_per_rank, world, seconds_max):
    """Whole-job throughput: all units processed by all ranks over the slowest rank's time."""
    return units_per_rank * world / seconds_max
