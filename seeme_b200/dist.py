"""Data-parallel sharding of the test protocol (one process per GPU) and the per-epoch metric
reduction -- the path's only collective.

The reference runs under Lightning DDP (``test.py:93-106``): a ``DistributedSampler`` shards the
sequences, every ``ComputeMetrics`` state is declared ``dist_reduce_fx="sum"``
(``mld/models/metrics/compute.py:106-178``) and torchmetrics all-gathers + sums them at
``compute()``.  Here: the (sequence, repetition) work list is split contiguously by sequence so all
repetitions of a sequence stay on the rank that holds its scene embedding, and
the state vector of ``seeme_b200.metrics.EgoMetric`` is summed with ONE ``all_reduce`` (NCCL over
NVLink on GPUs, gloo in the CPU tests).  Nothing in the sampling path itself communicates.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) block of rank `rank`; sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_work(n_sequences: int, replication_times: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """(sequence, repetition) pairs of this rank: every repetition of an owned sequence (``TEST.REPLICATION_TIMES``,
    test.py:116-136) runs on the owning rank, so its scene embedding is computed once."""
    lo, hi = shard_range(n_sequences, rank, world)
    return [(s, r) for s in range(lo, hi) for r in range(replication_times)]


def reduce_metric_state(metric, device=None) -> None:
    """Sum the metric's state vector over all ranks in place (no-op without an initialised process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    v = metric.state_vector()
    if device is not None:
        v = v.to(device)
    dist.all_reduce(v, op=dist.ReduceOp.SUM)
    metric.load_state_vector(v.cpu())


def gather_per_sequence(values: torch.Tensor) -> torch.Tensor:
    """all_gather of a ``[n_local, K]`` per-sequence record (ragged n_local allowed) -> ``[n_total, K]`` in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return values
    world = dist.get_world_size()
    n = torch.tensor([values.shape[0]], device=values.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(int(s) for s in sizes))
    pad = torch.zeros(m, values.shape[1], dtype=values.dtype, device=values.device)
    pad[: values.shape[0]] = values
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: int(s)] for o, s in zip(out, sizes)], dim=0)
