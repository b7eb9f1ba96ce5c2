"""world_size-2 gloo tests (CPU) of the multi-rank host logic: work-list sharding by (sequence, repetition)
and the metric-state all-reduce that replaces torchmetrics' dist_reduce_fx="sum" gathers."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _metric_inputs(seed, B=3, T=60):
    g = torch.Generator().manual_seed(seed)
    jr = torch.randn(B, T, 24, 3, generator=g) * 0.3
    jp = jr + 0.02 * torch.randn(B, T, 24, 3, generator=g)
    qr = torch.nn.functional.normalize(torch.randn(B * T, 4, generator=g), dim=1)
    qp = torch.nn.functional.normalize(qr + 0.05 * torch.randn(B * T, 4, generator=g), dim=1)
    ji = torch.randn(B, T, 24, 3, generator=g)
    qi = torch.nn.functional.normalize(torch.randn(B * T, 4, generator=g), dim=1)
    return jp, jr, qp, qr, ji[:, :, [0]], ji, qi, None, [60, 41, 60][:B], {}


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from seeme_b200 import dist as D
    from seeme_b200.metrics import EgoMetric
    work = D.shard_work(5, 3, rank, world)
    m = EgoMetric()
    # each rank owns different sequences: seeds 100 + rank
    m.update("test", *_metric_inputs(100 + rank))
    D.reduce_metric_state(m)
    rec = torch.full((len(work), 2), float(rank))
    allrec = D.gather_per_sequence(rec)
    if rank == 0:
        torch.save({"state": m.state_vector(), "compute": m.compute(), "n_rec": allrec.shape[0], "work0": work,
                    "rec_ranks": allrec[:, 0].tolist()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_is_a_partition():
    from seeme_b200 import dist as D
    for n, w in ((256, 8), (5, 2), (7, 4), (3, 8)):
        blocks = [D.shard_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    work = [D.shard_work(5, 10, r, 2) for r in range(2)]
    assert sorted(work[0] + work[1]) == [(s, r) for s in range(5) for r in range(10)]
    # all repetitions of a sequence live on one rank
    assert {s for s, _ in work[0]}.isdisjoint({s for s, _ in work[1]})


def test_metric_allreduce_world2_gloo(tmp_path):
    from seeme_b200.metrics import EgoMetric
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    ref = EgoMetric()
    for r in range(2):
        ref.update("test", *_metric_inputs(100 + r))
    assert torch.allclose(got["state"], ref.state_vector(), rtol=1e-12, atol=0)
    assert got["compute"]["MPJPE"] == pytest.approx(ref.compute()["MPJPE"], rel=1e-12)
    assert got["n_rec"] == 15 and got["work0"] == [(s, r) for s in range(3) for r in range(3)]
    assert got["rec_ranks"] == [0.0] * 9 + [1.0] * 6


def test_replication_statistics_match_reference_formula():
    """seeme_b200.driver.summarize = test.py:32-38,138-148 (mean, 1.96 std / sqrt(n), min, max + raw lists)"""
    import numpy as np
    from seeme_b200.driver import summarize
    vals = [12.5, 11.0, 13.25, 12.0]
    out = summarize({"Metrics/MPJPE": vals}, 4)
    a = np.array(vals)
    assert out["Metrics/MPJPE/mean"] == float(a.mean())
    assert abs(out["Metrics/MPJPE/conf_interval"] - 1.96 * a.std() / 2.0) < 1e-12
    assert out["Metrics/MPJPE/min"] == 11.0 and out["Metrics/MPJPE/max"] == 13.25 and out["Metrics/MPJPE"] == vals
