"""seeme_b200.metrics (batched, device-side) against numbers the UNMODIFIED reference metric classes produced on the same
seeded inputs (tests/golden/metrics.npz, written by oracle/make_golden_metrics.py).  Runs on the CPU here and on cuda:0 on
the GPU box -- the fixture travels, the reference tree does not."""
import math
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz")


def _check(dev):
    from seeme_b200.metrics import EgoMetric, MRMetric, vertice_pve
    g = dict(np.load(GOLDEN))
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    jp, jr, qp, qr, ji, qi = (t(k) for k in ("ego_jp", "ego_jr", "ego_qp", "ego_qr", "ego_ji", "ego_qi"))
    lengths = g["ego_lengths"].tolist()
    for split in ("test", "val"):
        m = EgoMetric(njoints=23)
        m.update(split, jp, jr, qp, qr, ji[:, :, [0]], ji, qi, None, lengths, {})
        m.update(split, jp.flip(0), jr.flip(0), qp.view(5, 60, 4).flip(0).reshape(-1, 4), qr.view(5, 60, 4).flip(0).reshape(-1, 4),
                 ji[:, :, [0]], ji, qi, None, lengths[::-1], {})
        s = m.state
        for k in ("MPJPE", "ROOT_ERROR", "ACCL", "HEAD_ORIENTATION_ERROR"):
            assert s[k] == pytest.approx(float(g[f"ego_{split}_{k}"]), rel=1e-5, abs=1e-6), (split, k)       # state sums
        for k in ("count", "count_seq", "count_seq_root", "count_seq_accl", "count_seq_head_orientation"):
            assert s[k] == float(g[f"ego_{split}_{k}"]), (split, k)
        res = m.compute()
        for k in ("MPJPE", "ROOT_ERROR", "ACCL", "HEAD_ORIENTATION_ERROR"):
            ref = float(g[f"ego_{split}_compute_{k}"])
            assert (math.isnan(ref) and math.isnan(res[k])) or res[k] == pytest.approx(ref, rel=1e-5), (split, k)
        if dev != "cpu":
            assert m._dev.is_cuda          # the sums never left the device before `state` / `compute`
    b = MRMetric(njoints=22, jointstype="humanml3d")
    for _ in range(2):
        b.update(t("mr_rst"), t("mr_ref"), g["mr_lengths"].tolist())
    rb = b.compute()
    for k in ("MPJPE", "PAMPJPE", "ACCEL"):
        assert rb[k] == pytest.approx(float(g[f"mr_{k}"]), rel=2e-5), k
    assert b.state["count"] == float(g["mr_count"]) and b.state["count_seq"] == float(g["mr_count_seq"])
    for al in ("none", "scale", "procrustes"):
        assert float(vertice_pve(t("pve_pred"), t("pve_target"), al)) == pytest.approx(float(g[f"pve_{al}"]), rel=1e-4), al


def test_metrics_vs_reference_fixture_cpu():
    _check("cpu")


@pytest.mark.gpu
def test_metrics_vs_reference_fixture_cuda():
    _check("cuda:0")


@pytest.mark.gpu
def test_metric_update_does_not_synchronise_the_host():
    """EgoMetric.update (the default METRIC.TYPE, called from the pipeline's retire loop) enqueues device work only: torch's
    sync-debug mode raises on any synchronising call.  (MRMetric keeps its sums on the device too, but its one batched
    torch.linalg.svd checks the cuSOLVER status on the host.)"""
    from seeme_b200.metrics import EgoMetric
    dev = "cuda:0"
    g = dict(np.load(GOLDEN))
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    jp, jr, qp, qr, ji, qi = (t(k) for k in ("ego_jp", "ego_jr", "ego_qp", "ego_qr", "ego_ji", "ego_qi"))
    m = EgoMetric(njoints=23)
    root_i, lens = ji[:, :, 0:1], g["ego_lengths"].tolist()
    m.update("test", jp, jr, qp, qr, root_i, ji, qi, None, lens, {})   # warm-up (allocations)
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")       # any synchronising call (.item(), .cpu(), a pageable copy, ...) raises
    try:
        m.update("test", jp, jr, qp, qr, root_i, ji, qi, None, lens, {})
    finally:
        torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()
    assert m.state["count"] == 2 * float(sum(g["ego_lengths"].tolist()))
