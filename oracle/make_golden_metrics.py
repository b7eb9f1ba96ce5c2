"""TEST INFRASTRUCTURE ONLY -- writes ``tests/golden/metrics.npz``: seeded inputs and the numbers the UNMODIFIED reference
metric code computes on them (``ComputeMetrics`` mld/models/metrics/compute.py:349-580, ``MRMetrics`` mr.py:73-96,
``vertice_pve`` metrics_utils_egobody.py:144-171), so the GPU box (no reference tree) can compare the device-side
``seeme_b200.metrics`` with reference results.

    python -m oracle.make_golden_metrics          # build container only; /root/reference must exist
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_modules as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "metrics.npz")


def ego_inputs():
    g = torch.Generator().manual_seed(11)
    B, T_ = 5, 60
    jr = torch.randn(B, T_, 24, 3, generator=g) * 0.3
    jp = jr + 0.02 * torch.randn(B, T_, 24, 3, generator=g)
    jp[3] += 1.0 * torch.randn(T_, 1, 3, generator=g)               # a sequence the root-error gate rejects
    qr = torch.nn.functional.normalize(torch.randn(B * T_, 4, generator=g), dim=1)
    qp = torch.nn.functional.normalize(qr + 0.05 * torch.randn(B * T_, 4, generator=g), dim=1)
    qp[T_:2 * T_] = torch.nn.functional.normalize(torch.randn(T_, 4, generator=g), dim=1)   # head-orientation gate
    ji = torch.randn(B, T_, 24, 3, generator=g)
    qi = torch.nn.functional.normalize(torch.randn(B * T_, 4, generator=g), dim=1)
    return jp, jr, qp, qr, ji, qi, [60, 45, 60, 60, 20]


def mr_inputs():
    g = torch.Generator().manual_seed(5)
    B, T_, J = 4, 60, 22
    ref_j = torch.randn(B, T_, J, 3, generator=g) * 0.4
    rst_j = 1.1 * ref_j @ torch.linalg.qr(torch.randn(3, 3, generator=g))[0] + 0.03 * torch.randn(B, T_, J, 3, generator=g) + 0.2
    pv = torch.randn(6, 500, 3, generator=g)
    tv = 0.9 * pv @ torch.linalg.qr(torch.randn(3, 3, generator=g))[0] + 0.01 * torch.randn(6, 500, 3, generator=g)
    return rst_j, ref_j, [60, 33, 60, 48], pv, tv


def main():
    R.import_mld()
    from mld.models.metrics.compute import ComputeMetrics
    from mld.models.metrics.mr import MRMetrics
    from mld.models.metrics import metrics_utils_egobody as U
    out = {}
    jp, jr, qp, qr, ji, qi, lengths = ego_inputs()
    out.update({"ego_jp": jp.numpy(), "ego_jr": jr.numpy(), "ego_qp": qp.numpy(), "ego_qr": qr.numpy(), "ego_ji": ji.numpy(),
                "ego_qi": qi.numpy(), "ego_lengths": np.asarray(lengths)})
    keys = ("MPJPE", "ROOT_ERROR", "ACCL", "HEAD_ORIENTATION_ERROR", "count", "count_seq", "count_seq_root", "count_seq_accl",
            "count_seq_head_orientation")
    for split in ("test", "val"):
        ref = ComputeMetrics(njoints=23, jointstype="humanml3d", dist_sync_on_step=False)
        ref.update(split, jp, jr, qp, qr, ji[:, :, [0]], ji, qi, None, lengths, {})
        ref.update(split, jp.flip(0), jr.flip(0), qp.view(5, 60, 4).flip(0).reshape(-1, 4), qr.view(5, 60, 4).flip(0).reshape(-1, 4),
                   ji[:, :, [0]], ji, qi, None, lengths[::-1], {})       # second batch: accumulation across updates
        for k in keys:
            out[f"ego_{split}_{k}"] = np.float64(float(getattr(ref, k)))
        res = ref.compute(sanity_flag=False)
        for k, v in res.items():
            out[f"ego_{split}_compute_{k}"] = np.float64(float(v))
    rst_j, ref_j, mr_lengths, pv, tv = mr_inputs()
    out.update({"mr_rst": rst_j.numpy(), "mr_ref": ref_j.numpy(), "mr_lengths": np.asarray(mr_lengths), "pve_pred": pv.numpy(),
                "pve_target": tv.numpy()})
    a = MRMetrics(njoints=22, jointstype="humanml3d", dist_sync_on_step=False)
    for _ in range(2):
        a.update(rst_j, ref_j, mr_lengths)
    for k, v in a.compute(sanity_flag=False).items():
        out[f"mr_{k}"] = np.float64(float(v))
    out["mr_count"], out["mr_count_seq"] = np.float64(float(a.count)), np.float64(float(a.count_seq))
    for al in ("none", "scale", "procrustes"):
        out[f"pve_{al}"] = np.float64(float(U.vertice_pve(pv.numpy(), tv.numpy(), alignment=al)))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: float(v) for k, v in out.items() if np.ndim(v) == 0})


if __name__ == "__main__":
    main()
