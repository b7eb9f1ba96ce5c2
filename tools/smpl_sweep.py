"""BASELINE config 5: standalone SMPL forward / LBS throughput sweep (frames per call 1k .. 64k): frames/s and output
GB/s against the measured HBM copy bandwidth.  Launch under torchrun for the multi-GPU (replicated, no collective) case.

    python tools/smpl_sweep.py [--frames 1024 2048 ...] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seeme_b200 import ops, synthetic as S  # noqa: E402

BYTES_PER_FRAME = 82_680 + 340      # 6890 x 3 fp32 written + pose/shape/translation read


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, nargs="*", default=[1024, 2048, 4096, 8192, 16384, 32768, 65536])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm = peaks.get("hbm_gbs", 6650.0)
    buf = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in S.smpl_buffers().items()}
    op = ops.SmplOp(buf, max_frames=max(args.frames))
    rows = []
    for F in args.frames:
        g = torch.Generator().manual_seed(F + rank)
        betas = (0.5 * torch.randn(F, 10, generator=g)).to(dev)
        pose = (0.3 * torch.randn(F, 69, generator=g)).to(dev)
        go = (0.3 * torch.randn(F, 3, generator=g)).to(dev)
        tr = torch.randn(F, 3, generator=g).to(dev)
        for _ in range(3):
            op.forward(betas, pose, go, tr)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            op.forward(betas, pose, go, tr)      # outputs (F x 82 680 B) exceed L2 from 2k frames up
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        ms = float(ms)
        gbs = F * BYTES_PER_FRAME / (ms / 1e3) / 1e9
        rows.append({"frames_per_gpu": F, "n_gpus": world, "ms": ms, "frames_per_s": world * F / (ms / 1e3), "gb_per_s_per_gpu": gbs,
                     "hbm_frac_of_measured": gbs / hbm})
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)
    if rank == 0 and args.json:
        json.dump({"hbm_peak_gbs": hbm, "rows": rows}, open(args.json, "w"), indent=1)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
