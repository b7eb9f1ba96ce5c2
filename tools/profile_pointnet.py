"""Scene encoder alone (one 32-sample chunk of 20 000-point clouds) for ncu captures of its kernels."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seeme_b200 import _lib, ops, synthetic as S  # noqa: E402

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = "cuda:0"
W = {k: v.to(dev) for k, v in S.pointnet_state(0).items()}
Wo = {k: v.to(dev) for k, v in S.output_scene_state(0).items()}
op = ops.PointNetOp(W, Wo, max_batch=B, max_points=20000, precision=prec)
p = S.egobody_scene(B, 20000, torch.Generator().manual_seed(3)).to(dev)
IT = int(os.environ.get("PN_ITERS", "3"))
for _ in range(2 if IT <= 3 else 10):
    op(p)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(IT):
    op(p)
e1.record()
torch.cuda.synchronize()
print(f"precision {prec} batch {B}: {e0.elapsed_time(e1) / IT:.3f} ms per pass")
if IT > 3:      # per-kernel device times (CUDA-event pairs inside the library): 6 = blocks 1..3, 7 = block 0
    _lib.prof_read(6), _lib.prof_read(7)
    _lib.prof_enable(True)
    for _ in range(IT):
        op(p)
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    for i, name in ((7, "block0"), (6, "blocks1-3")):
        ms, n = _lib.prof_read(i)
        print(f"  {name}: {ms / max(n, 1):.4f} ms per launch over {n} launches")
