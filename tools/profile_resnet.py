"""Image-backbone timing on one GPU: ms per batch of B crops through ``seeme_resnet50_forward`` (CUDA events, after
warm-up) and the achieved FLOP rate (8.18 GFLOP per 224x224 crop).  ``python tools/profile_resnet.py [B]``"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seeme_b200 import _lib, ops, synthetic as S  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda:0"
sd = {k: v.to(dev) for k, v in S.resnet50_state(0).items()}
out = {k: v.to(dev) for k, v in S.output_images_state(0).items()}
op = ops.ResNet50Op(sd, out, max_batch=B)
x = torch.randn(B, 3, 224, 224, device=dev)
for _ in range(3):
    op(x)
torch.cuda.synchronize()
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = int(os.environ.get("RN_ITERS", 5))
e0.record()
for _ in range(iters):
    op(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flop = 2 * 4.089e9 * B
print(f"resnet50 B={B}: {ms:.2f} ms per batch, {B / ms * 1e3:.0f} crops/s, {flop / ms / 1e9:.0f} TFLOP/s algorithmic "
      f"(x3 executed: split-bf16), {(_lib.launch_count() - n0) // iters} launches per batch")
if os.environ.get("RN_CHECK", "1") == "1":     # feature error against the reference golden (tests/golden/resnet50_image.npz)
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "resnet50_image.npz"))
    _, feat = op(S.images(int(g["batch"]), int(g["seed"])).to(dev), want_feat=True)
    ref = torch.from_numpy(g["feat"])
    d = (feat.cpu() - ref)
    print(f"precision {os.environ.get('SEEME_RESNET_PRECISION', '3')}: max|feat - reference| = {d.abs().max():.3e}, "
          f"rel l2 = {d.norm() / ref.norm():.3e} (|feat| max {ref.abs().max():.2f})")
