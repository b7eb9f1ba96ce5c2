"""One hot-path step at the bench configuration, bracketed by cudaProfilerStart/Stop so that
`ncu --profile-from-start off` captures exactly one warm step (see profiles/README.md)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import seeme_b200  # noqa: E402
from seeme_b200 import synthetic as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--warm", type=int, default=2)
args = ap.parse_args()
dev = torch.device("cuda", 0)
B = args.batch
model = seeme_b200.build_model("config_mld_egobody.yaml", device=dev, guidance_scale=bench.GUIDANCE, max_batch=B, n_points=bench.N_POINTS)
batch = tuple(x.to(dev) if torch.is_tensor(x) else x for x in S.make_batch(B, n_points=bench.N_POINTS))
noise = {k: v.to(dev) for k, v in bench.make_noise(B).items()}
for _ in range(args.warm):
    model.ego_eval(batch, noise)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model.ego_eval(batch, noise)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step, batch", B)
