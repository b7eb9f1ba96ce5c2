"""Throughput with D batches in flight (each on its own stream + kernel-side handles): does step k+1's scene encoder
overlap step k's latency-bound sampler chain?"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import seeme_b200  # noqa: E402
from seeme_b200 import modules as M, synthetic as S  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = 256
dev = torch.device("cuda", 0)
model = seeme_b200.build_model("config_mld_egobody.yaml", device=dev, guidance_scale=7.5, max_batch=B, n_points=20000, lanes=1)
batch = tuple(x.to(dev) if torch.is_tensor(x) else x for x in S.make_batch(B, n_points=20000))
noise = {k: v.to(dev) for k, v in bench.make_noise(B).items()}
streams = [torch.cuda.Stream() for _ in range(D)]
lengths = [60] * B


def submit(k):
    s = k % D
    M._LANE[0] = s
    try:
        with torch.cuda.stream(streams[s]):
            return model._ego_eval_one(batch, noise, defer_random=True, t_max=60, lengths=lengths)
    finally:
        M._LANE[0] = 0


for k in range(2 * D):
    submit(k)
torch.cuda.synchronize()
n = 24
t0 = time.perf_counter()
for k in range(n):
    submit(k)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n * 1e3
print(f"depth {D}: {dt:.2f} ms per step -> {B / dt * 1e3:.0f} sequences/s")
