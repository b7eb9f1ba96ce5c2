"""Sampler alone (config-2 shape: 256 sequences, CFG -> 512 rows, 50 steps): device time per run, per back-end.
    python tools/profile_sampler.py [B] [guidance] [backend ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seeme_b200 import ops, synthetic as S  # noqa: E402
from seeme_b200.modules import time_sinusoid  # noqa: E402
from seeme_b200.scheduler import DDIMScheduler  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
gs = float(sys.argv[2]) if len(sys.argv) > 2 else 7.5
backends = sys.argv[3:] or ["persistent", "tile", "graph"]
dev = "cuda:0"
op = ops.DenoiserOp({k: v.to(dev) for k, v in S.denoiser_state(0).items()}, max_rows=2 * B)
s = DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                  clip_sample=False, set_alpha_to_one=False, steps_offset=1)
s.set_timesteps(50)
ts = s.timesteps.tolist()
op.set_time_table(ts, time_sinusoid(s.timesteps))
g = torch.Generator().manual_seed(1)
R = 2 * B if gs > 1 else B
cond = torch.randn(2, R, 256, generator=g).to(dev)
xT = torch.randn(B, 256, generator=g).to(dev)
coef = s.step_coefficients()
zs = {}
for be in backends:
    op.set_backend(be)
    for _ in range(3):
        z = op.sample(xT, cond, gs, ts, coef)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        z = op.sample(xT, cond, gs, ts, coef)
    e1.record()
    torch.cuda.synchronize()
    zs[be] = z
    print(f"sampler[{be}] B={B} rows={R}: {e0.elapsed_time(e1) / n:.3f} ms per 50-step run; |z|max {float(z.abs().max()):.2f}")
ref = zs[backends[0]]
for be in backends[1:]:
    print(f"max|z[{be}] - z[{backends[0]}]| = {float((zs[be] - ref).abs().max()):.3e}")
