"""Small sampler run (B sequences, CFG, n steps) on every back-end -- for compute-sanitizer and quick agreement checks."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seeme_b200 import ops, synthetic as S  # noqa: E402
from seeme_b200.modules import time_sinusoid  # noqa: E402
from seeme_b200.scheduler import DDIMScheduler  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 70
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
backends = sys.argv[3:] or ["graph", "persistent", "tile"]
dev = "cuda:0"
op = ops.DenoiserOp({k: v.to(dev) for k, v in S.denoiser_state(0).items()}, max_rows=2 * B)
s = DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                  clip_sample=False, set_alpha_to_one=False, steps_offset=1)
s.set_timesteps(n)
ts = s.timesteps.tolist()
op.set_time_table(ts, time_sinusoid(s.timesteps))
g = torch.Generator().manual_seed(1)
cond = torch.randn(2, 2 * B, 256, generator=g).to(dev)
xT = torch.randn(B, 256, generator=g).to(dev)
zs = []
for be in backends:
    op.set_backend(be)
    zs.append(op.sample(xT, cond, 7.5, ts, s.step_coefficients()))
    torch.cuda.synchronize()
for be, z in zip(backends[1:], zs[1:]):
    print(f"{be} vs {backends[0]}: max|dz| = {float((z - zs[0]).abs().max()):.3e} (|z|max {float(zs[0].abs().max()):.2f})")
print("done")
