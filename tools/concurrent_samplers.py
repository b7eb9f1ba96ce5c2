"""Do two independent 50-step sampler graphs on two CUDA streams overlap?  (latency-bound chains of small kernels)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seeme_b200 import ops, synthetic as S  # noqa: E402
from seeme_b200.modules import time_sinusoid  # noqa: E402
from seeme_b200.scheduler import DDIMScheduler  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = 64
dev = "cuda:0"
sd = {k: v.to(dev) for k, v in S.denoiser_state(0).items()}
s = DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                  clip_sample=False, set_alpha_to_one=False, steps_offset=1)
s.set_timesteps(50)
ts = s.timesteps.tolist()
coef = s.step_coefficients()
opsl, streams, ins = [], [], []
g = torch.Generator().manual_seed(1)
for k in range(K):
    op = ops.DenoiserOp(sd, max_rows=2 * B)
    op.set_time_table(ts, time_sinusoid(s.timesteps))
    opsl.append(op)
    streams.append(torch.cuda.Stream())
    ins.append((torch.randn(B, 256, generator=g).to(dev), torch.randn(2, 2 * B, 256, generator=g).to(dev)))
for k in range(K):
    with torch.cuda.stream(streams[k]):
        for _ in range(2):
            opsl[k].sample(ins[k][0], ins[k][1], 7.5, ts, coef)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 5
for _ in range(n):
    for k in range(K):
        with torch.cuda.stream(streams[k]):
            opsl[k].sample(ins[k][0], ins[k][1], 7.5, ts, coef)
torch.cuda.synchronize()
print(f"{K} concurrent samplers (B={B} each): {(time.perf_counter() - t0) / n * 1e3:.2f} ms per round")
