"""One synchronous MLD.ego_eval of the benchmarked batch (configs[1]: 256 sequences, CFG 7.5, 20 000 points) between
cudaProfilerStart/Stop -- for `ncu --profile-from-start off` launch lists (tools/ncu_step.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seeme_b200  # noqa: E402
from seeme_b200 import synthetic as S  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda:0"
model = seeme_b200.build_model("config_mld_egobody.yaml", device=dev, guidance_scale=7.5, max_batch=B, n_points=20000)
batch = tuple(x.to(dev) if torch.is_tensor(x) else x for x in S.make_batch(B, n_points=20000))
g = torch.Generator().manual_seed(7)
noise = {k: torch.randn(*s, generator=g).to(dev) for k, s in (("eps_int", (1, B, 256)), ("eps_unc", (1, B, 256)), ("x_T", (B, 1, 256)))}
for _ in range(2):
    model.ego_eval(batch, noise)
torch.cuda.synchronize()
torch.cuda.profiler.start()
rs = model.ego_eval(batch, noise)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("joints", tuple(rs["joints_rst"].shape), float(rs["joints_rst"].abs().max()))
