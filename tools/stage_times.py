"""Per-stage device time of MLD.ego_eval at the bench configuration (each stage bracketed by a device synchronise)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import seeme_b200  # noqa: E402
from seeme_b200 import synthetic as S  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
LANES = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
model = seeme_b200.build_model("config_mld_egobody.yaml", device=dev, guidance_scale=bench.GUIDANCE, max_batch=B, n_points=bench.N_POINTS, lanes=LANES)
batch = tuple(x.to(dev) if torch.is_tensor(x) else x for x in S.make_batch(B, n_points=bench.N_POINTS))
noise = {k: v.to(dev) for k, v in bench.make_noise(B).items()}
acc = {}


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def w(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
        return r
    setattr(obj, name, w)


for _ in range(5):
    model.ego_eval(batch, noise)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 20
for _ in range(n):
    model.ego_eval(batch, noise)
torch.cuda.synchronize()
print(f"unwrapped: {(time.perf_counter() - t0) / n * 1e3:.2f} ms per step (host wall clock, lanes={LANES})")
st = torch.cuda.memory_stats()
print("cudaMalloc calls", st.get("num_device_alloc"), "frees", st.get("num_device_free"), "retries", st.get("num_alloc_retries"),
      "reserved GB", st.get("reserved_bytes.all.current", 0) / 1e9)
if LANES > 1:
    sys.exit(0)
n = 5
wrap(model, "_encode_scene", "scene encoder")
wrap(model.vae, "encode", "vae.encode (x2)")
wrap(model, "_diffusion_reverse", "sampler")
wrap(model.vae, "decode", "vae.decode")
wrap(model, "_body", "renorm+SMPL (x3)")
t0 = time.perf_counter()
for _ in range(n):
    model.ego_eval(batch, noise)
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / n * 1e3
for k, v in acc.items():
    print(f"  {k:22s} {v / n:8.3f} ms")
print(f"  {'sum of stages':22s} {sum(acc.values()) / n:8.3f} ms;  wrapped step {tot:.2f} ms")
