"""Host-side cost of one MLD.ego_eval_async submission (cProfile over N submissions with the pipeline full).
    python tools/host_profile.py [B] [config]"""
import cProfile
import os
import pstats
import sys
from collections import deque

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import seeme_b200  # noqa: E402
from seeme_b200 import synthetic as S  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = sys.argv[2] if len(sys.argv) > 2 else "config_mld_gimo.yaml"
dev = torch.device("cuda", 0)
model = seeme_b200.build_model(cfg, device=dev, guidance_scale=7.5, max_batch=B, n_points=20000)
if os.environ.get("HP_SKIP_SCENE"):
    _o = model._encode_scene
    _c = {}
    model._encode_scene = lambda sc: _c.setdefault(0, _o(sc))
model.prepare_pipeline()
batch = tuple(x.to(dev) if torch.is_tensor(x) else x for x in S.make_batch(B, n_points=20000))
if "interactee" not in model.condition:
    pass
noise = {k: v.to(dev) for k, v in bench.make_noise(B).items()}
D = int(model.pipeline_depth)


def run(n):
    pend = deque()
    for _ in range(n):
        pend.append(model.ego_eval_async(batch, noise))
        if len(pend) >= D:
            pend.popleft().synchronize()
    while pend:
        pend.popleft().synchronize()


run(6 * D)
torch.cuda.synchronize()
# where the allocations go: time of every torch.empty by size
import time  # noqa: E402
_empty = torch.empty
acc = {}


def timed_empty(*a, **k):
    t0 = time.perf_counter()
    r = _empty(*a, **k)
    dt = time.perf_counter() - t0
    key = r.numel() * r.element_size()
    e = acc.setdefault(key, [0, 0.0, 0.0])
    e[0] += 1; e[1] += dt; e[2] = max(e[2], dt)
    return r


torch.empty = timed_empty
t0 = time.perf_counter()
run(128)
print(f"unprofiled: {(time.perf_counter() - t0) / 128 * 1e3:.2f} ms per submission (incl. waiting when the pipeline is full)")
torch.empty = _empty
for kb, (n, tot, mx) in sorted(acc.items()):
    print(f"torch.empty {kb / 1e6:10.3f} MB: {n} calls, mean {tot / n * 1e6:7.1f} us, max {mx * 1e6:8.1f} us")
st0 = torch.cuda.memory_stats()
print("cudaMalloc calls so far", st0.get("num_device_alloc"), "frees", st0.get("num_device_free"))
pr = cProfile.Profile()
pr.enable()
run(128)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(40)
