import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, seeme_b200
from seeme_b200 import synthetic as S
dev = torch.device("cuda", 0); B = 256
model = seeme_b200.build_model("config_mld_egobody.yaml", device=dev, guidance_scale=7.5, max_batch=B, n_points=20000)
batch = tuple(x.to(dev) if torch.is_tensor(x) else x for x in S.make_batch(B, n_points=20000))
noise = {k: v.to(dev) for k, v in bench.make_noise(B).items()}
def ms(fn, n=5):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
step = lambda: model.ego_eval(batch, noise)
for _ in range(3): step()
print("full", [round(ms(step), 2) for _ in range(3)])
orig = model._encode_scene; cache = {}
def cached(scene):
    if 0 not in cache: cache[0] = orig(scene)
    return cache[0]
model._encode_scene = cached
step()
print("cached", [round(ms(step), 2) for _ in range(3)])
model._encode_scene = orig
print("full again", [round(ms(step), 2) for _ in range(3)])
