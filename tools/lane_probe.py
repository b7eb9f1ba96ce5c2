"""Which stage refuses to overlap across lanes?  Runs one stage at a time for K sub-batches on K streams."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import seeme_b200  # noqa: E402
from seeme_b200 import modules as M, synthetic as S  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = 256
dev = torch.device("cuda", 0)
model = seeme_b200.build_model("config_mld_egobody.yaml", device=dev, guidance_scale=7.5, max_batch=B, n_points=20000, lanes=1)
batch = tuple(x.to(dev) if torch.is_tensor(x) else x for x in S.make_batch(B, n_points=20000))
feats_ref, transl, beta, utils_, scene, length, _ = batch
streams = [torch.cuda.Stream() for _ in range(K)]
per = B // K
f_int = torch.cat([feats_ref[:, :, 1, :], transl[:, 1, :, :]], dim=-1).contiguous()
len_dev = length.reshape(-1).to(torch.int32)
eps = torch.randn(1, B, 256, device=dev)
cond = torch.randn(2 * B, 2, 256, device=dev)
xT = torch.randn(B, 1, 256, device=dev)
z = torch.randn(1, B, 256, device=dev)
lengths = [60] * B


def run(name, fn, n=5):
    for rep in range(2):   # first repetition warms the per-lane handles
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            for k in range(K):
                M._LANE[0] = k
                with torch.cuda.stream(streams[k]):
                    fn(k * per, (k + 1) * per)
            M._LANE[0] = 0
        th = (time.perf_counter() - t0) / n * 1e3
        torch.cuda.synchronize()
        tt = (time.perf_counter() - t0) / n * 1e3
    print(f"{name:12s} K={K}: host enqueue {th:7.2f} ms, total {tt:7.2f} ms per round")


run("scene", lambda a, b: model._encode_scene(scene[a:b]))
run("vae.encode", lambda a, b: model.vae.encode(f_int[a:b], None, lengths[a:b], eps=eps[:, a:b].contiguous(), lengths_dev=len_dev[a:b]))
run("sampler", lambda a, b: model._diffusion_reverse(torch.cat([cond[a:b], cond[B + a:B + b]]), lengths[a:b], latents=xT[a:b]))
run("vae.decode", lambda a, b: model.vae.decode(z[:, a:b].contiguous(), lengths[a:b], T=60, lengths_dev=len_dev[a:b]))

# the whole per-lane pipeline, as MLD.ego_eval dispatches it
noise = {"eps_int": eps, "eps_unc": torch.randn(1, B, 256, device=dev), "x_T": xT}
def full(a, b):
    sub = tuple((x[a:b] if torch.is_tensor(x) else x) for x in batch)
    sn = {"eps_int": noise["eps_int"][:, a:b].contiguous(), "eps_unc": noise["eps_unc"][:, a:b].contiguous(), "x_T": noise["x_T"][a:b]}
    return model._ego_eval_one(sub, sn, defer_random=True, t_max=60, lengths=lengths[a:b])
run("full lane", full, n=3)
# host timeline of one round
torch.cuda.synchronize()
t0 = time.perf_counter()
marks = []
for k in range(K):
    M._LANE[0] = k
    with torch.cuda.stream(streams[k]):
        full(k * per, (k + 1) * per)
    marks.append((time.perf_counter() - t0) * 1e3)
M._LANE[0] = 0
torch.cuda.synchronize()
print("host time after enqueueing lane k:", [f"{m:.1f}" for m in marks], f"all done {(time.perf_counter() - t0) * 1e3:.1f} ms")
