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
B = int(os.environ.get("PROBE_B", "256"))
dev = torch.device("cuda", 0)
CFG = os.environ.get("PROBE_CONFIG", "config_mld_egobody.yaml")
model = seeme_b200.build_model(CFG, device=dev, guidance_scale=7.5, max_batch=B, n_points=20000, lanes=1)
batch = tuple(x.to(dev) if torch.is_tensor(x) else x for x in S.make_batch(B, n_points=20000, **({"dataset": "gimo"} if "gimo" in CFG else {})))
noise = {k: v.to(dev) for k, v in bench.make_noise(B).items()}
if os.environ.get("PROBE_CACHE_SCENE"):       # replication-protocol case: scene embeddings reused, no scene encoder in the loop
    _orig = model._encode_scene
    _emb = {}

    def _cached(scene):
        if 0 not in _emb:
            _emb[0] = _orig(scene)
        return _emb[0]

    model._encode_scene = _cached
skip = os.environ.get("PROBE_SKIP", "").split(",")      # ablations: which stage bounds the pipelined throughput
if "sampler" in skip:
    model._diffusion_reverse = lambda enc, lengths=None, latents=None: latents.permute(1, 0, 2).contiguous()
if "vaedec" in skip:
    _z75 = torch.zeros(B, 60, 75, device=dev)
    model.vae.decode = lambda z, lengths, T=None, lengths_dev=None: _z75
if "vaeenc" in skip:
    _ze = torch.zeros(1, B, 256, device=dev)
    model.vae.encode = lambda f, images=None, lengths=None, eps=None, lengths_dev=None: (_ze, None)
    model._encode_uncond = lambda *a, **k: _ze
if "smpl" in skip:
    _o = model._body
    _cache = {}

    def _body_cached(*a, **k):
        if 0 not in _cache:
            _cache[0] = _o(*a, **k)
        return _cache[0]

    model._body = _body_cached
streams = [torch.cuda.Stream() for _ in range(D)]
lengths = [60] * B


if os.environ.get("PROBE_ASYNC"):      # the public async API (slots, sampler_group) instead of raw lanes
    from collections import deque
    model.pipeline_depth = D
    if os.environ.get("PROBE_BACKEND"):
        model.sampler_backend = os.environ["PROBE_BACKEND"]
    if os.environ.get("PROBE_ENC_HANDLES"):
        model.encoder_handles = int(os.environ["PROBE_ENC_HANDLES"])
    n = int(os.environ.get("PROBE_STEPS", "48"))
    model.prepare_pipeline()

    host_ms = []

    def run(n):
        pend = deque()
        for _ in range(n):
            h0 = time.perf_counter()
            pend.append(model.ego_eval_async(batch, noise))
            host_ms.append((time.perf_counter() - h0) * 1e3)
            if len(pend) >= D:
                pend.popleft().synchronize()
        while pend:
            pend.popleft().synchronize()

    run(3 * D)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(n)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    hm = host_ms[-n:]
    print(f"async depth {D} B {B} mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB alloc / {(torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 2**30:.1f} GiB device: {dt:.2f} ms per step -> {B / dt * 1e3:.0f} sequences/s | host submit ms: "
          f"mean {sum(hm) / len(hm):.2f} max {max(hm):.2f} first8 {[round(x, 1) for x in hm[:8]]}")
    sys.exit(0)


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
n = int(os.environ.get("PROBE_STEPS", "24"))
for rep in range(int(os.environ.get("PROBE_REPS", "1"))):
    host = []
    t0 = time.perf_counter()
    for k in range(n):
        h0 = time.perf_counter()
        submit(k)
        host.append((time.perf_counter() - h0) * 1e3)
    t_sub = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    print(f"depth {D}: {dt:.2f} ms per step -> {B / dt * 1e3:.0f} sequences/s | host submit mean {sum(host) / n:.2f} max {max(host):.2f} ms, "
          f"all submitted after {t_sub:.0f} ms of {dt * n:.0f} ms | cpus {os.cpu_count()} load {os.getloadavg()[0]:.1f}")
