#!/usr/bin/env python
"""Headline benchmark: sampled SMPL sequences / second through the SEE-ME inference hot path
(scene + interactee conditioning -> 50-step DDIM with CFG -> VAE decode -> SMPL LBS, 6890 verts).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port)

A "step" is one pass of the hot path (``MLD.ego_eval``) over one batch of synthetic sequences:
BASELINE.json configs[1] = config_mld_egobody.yaml, batch 256, guidance 7.5, 50 DDIM steps, 20 000
scene points, 60 frames.  ``value`` is measured with the batch resident in HBM; ``e2e`` goes through
the same public call with HOST (pinned) buffers, the H2D copies of the batch and the D2H read of the
predicted joints inside the timed region.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os

# 32 hardware work queues for the pipeline's streams.  The CUDA context reads the variable when it is created,
# which torch.cuda.set_device / the NCCL initialisation in _init_dist() do -- before `import seeme_b200` (which sets the same
# default) in the gimo / interactee / smpl-sweep modes -- so it is set here, first thing.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SMPL sequences/sec (50-step DDIM+VAE+LBS)"
UNIT = "sequences/s"
GUIDANCE = 7.5
N_POINTS = 20000
# algorithmic (essential) work per unit, SURVEY 8(d) / App. E
POINTNET_FLOP_PER_POINT = 2 * 919_040          # fc_pos + 4 blocks with the pooled half hoisted
SMPL_BYTES_PER_FRAME = 82_680 + 340


def workload(batch):
    return {"workload": f"config_mld_egobody.yaml (BASELINE configs[1]): scene+interactee cond, batch {batch}/GPU, "
                        f"CFG {GUIDANCE}, 50 DDIM steps, {N_POINTS} scene points, 60 frames, VAE decode + SMPL LBS (6890 verts) "
                        "of the predicted body, joints for GT and interactee bodies",
            "batch_per_gpu": batch, "ddim_steps": 50, "guidance_scale": GUIDANCE, "scene_points": N_POINTS, "frames": 60,
            "l2_policy": "per-step inputs+activations (>= 1.3 GB per 128-cloud chunk) exceed the 126 MB L2; no flush needed",
            "inputs": "8 distinct synthetic batches and noise draws per rank, rotated step by step",
            "pipeline": "MLD.ego_eval_async with up to pipeline_depth (default 32) batches in flight, each on its own CUDA stream "
                        "(CUDA_DEVICE_MAX_CONNECTIONS=32) and kernel-side handles, 4 shared scene-encoder handles; sampler back-end "
                        "'auto': the one-CTA-per-tile kernel inside the pipeline (least SM time next to the scene encoder), the "
                        "persistent cluster kernel for a single batch (kernels.single_batch_* and kernels.sampler_*: the unpipelined "
                        "numbers)"}


class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region, read in-process through NVML on a background thread
    (spawning nvidia-smi inside the timed region stalls the driver for tens of milliseconds)."""

    def __init__(self, gpu_index: int, period_s: float = 0.02):
        import threading
        self.samples, self.reasons, self.err = [], set(), None
        self.max_mhz = None
        self._stop = threading.Event()
        self._on = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:          # noqa: BLE001
            self.err = f"NVML unavailable: {e}"
            self.nv = None
        self.period = period_s
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        if self.nv is None:
            return
        nv = self.nv
        bad = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
               "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            if self._on.is_set():
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for name, bit in bad.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception as e:   # noqa: BLE001
                    self.err = str(e)
            self._stop.wait(self.period)

    def begin(self):
        self._on.set()

    def stop(self):
        self._on.clear()
        self._stop.set()
        self.t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [self.err or "no samples"], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def oracle_weights():
    from seeme_b200 import synthetic as S
    return {"denoiser": S.denoiser_state(0), "vae": S.vae_state(0), "pointnet": S.pointnet_state(0),
            "output_scene": S.output_scene_state(0)}


def cpu_reference_step(W, smpl, stats, batch, noise):
    """the reference's CPU algorithm for the same path (oracle/restate.py: as-written math, fp32, all host threads)"""
    import torch
    from oracle import restate as O
    with torch.no_grad():
        return O.ego_eval(W, smpl, stats, batch, noise, guidance_scale=GUIDANCE, want_vertices=True)


def make_noise(B, seed=7):
    import torch
    g = torch.Generator().manual_seed(seed)
    return {"eps_int": torch.randn(1, B, 256, generator=g), "eps_unc": torch.randn(1, B, 256, generator=g),
            "x_T": torch.randn(B, 1, 256, generator=g)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  /root/reference cannot travel to
    the GPU box and is pure Python on uninstallable deps, so this arm times the oracle port of its algorithm
    (pinned against the unmodified reference modules in tests/) on the box's host cores."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from seeme_b200 import synthetic as S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.ref_batch
    W, smpl, stats = oracle_weights(), S.smpl_buffers(), S.norm_stats()
    batch = S.make_batch(B, n_points=N_POINTS)
    noise = make_noise(B)
    for _ in range(args.warmup):
        cpu_reference_step(W, smpl, stats, batch, noise)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(W, smpl, stats, batch, noise)
    dt = time.perf_counter() - t0
    v = B * args.steps / dt
    sample = f"{args.steps} steps of a batch of {B} sequences (same per-sequence workload as configs[1]; configs[0] batch), fp32, as-written math"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload(args.batch),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- shared helpers of the extra configurations ------------------------------------------------------------------
DENOISER_FLOP_PER_ROW_STEP = 2 * 5_520_384      # essential MACs per denoiser row and step
VAE_DECODE_FLOP_PER_SEQ = 2 * 125_400_000       # essential, T = 60
VAE_ENCODE_FLOP_PER_SEQ = 2 * 128_800_000


def _peaks():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(pk)) if os.path.exists(pk) else {}


def _init_dist():
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return torch, dist, rank, world, local, dev


def _barrier(torch, dist, world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _device_ms(torch, fn, n=3):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def _max_over_ranks(torch, dist, world, dev, ms):
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def _pipelined_seconds(torch, dist, world, dev, submit, steps, depth, on_result=None):
    """K submissions with `depth` in flight, device-timed on the caller's stream, max over ranks (see main.timed_pipelined)"""
    from collections import deque
    from seeme_b200 import _lib
    _barrier(torch, dist, world)
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pend, last = deque(), None
    for _ in range(steps):
        pend.append(submit())
        if len(pend) >= depth:
            last = pend.popleft()
            last.synchronize()
            if on_result:
                on_result(last)
    while pend:
        last = pend.popleft()
        last.synchronize()
        if on_result:
            on_result(last)
    torch.cuda.current_stream().wait_event(last.event)
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - n0
    sec = _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1)) / 1e3
    _barrier(torch, dist, world)
    return sec, launches


def class_rooflines(model, dev, dev_batch, B):
    """Per-kernel-class achieved / peak entries SURVEY 8(d) asks for beyond the dominant kernel: the sampler at the
    configuration's real row count AND at a saturating row count (148 row tiles), the VAE stacks and the VAE attention
    kernel.  Essential FLOP counts (App. E); CUDA-event timings of standalone calls on resident inputs."""
    import torch
    from seeme_b200 import _lib, ops, synthetic as S
    from seeme_b200.modules import time_sinusoid
    from seeme_b200.scheduler import DDIMScheduler
    peaks = _peaks()
    tpeak = peaks.get("bf16_tflops_sustained", 1400.0)
    out = {}
    sched = DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                          clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    sched.set_timesteps(50)
    ts, coef = sched.timesteps.tolist(), sched.step_coefficients()
    sd = {k: v.to(dev) for k, v in S.denoiser_state(0).items()}
    g = torch.Generator().manual_seed(3)

    def sampler_entry(Bs, backend, n):
        R = 2 * Bs
        op = ops.DenoiserOp(sd, max_rows=R)
        try:
            op.set_backend(backend)
            op.set_time_table(ts, time_sinusoid(sched.timesteps))
            cond = torch.randn(2, R, 256, generator=g).to(dev)
            xT = torch.randn(Bs, 256, generator=g).to(dev)
            op.sample(xT, cond, GUIDANCE, ts, coef)
            ms = _device_ms(torch, lambda: op.sample(xT, cond, GUIDANCE, ts, coef), n)
        finally:
            op.close()
        tf = DENOISER_FLOP_PER_ROW_STEP * R * 50 / (ms / 1e3) / 1e12
        return {"bound": "tensor (latency-bound below ~19k rows)", "rows": R, "ms_per_50_steps": ms, "achieved": tf, "peak": tpeak,
                "unit": "TFLOP/s (essential: 11.04 MFLOP per row-step; the kernels execute 3 split-bf16 passes)", "frac": tf / tpeak,
                "backend": backend}

    out["sampler_rows_512_persistent"] = sampler_entry(B, "persistent", 3)
    out["sampler_rows_512_graph"] = sampler_entry(B, "graph", 3)
    out["sampler_rows_18944_persistent"] = sampler_entry(9472, "persistent", 2)     # 148 row tiles of 128: one per SM-octet wave
    out["sampler_rows_18944_graph"] = sampler_entry(9472, "graph", 2)
    out["sampler_rows_18944_tile"] = sampler_entry(9472, "tile", 2)                 # one CTA per tile: 148 CTAs, one wave
    out["sampler_rows_512_tile"] = sampler_entry(B, "tile", 3)
    # VAE stacks at the batch size of the step
    vae = model.vae
    lengths = [60] * B
    z = torch.randn(1, B, 256, device=dev)
    feats = torch.randn(B, 60, 75, device=dev)
    vae.decode(z, lengths)
    for i in range(8):
        _lib.prof_read(i)
    _lib.prof_enable(True)
    ms_dec = _device_ms(torch, lambda: vae.decode(z, lengths), 3)
    _lib.prof_enable(False)
    attn_ms, attn_n = _lib.prof_read(4)
    vae.encode(feats, None, lengths)
    ms_enc = _device_ms(torch, lambda: vae.encode(feats, None, lengths), 3)
    for name, ms, fl in (("vae_decode", ms_dec, VAE_DECODE_FLOP_PER_SEQ), ("vae_encode", ms_enc, VAE_ENCODE_FLOP_PER_SEQ)):
        tf = fl * B / (ms / 1e3) / 1e12
        out[name] = {"bound": "tensor", "sequences": B, "ms": ms, "achieved": tf, "peak": tpeak, "unit": "TFLOP/s (essential)",
                     "frac": tf / tpeak}
    if attn_n:
        per = attn_ms / attn_n
        # mha1_umma_kernel streams qkv (fp32, 62 x 768 per sequence) in and the split-bf16 context (62 x 256 x 2 x 2 B) out in
        # lock-step phases: the L2 / HBM stream bounds it, not the two 64 x 64 x 256 contractions (1.07 GFLOP per launch)
        by = (62 * 768 * 4 + 62 * 256 * 4) * B
        out["vae_attention_kernel"] = {"bound": "hbm", "kernel": "mha1_umma_kernel (tcgen05, split fp16 x3)", "launches": attn_n,
                                       "avg_launch_ms": per, "achieved": by / (per / 1e3) / 1e9, "unit": "GB/s",
                                       "peak": peaks.get("hbm_gbs", 6650.0), "frac": by / (per / 1e3) / 1e9 / peaks.get("hbm_gbs", 6650.0)}
    return out


def stock_pytorch_gpu(dev, Bs=64):
    """The honest same-box bar: the as-written fp32 PyTorch math of the reference (oracle/restate.py, the
    checker -- executed here only as a measured baseline, like the cpu_baseline leg) in eager mode on the same B200."""
    import torch
    from oracle import restate as O
    from seeme_b200 import synthetic as S
    W = {k: {n: t.to(dev) for n, t in sd.items()} for k, sd in oracle_weights().items()}
    smpl = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in S.smpl_buffers().items()}
    stats = tuple(t.to(dev) for t in S.norm_stats())
    batch = tuple(x.to(dev) if torch.is_tensor(x) else x for x in S.make_batch(Bs, n_points=N_POINTS))
    noise = {k: v.to(dev) for k, v in make_noise(Bs).items()}

    def run():
        with torch.no_grad(), torch.device(dev):
            return O.ego_eval(W, smpl, stats, batch, noise, guidance_scale=GUIDANCE, want_vertices=True)

    run()
    ms = _device_ms(torch, run, 2)
    return {"value": Bs / (ms / 1e3), "unit": UNIT, "batch": Bs, "ms_per_batch": ms,
            "what": "oracle/restate.ego_eval (as-written fp32 math of the reference: two scene-encoder passes under CFG, ~540 ATen "
                    "ops per denoiser step, three skinned bodies) in eager PyTorch on cuda:0, TF32 off"}


def bench_gimo(args):
    """BASELINE configs[2]: config_mld_gimo.yaml (scene conditioning only, Nc = 1, 63-dim body pose), 512 sequences in total,
    split over the ranks (strong scaling by default), CFG 7.5, 50 steps, 20 000-point GIMO-shaped clouds."""
    torch, dist, rank, world, local, dev = _init_dist()
    import seeme_b200
    from seeme_b200 import synthetic as S
    scaling = args.scaling or "strong"
    total = int(args.total)
    B = total // world if scaling == "strong" else total
    model = seeme_b200.build_model("config_mld_gimo.yaml", device=dev, guidance_scale=GUIDANCE, max_batch=B, n_points=N_POINTS)
    model.prepare_pipeline()                      # all slots' kernel-side handles up front (start-up cost, not a per-step one)
    depth = max(1, int(model.pipeline_depth))
    N_ROT = 4
    hb, nh, db, nd = [], [], [], []
    for i in range(N_ROT):
        b = tuple(x.pin_memory() if torch.is_tensor(x) else x for x in S.make_batch(B, seed=4321 + 100 * rank + i, n_points=N_POINTS, dataset="gimo"))
        n = {"x_T": make_noise(B, 17 + 100 * rank + i)["x_T"].pin_memory()}
        hb.append(b); nh.append(n)
        db.append(tuple(x.to(dev) if torch.is_tensor(x) else x for x in b)); nd.append({k: v.to(dev) for k, v in n.items()})
    cnt = [0]

    def submit():
        cnt[0] += 1
        return model.ego_eval_async(db[cnt[0] % N_ROT], nd[cnt[0] % N_ROT])

    jh = [torch.empty(B, 60, 24, 3).pin_memory() for _ in range(depth)]

    def submit_e2e():
        cnt[0] += 1
        dst = jh[cnt[0] % depth]
        return model.ego_eval_async(hb[cnt[0] % N_ROT], nh[cnt[0] % N_ROT]).then(lambda rs: dst.copy_(rs["joints_rst"], non_blocking=True))

    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(3):
        model.ego_eval(db[0], nd[0])
    _pipelined_seconds(torch, dist, world, dev, submit, max(args.warmup, 3) + 2 * depth, depth)
    if clocks:
        clocks.begin()
    sec, launches = _pipelined_seconds(torch, dist, world, dev, submit, args.steps, depth)
    clk = clocks.stop() if clocks else None
    n_global = world * B
    value = n_global * args.steps / sec
    # one batch at a time: the latency that bounds strong scaling (the sampler chain does not get shorter with fewer rows)
    from seeme_b200 import _lib
    for i in range(8):
        _lib.prof_read(i)
    _lib.prof_enable(True)
    lat_ms = _max_over_ranks(torch, dist, world, dev, _device_ms(torch, lambda: model.ego_eval(db[1], nd[1]), 3))
    _lib.prof_enable(False)
    samp = _lib.prof_read(3)
    p6, p7 = _lib.prof_read(6), _lib.prof_read(7)
    _pipelined_seconds(torch, dist, world, dev, submit_e2e, max(args.warmup, 3) + 2 * depth, depth)      # same pattern as the timed loop
    sec_e, _ = _pipelined_seconds(torch, dist, world, dev, submit_e2e, args.steps, depth)
    if rank == 0:
        h2d = sum(x.numel() * x.element_size() for x in hb[0] if torch.is_tensor(x)) + nh[0]["x_T"].numel() * 4
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "fp16 (scene encoder) / split-bf16 x3 (denoiser, VAE) tensor-core operands, fp32 accumulation; f32 elsewhere",
                "data": "synthetic",
                "config": {"workload": f"config_mld_gimo.yaml (BASELINE configs[2]): scene conditioning only (Nc = 1), {n_global} sequences per "
                                       f"step over {world} GPU(s) ({B} per GPU, {scaling} scaling), CFG {GUIDANCE}, 50 DDIM steps, {N_POINTS}-point "
                                       "GIMO-shaped clouds, 60 frames, VAE decode + SMPL LBS of the predicted body",
                           "batch_per_gpu": B, "global_batch": n_global, "ddim_steps": 50, "guidance_scale": GUIDANCE,
                           "scene_points": N_POINTS, "frames": 60, "inputs": f"{N_ROT} distinct batches per rank, rotated",
                           "l2_policy": "per-step inputs+activations exceed the 126 MB L2; no flush needed"},
                "clocks": clk, "gpu_launches": int(launches),
                "e2e": {"value": n_global * args.steps / sec_e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(jh[0].numel() * 4)},
                "kernels": {"single_batch_latency_ms": lat_ms, "sampler_ms_per_run": samp[0] / max(samp[1], 1),
                            "scene_encoder_ms_per_batch": (p6[0] + p7[0]) / 3.0,
                            "limiter": "the 50-step sampler is a dependent chain whose length does not depend on the row count: with "
                                       "512 / N sequences per GPU its latency (sampler_ms_per_run) stays while the scene encoder / VAE / SMPL "
                                       "shrink with N -- strong-scaling efficiency follows single_batch_latency_ms / pipeline_depth"},
                "roofline": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_interactee(args):
    """BASELINE configs[3]: config_mld_interactee.yaml test protocol (ESTIMATE interactee, MOTION_LENGTH 1, guidance 1.0) with
    REPLICATION_TIMES repetitions per sequence, data-parallel over the ranks; a step = ONE test epoch of this rank's batches
    through MLD.run_test_batches (metric update per batch) followed by the NCCL all-reduce of the metric state and
    on_test_epoch_end -- the collective is INSIDE the timed region.  Scene embeddings are reused across repetitions."""
    torch, dist, rank, world, local, dev = _init_dist()
    import seeme_b200
    from seeme_b200 import dist as sdist
    from seeme_b200.data import SyntheticDataModule
    from seeme_b200.driver import _SceneEmbeddingCache, _scene_fingerprint
    Bb, n_batches = 64, 8                         # TEST.BATCH_SIZE 64 (config_mld_interactee.yaml), 8 batches = 512 sequences per rank
    model = seeme_b200.build_model("config_mld_interactee.yaml", device=dev, max_batch=Bb, n_points=N_POINTS)
    model.prepare_pipeline(n_batches)             # an epoch has n_batches batches in flight at most
    dm = SyntheticDataModule(model.cfg, name=model.name_dataset, batch_size=Bb, n_batches=n_batches * world, n_points=N_POINTS,
                             T=int(model.cfg.MOTION_LENGTH))
    lo, hi = sdist.shard_range(n_batches * world, rank, world)
    host = [tuple(x.pin_memory() if torch.is_tensor(x) else x for x in dm.batch(i)) for i in range(lo, hi)]
    reps = int(args.replications)

    def epoch(cache):
        def keyed():
            for i, b in enumerate(host):
                cache.key, cache.fp = i, _scene_fingerprint(b[4])
                yield b
        for _ in model.run_test_batches(keyed()):
            pass
        for m in model.metrics_dict:
            sdist.reduce_metric_state(getattr(model, m), device=dev)        # the epoch's collective (NCCL all-reduce)
        return model.on_test_epoch_end()

    def protocol(n_epochs):
        cache = _SceneEmbeddingCache(model)
        model._encode_scene = cache
        try:
            _barrier(torch, dist, world)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = [epoch(cache) for _ in range(n_epochs)]
            e1.record()
            torch.cuda.synchronize()
        finally:
            model._encode_scene = cache.orig
        return _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1)) / 1e3, cache, res

    clocks = ClockSampler(local) if rank == 0 else None
    protocol(max(2, min(args.warmup, 3)))          # warm-up: handles, graphs, allocator (its cache is discarded)
    if clocks:
        clocks.begin()
    from seeme_b200 import _lib
    n0 = _lib.launch_count()
    sec, cache, res = protocol(reps)
    launches = _lib.launch_count() - n0
    clk = clocks.stop() if clocks else None
    n_seq = len(host) * Bb * world
    if rank == 0:
        h2d = sum(x.numel() * x.element_size() for b in host for x in b if torch.is_tensor(x))
        line = {"metric": METRIC, "value": n_seq * reps / sec, "unit": UNIT, "n_gpus": world, "steps": reps, "warmup": max(2, min(args.warmup, 3)),
                "ms_per_step": sec / reps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "fp16 (scene encoder) / split-bf16 x3 (denoiser, VAE) tensor-core operands, fp32 accumulation; f32 elsewhere",
                "data": "synthetic",
                "config": {"workload": f"config_mld_interactee.yaml (BASELINE configs[3]): ESTIMATE interactee, MOTION_LENGTH 1, scene + interactee "
                                       f"conditioning, guidance 1.0, 50 DDIM steps, {N_POINTS} scene points; a step = one test epoch of "
                                       f"{len(host)} batches x {Bb} sequences per GPU from HOST (pinned) batches incl. metric update, NCCL "
                                       f"all-reduce of the metric state and on_test_epoch_end; {reps} repetitions (TEST.REPLICATION_TIMES), "
                                       "value counts sequence-repetitions",
                           "batch_per_gpu": Bb, "batches_per_gpu": len(host), "replications": reps,
                           "l2_policy": "an epoch streams 8 x 15 MB of clouds per rank through HBM; repetitions hit the scene-embedding cache"},
                "clocks": clk, "gpu_launches": int(launches),
                "e2e": {"value": n_seq * reps / sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 12 * 8,
                        "note": "the timed region IS end to end: host batches in, metric state out"},
                "kernels": {"scene_embedding_cache": {"hits": cache.hits, "misses": cache.misses,
                                                      "hit_rate": cache.hits / max(1, cache.hits + cache.misses)},
                            "metric_collective": "all_reduce(sum) of the 12-float EgoMetric state per epoch (NCCL), inside the timed region",
                            "last_epoch_metrics": {k: (None if v != v else float(v)) for k, v in res[-1].items()}},
                "roofline": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_smpl_sweep(args):
    """BASELINE configs[4]: standalone SMPL forward (Rodrigues + blendshapes + chain + LBS, 6890 vertices written), 1k-64k
    frames, every rank its own frames (weak).  HBM-bound class: 82 680 B written + 340 B read per frame."""
    torch, dist, rank, world, local, dev = _init_dist()
    from seeme_b200 import ops, synthetic as S
    peaks = _peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    sizes = [1024, 2048, 4096, 8192, 16384, 32768, 65536]
    op = ops.SmplOp({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in S.smpl_buffers().items()}, max_frames=max(sizes))
    g = torch.Generator().manual_seed(5 + rank)
    clocks = ClockSampler(local) if rank == 0 else None
    rows = []
    if clocks:
        clocks.begin()
    for F in sizes:
        betas = (0.5 * torch.randn(F, 10, generator=g)).to(dev)
        pose = (0.3 * torch.randn(F, 69, generator=g)).to(dev)
        go = (0.3 * torch.randn(F, 3, generator=g)).to(dev)
        tr = torch.randn(F, 3, generator=g).to(dev)
        for _ in range(max(args.warmup, 3)):
            op.forward(betas, pose, go, tr)
        _barrier(torch, dist, world)
        ms = _max_over_ranks(torch, dist, world, dev, _device_ms(torch, lambda: op.forward(betas, pose, go, tr), max(3, min(args.steps, 10))))
        fps = world * F / (ms / 1e3)
        gbs = SMPL_BYTES_PER_FRAME * F / (ms / 1e3) / 1e9
        rows.append({"frames_per_gpu": F, "ms": ms, "frames_per_s": fps, "achieved_gbs_per_gpu": gbs, "frac_of_hbm": gbs / hbm})
    clk = clocks.stop() if clocks else None
    if rank == 0:
        best = max(rows, key=lambda r: r["frames_per_s"])
        line = {"metric": "SMPL frames/sec (standalone forward + LBS, 6890 vertices)", "value": best["frames_per_s"], "unit": "frames/s",
                "n_gpus": world, "steps": max(3, min(args.steps, 10)), "warmup": max(args.warmup, 3), "ms_per_step": best["ms"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (split-fp16 x3 tensor-core blend contractions)",
                "data": "synthetic",
                "config": {"workload": "BASELINE configs[4]: SMPL forward/LBS sweep, 1k-64k frames per GPU, SMPL-shaped random body model, "
                                       "vertices + joints + quaternions written to HBM", "frames_per_gpu": best["frames_per_gpu"],
                           "l2_policy": "outputs of >= 8k frames (>= 0.68 GB) exceed the 126 MB L2; the 1k-4k rows are L2-resident and say so"},
                "clocks": clk, "gpu_launches": 2 * len(sizes) * max(3, min(args.steps, 10)),
                "roofline": {"bound": "hbm", "achieved": best["achieved_gbs_per_gpu"], "peak": hbm, "unit": "GB/s", "frac": best["frac_of_hbm"],
                             "traffic": None, "kernel": "smpl_skin_tc2_kernel (+ smpl_pose_kernel)",
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)"},
                "kernels": {"sweep": rows}, "e2e": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="seeme_b200", choices=["seeme_b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="sequences per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=8, help="batch of the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--lanes", type=int, default=None, help="concurrent sub-batches per step (default: the model's)")
    ap.add_argument("--config", default="egobody", choices=["egobody", "gimo", "interactee", "smpl-sweep"],
                    help="egobody = BASELINE configs[1] (the headline, default); gimo = configs[2] (scene-only, 512 sequences, "
                         "strong scaling); interactee = configs[3] (single-frame protocol, replications, NCCL metric gather "
                         "inside the timed epoch); smpl-sweep = configs[4] (standalone SMPL forward, 1k-64k frames)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"], help="default: strong for gimo, weak otherwise")
    ap.add_argument("--replications", type=int, default=10, help="interactee: TEST.REPLICATION_TIMES (line 71)")
    ap.add_argument("--total", type=int, default=512, help="gimo: sequences per step over all ranks")
    ap.add_argument("--no-extras", action="store_true", help="egobody: skip the per-class roofline / stock-PyTorch sub-measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != "egobody":
        return {"gimo": bench_gimo, "interactee": bench_interactee, "smpl-sweep": bench_smpl_sweep}[args.config](args)

    import torch
    import torch.distributed as dist
    import seeme_b200
    from seeme_b200 import _lib, synthetic as S
    from seeme_b200.metrics import STATE_KEYS

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    extra = {} if args.lanes is None else {"lanes": args.lanes}
    model = seeme_b200.build_model("config_mld_egobody.yaml", device=dev, guidance_scale=GUIDANCE, max_batch=B, n_points=N_POINTS, **extra)
    model.prepare_pipeline()                      # all slots' kernel-side handles up front (start-up cost, not a per-step one)
    # every rank owns its own sequences (weak scaling: the work list is sharded by sequence).  The timed loops
    # rotate through N_ROT distinct batches and noise draws (a per-input cache could not leak into the measurement)
    N_ROT = 8
    host_batches, noises_h, dev_batches, noises_d = [], [], [], []
    for i in range(N_ROT):
        hb = S.make_batch(B, seed=1234 + 100 * rank + i, n_points=N_POINTS)
        hb = tuple(x.pin_memory() if torch.is_tensor(x) else x for x in hb)
        nh = {k: v.pin_memory() for k, v in make_noise(B, 7 + 100 * rank + i).items()}
        host_batches.append(hb); noises_h.append(nh)
        dev_batches.append(tuple(x.to(dev) if torch.is_tensor(x) else x for x in hb))
        noises_d.append({k: v.to(dev) for k, v in nh.items()})
    host_batch, noise_h, dev_batch, noise_d = host_batches[0], noises_h[0], dev_batches[0], noises_d[0]
    h2d = sum(x.numel() * x.element_size() for x in host_batch if torch.is_tensor(x)) + sum(v.numel() * 4 for v in noise_h.values())

    from collections import deque
    depth = max(1, int(model.pipeline_depth))
    rot = [0]

    def step_resident():                      # synchronous public call (one batch at a time)
        rot[0] += 1
        return model.ego_eval(dev_batches[rot[0] % N_ROT], noises_d[rot[0] % N_ROT])

    def submit_resident():                    # asynchronous public call: up to `depth` batches in flight
        rot[0] += 1
        return model.ego_eval_async(dev_batches[rot[0] % N_ROT], noises_d[rot[0] % N_ROT])

    joints_host = [torch.empty(B, 60, 24, 3).pin_memory() for _ in range(depth)]
    e2e_count = [0]

    def submit_e2e():
        # HOST (pinned) batch in, joints out to pinned host memory, both on the slot's stream, every step
        dst = joints_host[e2e_count[0] % depth]
        i = e2e_count[0] % N_ROT
        p = model.ego_eval_async(host_batches[i], noises_h[i]).then(lambda rs: dst.copy_(rs["joints_rst"], non_blocking=True))
        e2e_count[0] += 1
        return p

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_pipelined(submit, steps, read_host=False):
        """K steps with `depth` batches in flight; device-timed on the caller's stream: every slot's stream starts after
        e0 (it waits for the caller's stream at submission) and e1 is recorded after waiting for every slot's last event"""
        barrier()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pend, sink = deque(), 0.0
        last = None
        for _ in range(steps):
            pend.append(submit())
            if len(pend) >= depth:
                last = pend.popleft()
                last.synchronize()            # the host consumes step k's result while steps k+1.. run
                if read_host:
                    sink += float(joints_host[0][0, 0, 0, 0])
        while pend:
            last = pend.popleft()
            last.synchronize()
        torch.cuda.current_stream().wait_event(last.event)
        e1.record()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - n0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms) / 1e3, launches, last.rs_set

    def timed(fn, steps, prof=False):
        barrier()
        if prof:
            for i in range(8):
                _lib.prof_read(i)
            _lib.prof_enable(True)
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            rs = fn()
        e1.record()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - n0
        if prof:
            _lib.prof_enable(False)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms) / 1e3, launches, rs

    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step_resident()
    # warm every pipeline slot (handles, sampler graphs) with the SAME submit / hold / synchronize pattern as the timed
    # loop: the results of `depth` steps are alive at once there, and a caching-allocator miss (cudaMalloc = device-wide
    # sync) inside the timed region showed up as 2x run-to-run variance
    timed_pipelined(submit_resident, max(args.warmup, 3) + 2 * depth)
    if clocks:
        clocks.begin()
    t_res, launches, rs = timed_pipelined(submit_resident, args.steps)
    clk = clocks.stop() if clocks else None
    value = world * B * args.steps / t_res
    # per-kernel-class device times and the single-batch latency: one batch at a time (no overlap between batches)
    # the pipelined region leaves the GPU at its power cap and a pause lets it fall to idle clocks: the latency-type numbers
    # below are taken after a few untimed steps and as the best of three groups (the profiled group is the last one)
    torch.cuda.synchronize()
    for _ in range(5):
        step_resident()
    t_groups = [timed(step_resident, min(args.steps, 5))[0] for _ in range(2)]
    t_single, _, _ = timed(step_resident, min(args.steps, 5), prof=True)
    t_single_best = min(t_groups + [t_single])
    single_steps = min(args.steps, 5)
    prof = {name: _lib.prof_read(i) for i, name in enumerate(["pointnet_gemm", "smpl_skin", "smpl_pose", "sampler_graph"])}
    p6, p7 = _lib.prof_read(6), _lib.prof_read(7)      # 6: residual blocks 1..3, 7: block 0 (+ fc_pos)
    prof["pointnet_fused"] = (p6[0] + p7[0], p6[1] + p7[1])
    prof["pointnet_block0"], prof["pointnet_blocks123"] = p7, p6

    # the per-epoch metric gather (the path's only collective): all-reduce(sum) of the EgoMetric state vector
    model.EgoMetric.update("test", rs["joints_rst"], rs["joints_ref"], rs["orientation_quat_rst"], rs["orientation_quat_ref"],
                           rs["root_interactee"], rs["joints_interactee"], rs["orientation_quat_int"], None, rs["lengths"], {})
    state = model.EgoMetric.state_vector().to(dev)
    if world > 1:
        dist.all_reduce(state)
    n_seq_metric = float(state[STATE_KEYS.index("count")]) / 60.0

    # SURVEY 8(d) sub-metrics, one extra pass each (device-timed, batch resident):
    #  (ii) the reference's own timing window (interactee encode + 50-step reverse + VAE decode, mld.py:1267-1368)
    #  (iii) the chain with the scene embeddings cached (replication protocol: repetitions of a sequence reuse them)
    #  (iv) SMPL alone (predicted-body skinning of B x 60 frames)
    def device_ms(fn, n=3):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    sub = {}
    orig_encode = model._encode_scene
    try:
        scene_dev = dev_batch[4]
        t_scene = device_ms(lambda: orig_encode(scene_dev))
        emb_cache = {}

        def cached_encode(scene):        # per (lane) sub-batch: the replication protocol reuses a sequence's scene embedding
            k = (scene.data_ptr(), tuple(scene.shape))
            if k not in emb_cache:
                emb_cache[k] = orig_encode(scene)
            return emb_cache[k]

        model._encode_scene = cached_encode
        # the three back-to-back encoder passes above leave the GPU at its power cap (~1.5 GHz); this chain is latency-bound
        # and would read 46 ms instead of 29 ms: let the clocks recover first
        torch.cuda.synchronize()
        time.sleep(1.0)
        step_fixed = lambda: model.ego_eval(dev_batch, noise_d)     # ONE batch: every pass hits the cached embedding
        step_fixed()
        torch.cuda.synchronize()
        t_cached = device_ms(step_fixed)
        model._encode_scene = orig_encode
        F = B * 60
        bet = torch.zeros(F, 10, device=dev); pz = torch.zeros(F, 69, device=dev); gz = torch.zeros(F, 3, device=dev)
        sop = model.smpl_model.op
        t_smpl = device_ms(lambda: sop.forward(bet, pz, gz, None))
        sub = {"scene_encoder_ms": t_scene, "chain_cached_scene_ms": t_cached,
               "sequences_per_s_cached_scene": world * B / (t_cached / 1e3), "smpl_only_ms": t_smpl,
               "smpl_only_frames_per_s": world * F / (t_smpl / 1e3)}
        # image backbone (not part of the benchmarked conditioning): ResNet-50 on 64 crops of 224 x 224
        from seeme_b200 import ops as _ops, synthetic as _S
        rop = _ops.ResNet50Op({k: v.to(dev) for k, v in _S.resnet50_state(0).items()},
                              {k: v.to(dev) for k, v in _S.output_images_state(0).items()}, max_batch=64)
        crops = torch.randn(64, 3, 224, 224, device=dev)
        rop(crops)
        t_img = device_ms(lambda: rop(crops))
        sub["image_backbone_ms_per_64_crops"] = t_img
        sub["image_backbone_crops_per_s"] = world * 64 / (t_img / 1e3)
        rop.close()
        del rop, crops
    except Exception as e:      # noqa: BLE001
        sub = {"error": str(e)}
    finally:
        model._encode_scene = orig_encode

    extras = {}
    if not args.no_extras and rank == 0:
        try:
            extras = class_rooflines(model, dev, dev_batch, B)
        except Exception as e:      # noqa: BLE001
            extras = {"error": f"{type(e).__name__}: {e}"}
        try:
            extras["stock_pytorch_gpu"] = stock_pytorch_gpu(dev)
        except Exception as e:      # noqa: BLE001
            extras["stock_pytorch_gpu"] = {"error": f"{type(e).__name__}: {e}"}
    barrier()

    e2e = None
    if not args.no_e2e:
        timed_pipelined(submit_e2e, max(args.warmup, 3) + 2 * depth, read_host=True)      # warm-up, same pattern as the timed loop
        t_e2e, _, _ = timed_pipelined(submit_e2e, args.steps, read_host=True)
        e2e = {"value": world * B * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(joints_host[0].numel() * 4)}

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        # dominant kernel class: the scene encoder's fused residual-block kernels (tensor-bound class).
        # Algorithmic work: SURVEY 8(d)/App. E essential count, 1.838 MFLOP per point for the whole encoder (fc_pos + 4
        # residual blocks with the pooled half of each concat hoisted), i.e. 0.4595 MFLOP per point per block launch;
        # one launch processes a chunk of up to 128 clouds x 20 000 points.
        pn_ms, pn_n = prof["pointnet_fused"]
        flops = POINTNET_FLOP_PER_POINT * N_POINTS * B * single_steps
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        ach = flops / (pn_ms / 1e3) / 1e12 if pn_ms > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r1_pointnet_kernels_final_ncu.json")
        if os.path.exists(tp):
            try:
                k0 = [k for k in json.load(open(tp)) if "pointnet_block_kernel" in k["Kernel Name"]][0]
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                traffic = sum(k0[m]["value"] * scale[k0[m]["unit"]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            except Exception:
                traffic = None
        roofline = {"kernel": "scene-encoder fused residual-block kernels pointnet_block0_tc_kernel + pointnet_block_kernel<true> "
                              "(tcgen05 fp16 x fp16 -> fp32, TMA, TMEM-resident hidden activation)",
                    "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": (ach / peak_tf) if ach else None,
                    "traffic": traffic,
                    "traffic_note": "dram read+write bytes of one pointnet_block_kernel launch (128 clouds x 20000 points, the bench's launch size) from "
                                    "profiles/r1_pointnet_kernels_final_ncu.json; algorithmic 2.62 GB (fp16 tile in + out)",
                    "algorithmic_flop_per_launch": POINTNET_FLOP_PER_POINT * N_POINTS * min(B, 128) / 4,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)",
                    "frac_of_burst_peak": (ach / peaks["bf16_tflops"]) if (ach and "bf16_tflops" in peaks) else None,
                    "launches": pn_n, "avg_launch_ms": pn_ms / pn_n if pn_n else None,
                    "share_of_step": (pn_ms / 1e3) / t_single if t_single else None,
                    "measured_in": "the single-batch (unpipelined) timed region, CUDA-event pairs on the launching stream"}
        sk_ms, sk_n = prof["smpl_skin"]
        hbm = peaks.get("hbm_gbs", 6650.0)
        sk_ach = SMPL_BYTES_PER_FRAME * 60 * B * single_steps / (sk_ms / 1e3) / 1e9 if sk_ms > 0 else None
        other = {"smpl_skin": {"bound": "hbm", "achieved": sk_ach, "peak": hbm, "unit": "GB/s", "frac": (sk_ach / hbm) if sk_ach else None,
                               "launches": sk_n, "avg_launch_ms": sk_ms / sk_n if sk_n else None},
                 # one 50-step sampler run of the batch (persistent cluster kernel in the single-batch region); the key keeps
                 # its round-1 name
                 "sampler_graph_ms_per_step": prof["sampler_graph"][0] / max(prof["sampler_graph"][1], 1),
                 "smpl_pose_ms_per_launch": prof["smpl_pose"][0] / max(prof["smpl_pose"][1], 1),
                 "pointnet_fused_ms": prof["pointnet_fused"][0], "pointnet_fused_launches": prof["pointnet_fused"][1],
                 "pointnet_block0_ms_per_launch": prof["pointnet_block0"][0] / max(prof["pointnet_block0"][1], 1),
                 "pointnet_blocks123_ms_per_launch": prof["pointnet_blocks123"][0] / max(prof["pointnet_blocks123"][1], 1),
                 "single_batch_latency_ms": t_single_best / single_steps * 1e3,
                 "single_batch_sequences_per_s": world * B * single_steps / t_single_best,
                 "single_batch_latency_ms_profiled_group": t_single / single_steps * 1e3,
                 "pipeline_depth": depth, "sub_metrics": sub}
        other.update(extras)
        cpu_baseline = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            Bc = args.ref_batch
            W, smpl, stats = oracle_weights(), S.smpl_buffers(), S.norm_stats()
            cb = S.make_batch(Bc, n_points=N_POINTS)
            cn = make_noise(Bc)
            cpu_reference_step(W, smpl, stats, cb, cn)           # warm-up pass (allocator, thread pool)
            t0 = time.perf_counter()
            passes = 0
            while True:                                           # bounded sample: about 10 s of CPU work
                cpu_reference_step(W, smpl, stats, cb, cn)
                passes += 1
                dt = time.perf_counter() - t0
                if dt >= 10.0 or passes >= 16:
                    break
            cpu_baseline = {"value": Bc * passes / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                            "sample": f"{passes} passes over a batch of {Bc} sequences of the same per-sequence workload ({dt:.1f} s), "
                                      "oracle/restate.py (as-written fp32 math of the reference, all host threads)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "fp16 (scene encoder) / split-bf16 x3 (denoiser, VAE) tensor-core operands, fp32 accumulation; f32 elsewhere",
                "data": "synthetic", "config": workload(B), "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "kernels": other, "cpu_baseline": cpu_baseline,
                "metric_gather": {"collective": "all_reduce(sum) of the EgoMetric state vector", "sequences": n_seq_metric}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
