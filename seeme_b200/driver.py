"""Replication / test driver: the part of the reference's ``test.py`` around the hot path (``test.py:32-38,116-152``),
without Lightning.

``run_test_protocol`` runs ``cfg.TEST.REPLICATION_TIMES`` test epochs ("line-71 repetitions", config_mld_egobody.yaml:71);
one process per GPU, this rank's contiguous shard of the batches (``seeme_b200.dist``), ``model.pipeline_depth`` batches in
flight (``MLD.run_test_batches``), the scene embedding of a batch computed once and reused by the later repetitions (the
batches are identical across repetitions, only the sampling noise differs -- SURVEY 8e / App. H9), ONE all-reduce(sum) of
the metric state per epoch, and the mean / 95 % confidence interval / min / max table of ``get_metric_statistics``.

    torchrun --nproc-per-node 8 -m seeme_b200.driver --cfg config_mld_interactee.yaml --batches 8 --batch-size 64
"""
from __future__ import annotations

import json
from typing import Callable, Dict, Iterable, List, Optional

import numpy as np
import torch

from . import dist as sdist


def get_metric_statistics(values: np.ndarray, replication_times: int):
    """test.py:32-38"""
    mean = np.mean(values, axis=0)
    std = np.std(values, axis=0)
    conf_interval = 1.96 * std / np.sqrt(replication_times)
    return mean, conf_interval, np.min(values, axis=0), np.max(values, axis=0)


def summarize(all_metrics: Dict[str, List[float]], replication_times: int) -> Dict[str, object]:
    """test.py:138-148: per-metric mean / conf_interval / min / max, followed by the raw per-replication lists"""
    out: Dict[str, object] = {}
    for key, item in all_metrics.items():
        mean, ci, mn, mx = get_metric_statistics(np.array(item), replication_times)
        out[key + "/mean"], out[key + "/conf_interval"], out[key + "/min"], out[key + "/max"] = float(mean), float(ci), float(mn), float(mx)
    out.update(all_metrics)
    return out


def _scene_fingerprint(scene):
    """Cheap content check of a cloud batch.  Host tensors: exact bytes of a strided subset of the points (a re-drawn
    ``np.random.choice`` sample, as GIMO's loader does per item -- dataset.py:2018-2020 -- changes it); device tensors: the
    identity and version counter of the live tensor (the cache keeps a reference, so the address cannot be recycled)."""
    if scene.is_cuda:
        return ("cuda", scene.data_ptr(), scene._version)
    flat = scene.reshape(-1, scene.shape[-1])
    sub = flat[:: max(1, flat.shape[0] // 4096)]
    return ("cpu", sub.numpy().tobytes())


class _SceneEmbeddingCache:
    """Wraps ``MLD._encode_scene``: when the i-th batch of every epoch holds the same clouds, its embedding is computed in
    the first repetition only.  Keyed by the batch's position in the epoch (set by the driver before each submission) and
    verified against a content fingerprint: a batch whose clouds changed is re-encoded (counted as a miss)."""

    def __init__(self, model):
        self.model, self.orig, self.store, self.key = model, model._encode_scene, {}, None
        self.hits = self.misses = 0
        self.fp = None

    def __call__(self, scene):
        k = (self.key, tuple(scene.shape))
        fp = self.fp if self.fp is not None else _scene_fingerprint(scene)   # the driver fingerprints the batch as yielded
        ent = self.store.get(k)
        if ent is not None and ent[0] == fp:
            self.hits += 1
            torch.cuda.current_stream(ent[1].device).wait_event(ent[2])   # computed on another slot's stream
            return ent[1]
        self.misses += 1
        emb = self.orig(scene)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(emb.device))
        self.store[k] = (fp, emb, ev, scene if self.fp is None else None)
        return emb


def run_test_protocol(model, batches: Callable[[], Iterable], replication_times: Optional[int] = None,
                      cache_scene_embeddings: bool = False, out_json: Optional[str] = None) -> Dict[str, object]:
    """``batches()`` yields this rank's batches for one epoch (same order every call).  Returns the summary dict of
    ``summarize`` (identical on every rank: the metric state is all-reduced before ``compute``).

    ``cache_scene_embeddings`` is for datasets whose scenes are deterministic per item (EgoBody's preprocessed scene
    dict); GIMO re-draws its 20 000 points on every ``__getitem__`` (dataset.py:2018-2020), so its repetitions must
    re-encode -- the cache checks a content fingerprint and re-encodes on mismatch either way."""
    reps = int(replication_times if replication_times is not None else model.cfg.TEST.REPLICATION_TIMES)
    model.prepare_pipeline()          # the slots' kernel-side handles: built once here instead of inside the first epochs
    cache = _SceneEmbeddingCache(model) if cache_scene_embeddings and "scene" in model.condition else None
    if cache is not None:
        model._encode_scene = cache
    all_metrics: Dict[str, List[float]] = {}
    try:
        for rep in range(reps):
            def keyed():
                for i, b in enumerate(batches()):
                    if cache is not None:
                        cache.key = i
                        cache.fp = _scene_fingerprint(b[4]) if torch.is_tensor(b[4]) else None   # scene: item 4 of the tuple
                    yield b
            for _ in model.run_test_batches(keyed()):
                pass
            for m in model.metrics_dict:
                sdist.reduce_metric_state(getattr(model, m), device=next(model.parameters()).device)
            for key, val in model.on_test_epoch_end().items():
                all_metrics.setdefault(key, []).append(float(val))
    finally:
        if cache is not None:
            model._encode_scene = cache.orig
    summary = summarize(all_metrics, reps)
    if cache is not None:
        summary["_scene_embedding_cache"] = {"hits": cache.hits, "misses": cache.misses}
    if out_json and (not torch.distributed.is_initialized() or torch.distributed.get_rank() == 0):
        with open(out_json, "w", encoding="utf-8") as f:
            json.dump(summary, f, indent=4)
    return summary


def main():
    import argparse
    import os
    import seeme_b200
    from .data import SyntheticDataModule
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="config_mld_interactee.yaml")
    ap.add_argument("--batches", type=int, default=8, help="batches per epoch over ALL ranks")
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--points", type=int, default=20000)
    ap.add_argument("--replication-times", type=int, default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    model = seeme_b200.build_model(args.cfg, device=dev, max_batch=args.batch_size, n_points=args.points)
    dm = SyntheticDataModule(model.cfg, name=model.name_dataset, batch_size=args.batch_size, n_batches=args.batches,
                             n_points=args.points, T=int(model.cfg.MOTION_LENGTH))
    lo, hi = sdist.shard_range(args.batches, rank, world)       # all repetitions of a batch stay on its rank
    host = [tuple(x.pin_memory() if torch.is_tensor(x) else x for x in dm.batch(i)) for i in range(lo, hi)]
    summary = run_test_protocol(model, lambda: iter(host), args.replication_times, cache_scene_embeddings=True,
                                out_json=args.out)   # synthetic batches: identical clouds every epoch
    if rank == 0:
        print(json.dumps({k: v for k, v in summary.items() if not isinstance(v, list)}, indent=1))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
