"""Per-epoch wall time of the interactee replication protocol (bench.py --config interactee) to see cold-slot effects.
    python tools/interactee_probe.py [epochs]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seeme_b200  # noqa: E402
from seeme_b200.data import SyntheticDataModule  # noqa: E402
from seeme_b200.driver import _SceneEmbeddingCache, _scene_fingerprint  # noqa: E402

n_ep = int(sys.argv[1]) if len(sys.argv) > 1 else 14
dev = torch.device("cuda", 0)
Bb, nb = 64, 8
model = seeme_b200.build_model("config_mld_interactee.yaml", device=dev, max_batch=Bb, n_points=20000)
dm = SyntheticDataModule(model.cfg, name=model.name_dataset, batch_size=Bb, n_batches=nb, n_points=20000, T=int(model.cfg.MOTION_LENGTH))
host = [tuple(x.pin_memory() if torch.is_tensor(x) else x for x in dm.batch(i)) for i in range(nb)]
model.prepare_pipeline(nb)
cache = _SceneEmbeddingCache(model)
model._encode_scene = cache
ts = []
from seeme_b200 import ops as _ops, _lib  # noqa: E402
chosen = []
_orig_sb = _ops.DenoiserOp.set_backend


def _sb(self, name):
    chosen.append(name[0])
    return _orig_sb(self, name)


_ops.DenoiserOp.set_backend = _sb
from seeme_b200 import modules as _mods  # noqa: E402
created = []
_orig_hi = _ops._Handle.__init__


def _hi(self, *a, **k):
    created.append((type(self).__name__, _mods._LANE[0]))
    return _orig_hi(self, *a, **k)


_ops._Handle.__init__ = _hi
info = []
for e in range(n_ep):
    chosen.clear()
    created.clear()
    a0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
    l0 = _lib.launch_count()
    torch.cuda.synchronize()
    t0 = time.perf_counter()

    def keyed():
        for i, b in enumerate(host):
            cache.key, cache.fp = i, _scene_fingerprint(b[4])
            yield b
    for _ in model.run_test_batches(keyed()):
        pass
    model.on_test_epoch_end()
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
    info.append(("".join(chosen), torch.cuda.memory_stats().get("num_device_alloc", 0) - a0, _lib.launch_count() - l0, list(created)))
print(f"depth {model.pipeline_depth} backend {model.sampler_backend} connections {os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS')}: epoch ms",
      [round(t, 1) for t in ts], "slots used", len(model.__dict__.get("_slot_streams", {})), "| per epoch (backends, cudaMallocs, launches):", info)
