"""seeme_b200 -- B200-native (sm_100a) implementation of SEE-ME's inference hot path:
scene / interactee-conditioned MLD latent-diffusion sampler -> motion-VAE decoder -> SMPL LBS.

Python is the host language (the reference is pure Python); all compute runs in hand-written CUDA
kernels behind the C ABI of ``include/seeme_b200.h`` (``seeme_b200/lib/libseeme_b200.so``).
"""
from __future__ import annotations

import os

__all__ = ["MLD", "load_config", "build_model", "CONFIG_DIR"]

# The batch pipeline (MLD.ego_eval_async) keeps up to 32 CUDA streams busy; with the driver's default of 8 hardware work queues
# streams alias onto the same queue and one slot's scene encoder waits behind another slot's sampler (16.3 k ->
# 20.2 k sequences/s).  The variable is read when the CUDA context is created, so it has to be set before the first CUDA call.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs")


def __getattr__(name):
    if name == "MLD":
        from .mld import MLD
        return MLD
    if name == "load_config":
        from .config import load_config
        return load_config
    raise AttributeError(name)


def build_model(config: str = "config_mld_egobody.yaml", device="cuda", guidance_scale=None, condition=None,
                max_batch: int = 8, n_points: int = 20000, seed: int = 0, datamodule=None, sparse_lbs: bool = True,
                overrides=None, **kw):
    """Synthetic-weights model of the named architecture (no checkpoints exist offline): the reference-keyed
    ``state_dict`` from ``seeme_b200.synthetic`` is loaded with ``load_state_dict`` exactly like ``test.py:111-113``."""
    import torch
    from . import synthetic
    from .config import load_config
    from .data import SyntheticDataModule
    from .mld import MLD
    path = config if os.path.isabs(config) else os.path.join(CONFIG_DIR, config)
    over = {"model": {}, "TEST": {"BATCH_SIZE": max_batch}}
    if guidance_scale is not None:
        over["model"]["guidance_scale"] = guidance_scale
    if condition is not None:
        over["model"]["condition"] = list(condition)
    for k, v in (overrides or {}).items():          # e.g. {"TEST": {"GLOBAL_ORIENT_PRED": False}} (one level of nesting)
        if isinstance(v, dict):
            over.setdefault(k, {}).update(v)
        else:
            over[k] = v
    cfg = load_config(path, overrides=over)
    dm = datamodule or SyntheticDataModule(cfg, name=cfg.DATASET_NAME, batch_size=max_batch, n_points=n_points,
                                           T=int(cfg.MOTION_LENGTH))
    model = MLD(cfg, dm, smpl_buffers=synthetic.smpl_buffers(sparse_lbs=sparse_lbs), max_batch=max_batch, max_points=n_points, **kw)
    sd = {}
    sd.update({"denoiser." + k: v for k, v in synthetic.denoiser_state(seed).items()})
    sd.update({"vae." + k: v for k, v in synthetic.vae_state(seed).items()})
    if "scene" in model.condition:
        sd.update({"proscene.scene_enc." + k: v for k, v in synthetic.pointnet_state(seed).items()})
        sd.update({"output_scene." + k: v for k, v in synthetic.output_scene_state(seed).items()})
    if "image" in model.condition:
        sd.update({"proscene.backbone." + k: v for k, v in synthetic.resnet50_state(seed).items()})
        sd.update({"output_images." + k: v for k, v in synthetic.output_images_state(seed).items()})
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if not m.startswith("smpl_model.")]
    if missing or unexpected:
        raise RuntimeError(f"state_dict mismatch: missing={missing[:5]} unexpected={unexpected[:5]}")
    return model.to(device)
