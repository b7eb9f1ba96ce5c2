"""MPJPE drift of the image-conditioned chain against the oracle for the backbone's operand formats
(SEEME_RESNET_PRECISION=3 split-bf16 / 16 fp16).  ``python tools/image_precision_probe.py`` on a GPU box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import seeme_b200  # noqa: E402
from oracle import restate as O  # noqa: E402  (a measurement tool, not the product path)
from seeme_b200 import synthetic as S  # noqa: E402

dev, B = "cuda:0", 4
cond = ("text", "image", "scene", "interactee")
feats_ref, transl, beta, utils_, scene, length, _ = S.make_batch(B, n_points=600, ragged=True)
batch = (feats_ref, transl, beta, utils_, scene, S.images(B, 3), length)
g = torch.Generator().manual_seed(5)
noise = {"eps_int": torch.randn(1, B, 256, generator=g), "x_T": torch.randn(B, 1, 256, generator=g)}
W = {"denoiser": S.denoiser_state(0), "vae": S.vae_state(0), "pointnet": S.pointnet_state(0), "output_scene": S.output_scene_state(0),
     "resnet50": S.resnet50_state(0), "output_images": S.output_images_state(0)}
with torch.no_grad():
    ref = O.ego_eval(W, S.smpl_buffers(), S.norm_stats(), batch, noise, condition=cond, guidance_scale=1.0)
for prec in ("3", "16"):
    os.environ["SEEME_RESNET_PRECISION"] = prec
    model = seeme_b200.build_model("config_mld_egobody.yaml", device=dev, guidance_scale=1.0, condition=cond, max_batch=B, n_points=600,
                                   scene_precision="split-bf16")
    rs = model.ego_eval(tuple(x.to(dev) for x in batch), {k: v.to(dev) for k, v in noise.items()})
    d = rs["joints_rst"].cpu() - ref["joints_rst"]
    print(f"backbone precision {prec}: MPJPE drift {d.norm(dim=-1).mean() * 1e3:.4f} mm, max-abs {d.abs().max() * 1e3:.4f} mm")
