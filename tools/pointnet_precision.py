"""Scene-embedding error of the tcgen05 scene encoder at each precision mode vs the fp32 CPU oracle,
and its effect on the final joints (run on the GPU box)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
mode = os.environ.get("SEEME_POINTNET_PRECISION", "3")
import seeme_b200
from oracle import restate as O
from seeme_b200 import ops, synthetic as S
dev = "cuda:0"
W = {"pointnet": S.pointnet_state(0), "output_scene": S.output_scene_state(0)}
p = S.egobody_scene(4, 20000, torch.Generator().manual_seed(3))
with torch.no_grad():
    ref = O.scene_embed(W["pointnet"], W["output_scene"], p)
op = ops.PointNetOp({k: v.to(dev) for k, v in W["pointnet"].items()}, {k: v.to(dev) for k, v in W["output_scene"].items()}, 4, 20000)
emb = op(p.to(dev)).cpu()
print(json.dumps({"precision": mode, "emb_absmax": float(ref.abs().max()), "emb_max_err": float((emb - ref).abs().max()),
                  "emb_rel_rms": float(((emb - ref).pow(2).mean() / ref.pow(2).mean()).sqrt())}))
