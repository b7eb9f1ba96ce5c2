"""Seeded synthetic weights, SMPL-shaped body-model buffers and EgoBody/GIMO-shaped batches.

Checkpoints, datasets and the licensed SMPL model are unavailable offline (reference
``README.md:25-29``), so every test and benchmark runs on synthetic data of the exact shapes
and ``state_dict`` keys of the reference:

* denoiser keys/shapes  : ``mld/models/architectures/mld_denoiser.py:44-149`` + ``mdiff_transformer.py:137-284``
* VAE keys/shapes       : ``mld/models/architectures/mld_vae.py:51-114``
* scene encoder         : ``EgoHMR/models/respointnet.py:13-26,70-86``; ``mld/models/modeltype/mld.py:257-261``
* batch tuple           : ``mld/data/humanml/data/dataset.py:1778-1790`` (EgoBody), ``:2493-2501`` (GIMO)

Everything is generated on the CPU from a ``torch.Generator`` / ``numpy.RandomState`` so the same
seed gives bit-identical tensors in the build container and on the GPU box (tests compare against
golden outputs produced by the reference modules loaded with these same weights).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np
import torch

D = 256          # latent width (configs/config_mld_egobody.yaml:115)
NFEATS = 75      # 72 axis-angle + 3 translation (config_mld_egobody.yaml:123)
T_MAX = 60       # MOTION_LENGTH
N_POINTS = 20000
N_VERTS = 6890
N_JOINTS = 24
SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

BLOCKS = ["input_blocks.0", "input_blocks.1", "middle_block", "output_blocks.0", "output_blocks.1"]


# --------------------------------------------------------------------------------------------
# state_dict specs (key -> shape), in the reference's registration order where it matters little
# --------------------------------------------------------------------------------------------
def _lin(spec, name, out_f, in_f, bias=True):
    spec[name + ".weight"] = (out_f, in_f)
    if bias:
        spec[name + ".bias"] = (out_f,)


def _ln(spec, name, d=D):
    spec[name + ".weight"] = (d,)
    spec[name + ".bias"] = (d,)


def _mha(spec, name, d=D):
    spec[name + ".in_proj_weight"] = (3 * d, d)
    spec[name + ".in_proj_bias"] = (3 * d,)
    _lin(spec, name + ".out_proj", d, d)


def denoiser_spec() -> "OrderedDict[str, Tuple[int, ...]]":
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    _lin(s, "time_embedding.linear_1", D, D)
    _lin(s, "time_embedding.linear_2", D, D)
    s["query_pos.pe"] = (500, 1, D)
    s["mem_pos.pe"] = (500, 1, D)
    _ln(s, "encoder.norm")
    for b in BLOCKS:
        p = f"encoder.{b}."
        _ln(s, p + "ca_block.norm")
        _ln(s, p + "ca_block.text_norm")
        _lin(s, p + "ca_block.query", D, D)
        _lin(s, p + "ca_block.key", D, D)
        _lin(s, p + "ca_block.value", D, D)
        _lin(s, p + "ca_block.proj_out.emb_layers.1", 2 * D, D)
        _ln(s, p + "ca_block.proj_out.norm")
        _lin(s, p + "ca_block.proj_out.out_layers.2", D, D)
        _lin(s, p + "ffn.linear1", 128, D)
        _lin(s, p + "ffn.linear2", D, 128)
        _lin(s, p + "ffn.proj_out.emb_layers.1", 2 * D, D)
        _ln(s, p + "ffn.proj_out.norm")
        _lin(s, p + "ffn.proj_out.out_layers.2", D, D)
        _mha(s, p + "sa_block.self_attn")
        _lin(s, p + "sa_block.linear1", 1024, D)
        _lin(s, p + "sa_block.linear2", D, 1024)
        _ln(s, p + "sa_block.norm1")
        _ln(s, p + "sa_block.norm2")
    for i in range(2):
        _lin(s, f"encoder.linear_blocks.{i}", D, 2 * D)
    return s


def vae_spec(nfeats: int = NFEATS) -> "OrderedDict[str, Tuple[int, ...]]":
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["global_motion_token"] = (2, D)
    s["query_pos_encoder.pe"] = (500, 1, D)
    s["query_pos_decoder.pe"] = (500, 1, D)
    for stack, has_cross in (("encoder", False), ("decoder", True)):
        for b in BLOCKS:
            p = f"{stack}.{b}."
            _mha(s, p + "self_attn")
            if has_cross:
                _mha(s, p + "multihead_attn")
            _lin(s, p + "linear1", 128, D)
            _lin(s, p + "linear2", D, 128)
            _ln(s, p + "norm1")
            _ln(s, p + "norm2")
            if has_cross:
                _ln(s, p + "norm3")
        for i in range(2):
            _lin(s, f"{stack}.linear_blocks.{i}", D, 2 * D)
        _ln(s, f"{stack}.norm")
    _lin(s, "skel_embedding", D, nfeats)
    _lin(s, "final_layer", nfeats, D)
    return s


def pointnet_spec() -> "OrderedDict[str, Tuple[int, ...]]":
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    _lin(s, "fc_pos_0", 512, 3)
    for i in range(4):
        _lin(s, f"block_{i}.fc_0", 256, 512)
        _lin(s, f"block_{i}.fc_1", 256, 256)
        _lin(s, f"block_{i}.shortcut", 256, 512, bias=False)
    _lin(s, "fc_c", 512, 256)
    return s


def output_scene_spec() -> "OrderedDict[str, Tuple[int, ...]]":
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    _lin(s, "1", D, 512)
    return s


def _fill(spec, seed: int, style: str) -> "OrderedDict[str, torch.Tensor]":
    """style 'xavier': transformer stacks after ``_reset_parameters`` (cross_attention.py:36-39);
    style 'default': ``nn.Linear`` default init (U(+-1/sqrt(fan_in))).  Biases and LayerNorm affine
    terms are given small non-trivial values so every code path is exercised by parity tests
    (a trained checkpoint has non-zero values there; the reference's zero-inits would hide bugs)."""
    g = torch.Generator().manual_seed(seed)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def u(shape, b):
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * b

    for k, shape in spec.items():
        if k.endswith(".pe"):
            out[k] = torch.rand(shape, generator=g, dtype=torch.float32)         # nn.init.uniform_
        elif k == "global_motion_token":
            out[k] = torch.randn(shape, generator=g, dtype=torch.float32)
        elif "." in k and "norm" in k.split(".")[-2]:          # LayerNorm affine terms
            if k.endswith("weight"):
                out[k] = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
            else:
                out[k] = 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        elif len(shape) >= 2:
            fan_out, fan_in = shape[0], shape[1]
            if style == "xavier":
                out[k] = u(shape, math.sqrt(6.0 / (fan_in + fan_out)))
            else:
                out[k] = u(shape, 1.0 / math.sqrt(fan_in))
        else:  # biases
            out[k] = u(shape, 0.05)
    return out


def denoiser_state(seed: int = 0):
    return _fill(denoiser_spec(), 1000 + seed, "xavier")


def vae_state(seed: int = 0, nfeats: int = NFEATS):
    return _fill(vae_spec(nfeats), 2000 + seed, "xavier")


def pointnet_state(seed: int = 0):
    return _fill(pointnet_spec(), 3000 + seed, "default")


def output_scene_state(seed: int = 0):
    return _fill(output_scene_spec(), 4000 + seed, "default")


# --------------------------------------------------------------------------------------------
# image backbone: ResNet-50 state_dict keys of EgoHMR/models/resnet.py:99-150
# --------------------------------------------------------------------------------------------
RESNET50_LAYERS = (3, 4, 6, 3)
_BN_FIELDS = ("weight", "bias", "running_mean", "running_var")


def resnet50_convs() -> List[Tuple[str, str, int, int, int]]:
    """(conv key prefix, bn key prefix, Cout, Cin, kernel) in the tensor order of ``seeme_resnet50_create``"""
    out = [("conv1", "bn1", 64, 3, 7)]
    cin = 64
    for L, n in enumerate(RESNET50_LAYERS):
        planes = 64 << L
        for b in range(n):
            p = f"layer{L + 1}.{b}."
            out += [(p + "conv1", p + "bn1", planes, cin, 1), (p + "conv2", p + "bn2", planes, planes, 3),
                    (p + "conv3", p + "bn3", planes * 4, planes, 1)]
            if b == 0:
                out.append((p + "downsample.0", p + "downsample.1", planes * 4, cin, 1))
            cin = planes * 4
    return out


def resnet50_state(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Convolutions as the reference initialises them (N(0, sqrt(2 / (k*k*Cout))), resnet.py:115-118); BatchNorm
    affine terms and running statistics get non-trivial values (a trained checkpoint has them; the reference's
    1/0/0/1 init would make the BatchNorm fold untestable)."""
    g = torch.Generator().manual_seed(5000 + seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for conv, bn, cout, cin, k in resnet50_convs():
        sd[conv + ".weight"] = torch.randn((cout, cin, k, k), generator=g) * math.sqrt(2.0 / (k * k * cout))
        last = conv.endswith("conv3")
        lo, hi = (0.2, 0.5) if last else (0.7, 1.1)
        sd[bn + ".weight"] = lo + (hi - lo) * torch.rand((cout,), generator=g)
        sd[bn + ".bias"] = 0.1 * torch.randn((cout,), generator=g)
        sd[bn + ".running_mean"] = 0.1 * torch.randn((cout,), generator=g)
        sd[bn + ".running_var"] = 0.5 + torch.rand((cout,), generator=g)
        sd[bn + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return sd


def output_images_state(seed: int = 0):
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    _lin(s, "1", D, 2048)
    return _fill(s, 6000 + seed, "default")


def images(batch: int, seed: int = 0) -> torch.Tensor:
    """[B,3,224,224] normalised crops (the reference feeds ImageNet-normalised 224x224 crops)"""
    g = torch.Generator().manual_seed(7000 + seed)
    return torch.randn((batch, 3, 224, 224), generator=g)


# --------------------------------------------------------------------------------------------
# SMPL-shaped body model buffers (smplx==0.1.28 buffer names)
# --------------------------------------------------------------------------------------------
_REST_JOINTS = np.array([
    [0.00, 0.00, 0.00], [0.07, -0.09, 0.00], [-0.07, -0.09, 0.00], [0.00, 0.11, -0.02],
    [0.10, -0.47, 0.00], [-0.10, -0.47, 0.00], [0.00, 0.25, 0.00], [0.09, -0.87, -0.03],
    [-0.09, -0.87, -0.03], [0.00, 0.30, 0.02], [0.11, -0.93, 0.09], [-0.11, -0.93, 0.09],
    [0.00, 0.52, -0.02], [0.08, 0.42, -0.01], [-0.08, 0.42, -0.01], [0.00, 0.60, 0.03],
    [0.17, 0.44, -0.02], [-0.17, 0.44, -0.02], [0.43, 0.43, -0.03], [-0.43, 0.43, -0.03],
    [0.68, 0.43, -0.03], [-0.68, 0.43, -0.03], [0.76, 0.42, -0.03], [-0.76, 0.42, -0.03]],
    dtype=np.float64)


def smpl_buffers(seed: int = 1, sparse_lbs: bool = True) -> Dict[str, torch.Tensor]:
    """A 1.7 m humanoid with SMPL's buffer shapes: vertices scattered around the 23 bones,
    row-stochastic ``J_regressor`` (<=32 nnz/row) and ``lbs_weights`` (<=4 nnz/row like the real
    model, or dense when ``sparse_lbs=False`` to exercise the dense fallback)."""
    rs = np.random.RandomState(seed)
    V, J = N_VERTS, N_JOINTS
    parents = np.array(SMPL_PARENTS)
    # vertices: pick a bone (child joint c>0), a position along it and a radial offset
    bone = rs.randint(1, J, size=V)
    bone = np.sort(bone)                       # real SMPL vertex order has strong part locality
    tpar = rs.rand(V, 1)
    a, b = _REST_JOINTS[parents[bone]], _REST_JOINTS[bone]
    center = a * (1 - tpar) + b * tpar
    off = rs.randn(V, 3)
    off /= np.linalg.norm(off, axis=1, keepdims=True) + 1e-9
    v_template = center + off * (0.03 + 0.05 * rs.rand(V, 1))
    # skinning weights: bone's child joint, its parent, and up to two random neighbours in the tree
    W = np.zeros((V, J))
    if sparse_lbs:
        for v in range(V):
            c = bone[v]
            cand = [c, parents[c]]
            gp = parents[parents[c]]
            if gp >= 0:
                cand.append(gp)
            kids = np.nonzero(parents == c)[0]
            if len(kids):
                cand.append(kids[rs.randint(len(kids))])
            cand = list(dict.fromkeys(cand))[:4]
            w = rs.dirichlet(np.ones(len(cand)) * 0.7)
            W[v, cand] = w
    else:
        W = rs.dirichlet(np.ones(J) * 0.3, size=V)
    # joint regressor: 32 vertices of the joint's own bone(s), random convex weights
    Jreg = np.zeros((J, V))
    for j in range(J):
        d = np.linalg.norm(v_template - _REST_JOINTS[j], axis=1)
        idx = np.argsort(d)[:32]
        Jreg[j, idx] = rs.dirichlet(np.ones(32))
    shapedirs = 0.01 * rs.randn(V, 3, 10)
    posedirs = 0.001 * rs.randn(207, V * 3)
    f32 = lambda x: torch.from_numpy(np.ascontiguousarray(x)).float()
    return {
        "v_template": f32(v_template), "shapedirs": f32(shapedirs), "posedirs": f32(posedirs),
        "J_regressor": f32(Jreg), "lbs_weights": f32(W),
        "parents": torch.tensor(SMPL_PARENTS, dtype=torch.long),
        "faces_tensor": torch.from_numpy(rs.randint(0, V, size=(13776, 3))).long(),
    }


# --------------------------------------------------------------------------------------------
# batches and dataset statistics
# --------------------------------------------------------------------------------------------
def norm_stats(nfeats: int = NFEATS + 3) -> Tuple[torch.Tensor, torch.Tensor]:
    """float64 ``[1,78]`` mean/std like the npy stats of the datamodule (``mld/data/EgoBody.py:151-157``
    promotes to float64).  mean 0; std 0.2 on the 72 pose dims, 0.5 on translation."""
    mean = torch.zeros(1, nfeats, dtype=torch.float64)
    std = torch.full((1, nfeats), 0.5, dtype=torch.float64)
    std[0, :72] = 0.2
    return mean, std


def egobody_scene(B: int, n_points: int, g: torch.Generator) -> torch.Tensor:
    xy = torch.rand(B, n_points, 2, generator=g) * 6.0 - 3.0
    z = torch.rand(B, n_points, 1, generator=g) * 5.7 + 0.3
    return torch.cat([xy, z], dim=-1)


def gimo_scene(B: int, n_points: int, g: torch.Generator) -> torch.Tensor:
    """Points sampled with replacement from a synthetic 8x3x8 m room (floor + 4 walls + boxes), scaled
    1/1.03 and passed through a random rigid transform (mirrors ``dataset.py:1994-2026``)."""
    n_room = 200_000
    face = torch.randint(0, 6, (n_room,), generator=g)
    uvw = torch.rand(n_room, 3, generator=g)
    p = uvw * torch.tensor([8.0, 3.0, 8.0])
    p[face == 0, 1] = 0.0
    p[face == 1, 0] = 0.0
    p[face == 2, 0] = 8.0
    p[face == 3, 2] = 0.0
    p[face == 4, 2] = 8.0
    box = face == 5
    p[box] = p[box] * torch.tensor([0.15, 0.3, 0.15]) + torch.tensor([3.0, 0.0, 4.0])
    p = (p - torch.tensor([4.0, 0.0, 4.0])) / 1.03
    out = torch.empty(B, n_points, 3)
    for b in range(B):
        idx = torch.randint(0, n_room, (n_points,), generator=g)
        ang = float(torch.rand(1, generator=g)) * 2 * math.pi
        c, s = math.cos(ang), math.sin(ang)
        R = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
        tr = torch.randn(3, generator=g) * 0.5
        out[b] = p[idx] @ R.T + tr
    return out


def make_batch(B: int, seed: int = 1234, n_points: int = N_POINTS, T: int = T_MAX,
               ragged: bool = False, dataset: str = "egobody", with_images: bool = False):
    """The 7-tuple ``ego_eval`` unpacks at ``mld/models/modeltype/mld.py:1135-1137``:
    (feats_ref[B,T,2,72], transl[B,2,T,3], beta[B,2,T,10], utils_[B,T,6], scene[B,N,3],
    length[B,1] int32, dict_images: T tuples of B strings); ``with_images``: the image-conditioned item tuple
    (..., scene, images[B,3,224,224], length) of ``dataset.py:1788-1790``."""
    g = torch.Generator().manual_seed(seed)
    # GIMO rows carry 21 body joints: 3 + 63 = 66 pose dims (mld.py:1656, Gimo.py numdims 69 with transl)
    feats_ref = torch.randn(B, T, 2, 72 if dataset == "egobody" else 66, generator=g)
    transl = torch.randn(B, 2, T, 3, generator=g)
    beta = (0.5 * torch.randn(B, 2, 1, 10, generator=g)).expand(B, 2, T, 10).contiguous()
    utils_ = torch.rand(B, T, 6, generator=g)
    scene = egobody_scene(B, n_points, g) if dataset == "egobody" else gimo_scene(B, n_points, g)
    if ragged:
        length = torch.randint(20, T + 1, (B, 1), generator=g, dtype=torch.int32)
        length[0, 0] = T   # the reference sizes the decode by max(lengths)
    else:
        length = torch.full((B, 1), T, dtype=torch.int32)
    if with_images:
        return feats_ref, transl, beta, utils_, scene, images(B, seed), length
    dict_images = [tuple(f"img_{t:03d}_{b:04d}.jpg" for b in range(B)) for t in range(T)]
    return feats_ref, transl, beta, utils_, scene, length, dict_images
