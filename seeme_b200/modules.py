"""Host-side mirrors of the reference's operator classes on the hot path.

Each class keeps the reference's constructor signature, ``state_dict`` keys (so reference checkpoints
load with ``load_state_dict``) and call surface, but its ``forward`` runs the sm_100a kernels through
the C ABI (``seeme_b200.ops``).  Parameters are plain ``nn.Parameter``s holding fp32 weights; the
kernels' packed copies are (re)built lazily whenever the parameters change (``_version`` bump) or
move device.  Unsupported ablations raise, as the reference does for its own unsupported
combinations (``mld_denoiser.py:96,149,190``) -- nothing falls back to PyTorch math.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops, synthetic


def _register_spec(root: nn.Module, spec: Dict[str, tuple]) -> None:
    """Create nested containers so that ``root.state_dict()`` has exactly the dotted keys of ``spec``."""
    for key, shape in spec.items():
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape), requires_grad=False))


# Execution lane (see MLD.ego_eval): kernel-side handles own their workspace and are not re-entrant, so every lane --
# a sub-batch running on its own CUDA stream -- gets its own handle.  Set by ``MLD`` around a lane's calls.
_LANE = [0]


class _PackedModule(nn.Module):
    """Tracks parameter versions/devices so the C-side packed weights are rebuilt when they change."""

    def _signature(self):
        # nn.Module.parameters() walks the module tree with de-duplication (~0.5 ms per call for the denoiser, three calls per
        # batch): the (owning module, name) slots are cached instead and the CURRENT parameter object is read from each slot, so
        # in-place updates (version counter), replaced parameter objects and moved devices all change the signature
        slots = self.__dict__.get("_pslots")
        if slots is None:
            slots = self.__dict__["_pslots"] = [(m._parameters, n) for m in self.modules() for n in m._parameters]
        sig = []
        for d, n in slots:
            p = d[n]
            sig.append((p.data_ptr(), p._version, p.device.index) if p is not None else None)
        return tuple(sig)

    def _op(self, key, factory):
        key = (key, _LANE[0])
        cache = self.__dict__.setdefault("_op_cache", {})
        sig = self._signature()
        ent = cache.get(key)
        if ent is None or ent[0] != sig:
            if ent is not None:
                ent[1].close()
            cache[key] = (sig, factory())
        return cache[key][1]

    def _require_cuda(self):
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError(f"{type(self).__name__}: parameters are on {p.device}; seeme_b200 runs on CUDA (sm_100a) only")
        return p.device


def _attr(obj, name, default=None):
    if isinstance(obj, dict):
        return obj.get(name, default)
    return getattr(obj, name, default)


class MldDenoiser(_PackedModule):
    """``mld.models.architectures.mld_denoiser.MldDenoiser`` for the north-star configuration:
    trans_enc + SKIP_CONNECT + MD_TRANS + learned PE, 5 blocks, 1 head, latent [1,256]."""

    def __init__(self, ablation, nfeats: int = 72, condition="text", latent_dim: list = [1, 256], ff_size: int = 128,
                 num_layers: int = 6, num_heads: int = 4, dropout: float = 0.1, normalize_before: bool = False,
                 activation: str = "gelu", flip_sin_to_cos: bool = True, return_intermediate_dec: bool = False,
                 position_embedding: str = "learned", arch: str = "trans_enc", freq_shift: int = 0,
                 guidance_scale: float = 7.5, guidance_uncondp: float = 0.1, text_encoded_dim: int = 256,
                 nclasses: int = 10, max_rows: int = 1024, **kwargs) -> None:
        super().__init__()
        self.latent_dim = latent_dim[-1]
        self.text_encoded_dim = text_encoded_dim
        self.condition = condition
        self.arch = arch
        self.MD_trans = _attr(ablation, "MD_TRANS", False)
        self.ablation_skip_connection = _attr(ablation, "SKIP_CONNECT", False)
        self.pe_type = _attr(ablation, "DIFF_PE_TYPE", "mld")
        self.max_rows = max_rows
        if "text" not in self.condition:
            raise TypeError(f"condition type {self.condition} not supported")          # mld_denoiser.py:190
        unsupported = []
        if not (self.MD_trans and self.ablation_skip_connection and arch == "trans_enc"):
            unsupported.append("only arch=trans_enc with SKIP_CONNECT and MD_TRANS is built")
        if self.pe_type != "mld" or position_embedding not in ("learned", "v3"):
            unsupported.append("only DIFF_PE_TYPE=mld with learned position embedding is built")
        if _attr(ablation, "VAE_TYPE", "actor") == "no":
            unsupported.append("diffusion-only (VAE_TYPE=no) is not built")
        if list(latent_dim) != [1, 256] or text_encoded_dim != 256 or num_layers != 5 or num_heads != 1 or ff_size != 128:
            unsupported.append("kernels are specialised for latent_dim [1,256], text_encoded_dim 256, 5 layers, 1 head, ff 128")
        if not flip_sin_to_cos or freq_shift != 0:
            unsupported.append("timestep embedding: flip_sin_to_cos=True, freq_shift=0 only")
        if unsupported:
            raise NotImplementedError("MldDenoiser (sm_100a): " + "; ".join(unsupported))
        _register_spec(self, synthetic.denoiser_spec())

    @property
    def op(self) -> ops.DenoiserOp:
        self._require_cuda()
        return self._op("den", lambda: ops.DenoiserOp(self.state_dict(), self.max_rows))

    def op_rows(self, rows: int) -> ops.DenoiserOp:
        """a handle of the current lane with room for ``rows`` denoiser rows (the coalesced sampler of several batches)"""
        self._require_cuda()
        rows = max(int(rows), 1)
        return self._op(("den", rows), lambda: ops.DenoiserOp(self.state_dict(), rows))

    def forward(self, sample, timestep, encoder_hidden_states, lengths=None, **kwargs):
        """sample [B',1,256], timestep 0-d tensor/int, encoder_hidden_states [Nc,B',256] -> ([B',1,256],)"""
        if sample.dim() != 3 or sample.shape[1] != 1 or sample.shape[2] != 256:
            raise ValueError(f"sample must be [B,1,256], got {tuple(sample.shape)}")
        R = sample.shape[0]
        if encoder_hidden_states.dim() != 3 or encoder_hidden_states.shape[1] != R:
            raise ValueError("encoder_hidden_states must be [Nc,B,256] with the batch of `sample`")
        t = int(timestep)
        op = self.op
        op.set_time_table([t], time_sinusoid(torch.tensor([t])))
        out = op.forward(sample.reshape(R, 256), t, encoder_hidden_states.contiguous())
        return (out.view(R, 1, 256),)


def time_sinusoid(timesteps: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """``get_timestep_embedding(t, 256, flip_sin_to_cos=True, downscale_freq_shift=0)``
    (mld/models/architectures/tools/embeddings.py:245-285) on the host, with the same tensor ops."""
    import math
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32)
    exponent = exponent / (half - 0)
    emb = torch.exp(exponent)
    emb = timesteps.to("cpu")[:, None].float() * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    return torch.cat([emb[:, half:], emb[:, :half]], dim=-1)


class MldVae(_PackedModule):
    """``mld.models.architectures.mld_vae.MldVae`` (encoder_decoder arch, learned PE; the constructor
    hard-codes 5 layers / 1 head / ff 128 like the reference, mld_vae.py:51-53)."""

    def __init__(self, ablation, nfeats: int, latent_dim: list = [1, 256], ff_size: int = 1024, num_layers: int = 9,
                 num_heads: int = 4, dropout: float = 0.1, arch: str = "all_encoder", normalize_before: bool = False,
                 activation: str = "gelu", position_embedding: str = "learned", max_batch: int = 512, max_frames: int = 60,
                 **kwargs) -> None:
        super().__init__()
        self.latent_size = latent_dim[0]
        self.latent_dim = latent_dim[-1]
        self.nfeats = nfeats
        self.arch = arch
        self.mlp_dist = _attr(ablation, "MLP_DIST", False)
        self.pe_type = _attr(ablation, "PE_TYPE", "mld")
        self.max_batch, self.max_frames = max_batch, max_frames
        unsupported = []
        if arch != "encoder_decoder":
            unsupported.append("only arch=encoder_decoder is built")
        if self.pe_type != "mld" or position_embedding not in ("learned", "v3"):
            unsupported.append("only PE_TYPE=mld with learned position embedding is built")
        if self.mlp_dist:
            unsupported.append("MLP_DIST=True is not built")
        if list(latent_dim) != [1, 256] or normalize_before or activation != "gelu":
            unsupported.append("kernels are specialised for latent_dim [1,256], post-norm, gelu")
        if unsupported:
            raise NotImplementedError("MldVae (sm_100a): " + "; ".join(unsupported))
        _register_spec(self, synthetic.vae_spec(nfeats))

    @property
    def op(self) -> ops.VaeOp:
        self._require_cuda()
        return self._op("vae", lambda: ops.VaeOp(self.state_dict(), self.nfeats, self.max_batch, self.max_frames))

    def forward(self, features, lengths=None):
        print("Should Not enter here")                                             # mld_vae.py:118-126
        z, dist = self.encode(features, None, lengths)
        return self.decode(z, lengths), z, dist

    def encode(self, features, images=None, lengths: Optional[List[int]] = None, eps: Optional[torch.Tensor] = None,
               lengths_dev: Optional[torch.Tensor] = None):
        """[B,T,nfeats] -> (latent [1,B,256], Normal(mu,std)).  ``eps`` ([1,B,256]) is the N(0,1) draw of
        ``rsample``; when None it is drawn with ``torch.randn`` on the features' device.  ``lengths_dev``: the same
        lengths already on the device (avoids a blocking host-to-device copy in the middle of the stream)."""
        if lengths is None:
            lengths = [len(f) for f in features]
        B = features.shape[0]
        if eps is None:
            eps = torch.randn(1, B, self.latent_dim, device=features.device, dtype=torch.float32)
        z, mu, std = self.op.encode(features, lengths_dev if lengths_dev is not None else torch.as_tensor(lengths), eps)
        dist = torch.distributions.Normal(mu.unsqueeze(0), std.unsqueeze(0), validate_args=False)
        return z.unsqueeze(0), dist

    def decode(self, z, lengths: List[int], T: Optional[int] = None, lengths_dev: Optional[torch.Tensor] = None):
        """z [1,B,256] -> [B,max(lengths),nfeats]  (``T`` overrides max(lengths): a sub-batch decoded to the frame count of
        the whole batch; ``lengths_dev``: the lengths already on the device)"""
        T = int(max(lengths)) if T is None else int(T)
        if lengths_dev is not None:
            return self.op.decode(z.reshape(-1, self.latent_dim), lengths_dev, T)
        return self.op.decode(z.reshape(-1, self.latent_dim), torch.as_tensor(lengths), T)


class ResnetPointnet(_PackedModule):
    """``EgoHMR.models.respointnet.ResnetPointnet`` (out_dim 512, hidden 256).  ``output_scene`` is fused
    behind the same handle; ``forward`` returns the 512-d code like the reference."""

    def __init__(self, out_dim: int = 512, hidden_dim: int = 256, max_batch: int = 512, max_points: int = 20000,
                 precision: int = -1):
        super().__init__()
        if out_dim != 512 or hidden_dim != 256:
            raise NotImplementedError("ResnetPointnet (sm_100a): out_dim 512 / hidden_dim 256 only")
        self.out_dim = out_dim
        self.precision = int(precision)      # operand format of the per-point GEMMs (seeme_pointnet_create_ex)
        self.max_batch, self.max_points = max_batch, max_points
        _register_spec(self, synthetic.pointnet_spec())

    def op(self, output_scene: Optional[nn.Module] = None) -> ops.PointNetOp:
        """One handle serves both ``encode_scene`` (512-d) and the fused ``output_scene`` tail (256-d).
        The handle is rebuilt when either this module's or ``output_scene``'s parameters change."""
        dev = self._require_cuda()
        if output_scene is None:
            output_scene = self.__dict__.get("_out_scene")
            if output_scene is None:   # standalone use: a zero tail, only the 512-d code is read
                output_scene = nn.Sequential(nn.ReLU(), nn.Linear(512, 256))
                for p in output_scene.parameters():
                    p.requires_grad_(False).zero_()
                self.__dict__["_out_scene"] = output_scene
            output_scene.to(dev)
        else:
            self.__dict__["_out_scene"] = output_scene
        lin = output_scene[1]
        sig_extra = (lin.weight.data_ptr(), lin.weight._version, lin.bias.data_ptr(), lin.bias._version)
        cache = self.__dict__.setdefault("_op_cache", {})
        key = ("pn", _LANE[0])
        ent = cache.get(key)
        sig = (self._signature(), sig_extra, self.precision)
        if ent is None or ent[0] != sig:
            if ent is not None:
                ent[1].close()
            cache[key] = (sig, ops.PointNetOp(self.state_dict(), {"1.weight": lin.weight, "1.bias": lin.bias},
                                              self.max_batch, self.max_points, precision=self.precision))
        return cache[key][1]

    def forward(self, p):
        _, feat = self.op()(p, want_feat=True)
        return feat


class ResNet50Backbone(_PackedModule):
    """``EgoHMR.models.resnet.resnet50`` (resnet.py:99-180, 211-224) at inference: same ``state_dict`` keys
    (``conv1.weight``, ``bn1.*``, ``layer{1..4}.{i}.conv{1,2,3}.weight`` / ``bn{1,2,3}.*`` / ``downsample.{0,1}.*``),
    eval-mode BatchNorm folded into the packed convolution weights by the kernel-side handle."""

    def __init__(self, max_batch: int = 512):
        super().__init__()
        self.max_batch = max_batch
        for conv, bn, cout, cin, k in synthetic.resnet50_convs():
            for key, shape, buf in ((conv + ".weight", (cout, cin, k, k), False), (bn + ".weight", (cout,), False),
                                    (bn + ".bias", (cout,), False), (bn + ".running_mean", (cout,), True),
                                    (bn + ".running_var", (cout,), True), (bn + ".num_batches_tracked", (), True)):
                parts = key.split(".")
                mod = self
                for p in parts[:-1]:
                    if p not in mod._modules:
                        mod.add_module(p, nn.Module())
                    mod = mod._modules[p]
                if buf:
                    dt = torch.long if parts[-1] == "num_batches_tracked" else torch.float32
                    init = torch.ones(shape) if parts[-1] == "running_var" else torch.zeros(shape, dtype=dt)
                    mod.register_buffer(parts[-1], init.to(dt))
                else:
                    mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape), requires_grad=False))

    def _signature(self):
        return tuple((p.data_ptr(), p._version, str(p.device)) for p in list(self.parameters()) + list(self.buffers()))

    def op(self, output_images: Optional[nn.Module] = None) -> ops.ResNet50Op:
        """One handle serves ``encode_image`` (2048-d) and the fused ``output_images`` tail (256-d); rebuilt when either
        module's tensors change."""
        dev = self._require_cuda()
        if output_images is None:
            output_images = self.__dict__.get("_out_images")
            if output_images is None:   # standalone use: a zero tail, only the 2048-d code is read
                output_images = nn.Sequential(nn.ReLU(), nn.Linear(2048, 256))
                for p in output_images.parameters():
                    p.requires_grad_(False).zero_()
                self.__dict__["_out_images"] = output_images
            output_images.to(dev)
        else:
            self.__dict__["_out_images"] = output_images
        lin = output_images[1]
        sig = (self._signature(), (lin.weight.data_ptr(), lin.weight._version, lin.bias.data_ptr(), lin.bias._version))
        cache = self.__dict__.setdefault("_op_cache", {})
        key = ("rn50", _LANE[0])
        ent = cache.get(key)
        if ent is None or ent[0] != sig:
            if ent is not None:
                ent[1].close()
            cache[key] = (sig, ops.ResNet50Op(self.state_dict(), {"1.weight": lin.weight, "1.bias": lin.bias}, self.max_batch))
        return cache[key][1]

    def forward(self, x):
        """[B,3,224,224] -> [B,2048]"""
        _, feat = self.op()(x, want_feat=True)
        return feat


class ProHMRScene(nn.Module):
    """The slice of ``EgoHMR.models.prohmr.prohmr_scene.ProHMRScene`` on the path: ``scene_enc`` / ``encode_scene``
    (:51,102-104) and, with ``with_backbone=True`` (the image-conditioned variants), ``backbone`` /
    ``encode_image`` (:33-34,99-100).  The flow / discriminator sub-trees are never run at SEE-ME test time and are
    not built (load reference checkpoints with ``strict=False``)."""

    def __init__(self, cfg=None, max_batch: int = 512, max_points: int = 20000, precision: int = -1,
                 with_backbone: bool = False, **kwargs):
        super().__init__()
        self.scene_enc = ResnetPointnet(512, 256, max_batch, max_points, precision=precision)
        if with_backbone:
            self.backbone = ResNet50Backbone(max_batch)

    def encode_scene(self, scene_pcd_verts):
        return self.scene_enc(scene_pcd_verts)

    def encode_image(self, x):
        if "backbone" not in self._modules:
            raise NotImplementedError("this ProHMRScene was built without the image backbone (with_backbone=True builds it)")
        return self.backbone(x)


SMPL_EXTRA_VERTEX_IDS = [332, 6260, 2800, 4071, 583,                     # nose, reye, leye, rear, lear
                         3216, 3226, 3387, 6617, 6624, 6787,             # L/R big toe, small toe, heel
                         2746, 2319, 2445, 2556, 2673,                   # left finger tips
                         6191, 5782, 5905, 6016, 6133]                   # right finger tips


class SMPL(_PackedModule):
    """``smplx.SMPL`` forward (smplx==0.1.28) for ``pose2rot=True`` axis-angle input.  Buffers use smplx's
    names so ``smpl_model.*`` checkpoint entries load.  ``model_path`` may be a dict of buffers (tests /
    synthetic) or a path to an SMPL ``.pkl``/``.npz`` (needs the licensed download)."""

    def __init__(self, model_path=None, batch_size: int = 1, gender: str = "neutral", max_frames: int = 65536, **kwargs):
        super().__init__()
        buf = model_path if isinstance(model_path, dict) else self._load(model_path)
        for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights"):
            self.register_buffer(k, buf[k].float().contiguous())
        self.register_buffer("parents", buf["parents"].long())
        self.register_buffer("faces_tensor", buf.get("faces_tensor", torch.zeros(13776, 3, dtype=torch.long)))
        self.max_frames = max_frames

    @staticmethod
    def _load(path):
        import os
        import pickle
        import numpy as np
        if path is None or not os.path.exists(path):
            raise FileNotFoundError(f"SMPL model file {path!r} not found (the licensed SMPL download is required)")
        if path.endswith(".npz"):
            d = dict(np.load(path, allow_pickle=True))
        else:
            with open(path, "rb") as f:
                d = pickle.load(f, encoding="latin1")
        V = np.asarray(d["v_template"]).shape[0]
        par = np.asarray(d["kintree_table"])[0].astype(np.int64)
        par[0] = -1
        J_reg = d["J_regressor"]
        J_reg = np.asarray(J_reg.todense()) if hasattr(J_reg, "todense") else np.asarray(J_reg)
        posedirs = np.asarray(d["posedirs"]).reshape(V * 3, -1).T           # smplx: reshape([-1, 207]).T
        t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))
        return {"v_template": t(d["v_template"]), "shapedirs": t(np.asarray(d["shapedirs"])[:, :, :10]),
                "posedirs": t(posedirs), "J_regressor": t(J_reg), "lbs_weights": t(d["weights"]),
                "parents": torch.from_numpy(par), "faces_tensor": torch.from_numpy(np.asarray(d["f"]).astype(np.int64))}

    def _signature(self):
        return tuple((b.data_ptr(), b._version, str(b.device)) for b in self.buffers())

    @property
    def op(self) -> ops.SmplOp:
        if not self.v_template.is_cuda:
            raise RuntimeError("SMPL: buffers are on the CPU; seeme_b200 runs on CUDA (sm_100a) only")
        return self._op("smpl", lambda: ops.SmplOp({k: getattr(self, k) for k in
                                                    ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights", "parents")},
                                                   self.max_frames))

    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None, pose2rot: bool = True, **kwargs):
        if not pose2rot:
            raise NotImplementedError("SMPL (sm_100a): only pose2rot=True (axis-angle) is on the path (DATA_TYPE: angle)")
        verts, j24, _ = self.op.forward(betas, body_pose, global_orient, transl, want_vertices=True, want_quat=False)
        extra = verts[:, SMPL_EXTRA_VERTEX_IDS]                      # VertexJointSelector (21 vertex-picked joints)
        return SimpleNamespace(vertices=verts, joints=torch.cat([j24, extra], dim=1), betas=betas, body_pose=body_pose,
                               global_orient=global_orient, transl=transl)
