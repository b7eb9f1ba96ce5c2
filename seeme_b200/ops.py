"""Operator handles over the C ABI, fed with ``torch`` CUDA tensors (PyTorch is only the owner of
device memory and streams here).  Key orders below are the tensor orders documented in
``include/seeme_b200.h``; keys are the reference's ``state_dict`` names."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib

BLOCKS = ["input_blocks.0", "input_blocks.1", "middle_block", "output_blocks.0", "output_blocks.1"]


def _wb(name: str) -> List[str]:
    return [name + ".weight", name + ".bias"]


def pointnet_keys() -> List[str]:
    """keys of ``proscene.scene_enc.*`` followed by ``output_scene.1.*`` (prefix 'output_scene.')"""
    k = _wb("fc_pos_0")
    for i in range(4):
        k += _wb(f"block_{i}.fc_0") + _wb(f"block_{i}.fc_1") + [f"block_{i}.shortcut.weight"]
    k += _wb("fc_c")
    return k


def _mha(name: str) -> List[str]:
    return [name + ".in_proj_weight", name + ".in_proj_bias"] + _wb(name + ".out_proj")


def vae_keys() -> List[str]:
    k = ["global_motion_token", "query_pos_encoder.pe", "query_pos_decoder.pe"] + _wb("skel_embedding") + _wb("final_layer")
    for stack, cross in (("encoder", False), ("decoder", True)):
        k += _wb(f"{stack}.norm") + _wb(f"{stack}.linear_blocks.0") + _wb(f"{stack}.linear_blocks.1")
        for b in BLOCKS:
            p = f"{stack}.{b}."
            k += _mha(p + "self_attn")
            if cross:
                k += _mha(p + "multihead_attn")
            k += _wb(p + "linear1") + _wb(p + "linear2") + _wb(p + "norm1") + _wb(p + "norm2")
            if cross:
                k += _wb(p + "norm3")
    return k


def denoiser_keys() -> List[str]:
    k = _wb("time_embedding.linear_1") + _wb("time_embedding.linear_2") + ["query_pos.pe"] + _wb("encoder.norm")
    k += _wb("encoder.linear_blocks.0") + _wb("encoder.linear_blocks.1")
    for b in BLOCKS:
        p = f"encoder.{b}."
        k += _mha(p + "sa_block.self_attn") + _wb(p + "sa_block.linear1") + _wb(p + "sa_block.linear2")
        k += _wb(p + "sa_block.norm1") + _wb(p + "sa_block.norm2")
        k += _wb(p + "ca_block.norm") + _wb(p + "ca_block.text_norm") + _wb(p + "ca_block.query") + _wb(p + "ca_block.key")
        k += _wb(p + "ca_block.value") + _wb(p + "ca_block.proj_out.emb_layers.1") + _wb(p + "ca_block.proj_out.norm")
        k += _wb(p + "ca_block.proj_out.out_layers.2")
        k += _wb(p + "ffn.linear1") + _wb(p + "ffn.linear2") + _wb(p + "ffn.proj_out.emb_layers.1")
        k += _wb(p + "ffn.proj_out.norm") + _wb(p + "ffn.proj_out.out_layers.2")
    return k


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (seeme_b200 has no CPU path)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


def _ptr_array(tensors: Sequence[torch.Tensor]):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class _Handle:
    _destroy = None

    def __init__(self):
        self.h = C.c_void_p()

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            getattr(_lib.lib(), self._destroy)(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PointNetOp(_Handle):
    """``output_scene(proscene.encode_scene(pcd))`` -- prohmr_scene.py:102-104, mld.py:257-261,1153-1154"""
    _destroy = "seeme_pointnet_destroy"

    def __init__(self, scene_enc_sd: Dict[str, torch.Tensor], output_scene_sd: Dict[str, torch.Tensor], max_batch: int,
                 max_points: int = 20000, precision: int = -1):
        """precision: GEMM operand format of the per-point contractions (include/seeme_b200.h,
        ``seeme_pointnet_create_ex``); -1 = library default (split-bf16 unless SEEME_POINTNET_PRECISION is set)."""
        super().__init__()
        ts = [_dev_f32(scene_enc_sd[k], k) for k in pointnet_keys()]
        ts += [_dev_f32(output_scene_sd["1.weight"], "output_scene.1.weight"), _dev_f32(output_scene_sd["1.bias"], "output_scene.1.bias")]
        self.device = ts[0].device
        with torch.cuda.device(self.device):
            if precision < 0:
                _lib.check(_lib.lib().seeme_pointnet_create(C.byref(self.h), _ptr_array(ts), len(ts), max_batch, max_points),
                           "seeme_pointnet_create")
            else:
                _lib.check(_lib.lib().seeme_pointnet_create_ex(C.byref(self.h), _ptr_array(ts), len(ts), max_batch, max_points,
                                                               int(precision)), "seeme_pointnet_create_ex")
        self.precision = precision
        self.max_batch, self.max_points = max_batch, max_points

    def __call__(self, pcd: torch.Tensor, want_feat: bool = False):
        pcd = _dev_f32(pcd, "pcd")
        B, N, _ = pcd.shape
        emb = torch.empty(B, 256, device=pcd.device, dtype=torch.float32)
        feat = torch.empty(B, 512, device=pcd.device, dtype=torch.float32) if want_feat else None
        with torch.cuda.device(pcd.device):
            _lib.check(_lib.lib().seeme_pointnet_forward(self.h, pcd.data_ptr(), B, N, feat.data_ptr() if want_feat else None,
                                                         emb.data_ptr(), _stream()), "seeme_pointnet_forward")
        return (emb, feat) if want_feat else emb


def resnet50_keys() -> List[str]:
    """``backbone.*`` keys in the tensor order of ``seeme_resnet50_create`` (include/seeme_b200.h)"""
    from . import synthetic
    k: List[str] = []
    for conv, bn, _, _, _ in synthetic.resnet50_convs():
        k += [conv + ".weight"] + [f"{bn}.{f}" for f in ("weight", "bias", "running_mean", "running_var")]
    return k


class ResNet50Op(_Handle):
    """``output_images(proscene.encode_image(images))`` -- prohmr_scene.py:99-100, EgoHMR/models/resnet.py:168-180,
    mld.py:251-255,1083-1086"""
    _destroy = "seeme_resnet50_destroy"

    def __init__(self, backbone_sd: Dict[str, torch.Tensor], output_images_sd: Dict[str, torch.Tensor], max_batch: int):
        super().__init__()
        ts = [_dev_f32(backbone_sd[k], k) for k in resnet50_keys()]
        ts += [_dev_f32(output_images_sd["1.weight"], "output_images.1.weight"), _dev_f32(output_images_sd["1.bias"], "output_images.1.bias")]
        self.device = ts[0].device
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().seeme_resnet50_create(C.byref(self.h), _ptr_array(ts), len(ts), max_batch), "seeme_resnet50_create")
        self.max_batch = max_batch

    def __call__(self, images: torch.Tensor, want_feat: bool = False):
        images = _dev_f32(images, "images")
        if images.dim() != 4 or tuple(images.shape[1:]) != (3, 224, 224):
            raise ValueError(f"images must be [B,3,224,224], got {tuple(images.shape)}")
        B = images.shape[0]
        emb = torch.empty(B, 256, device=images.device, dtype=torch.float32)
        feat = torch.empty(B, 2048, device=images.device, dtype=torch.float32) if want_feat else None
        with torch.cuda.device(images.device):
            _lib.check(_lib.lib().seeme_resnet50_forward(self.h, images.data_ptr(), B, feat.data_ptr() if want_feat else None,
                                                         emb.data_ptr(), _stream()), "seeme_resnet50_forward")
        return (emb, feat) if want_feat else emb


class VaeOp(_Handle):
    """``MldVae.encode`` / ``MldVae.decode`` -- mld_vae.py:128-256"""
    _destroy = "seeme_vae_destroy"

    def __init__(self, sd: Dict[str, torch.Tensor], nfeats: int, max_batch: int, max_frames: int = 60):
        super().__init__()
        ts = [_dev_f32(sd[k], k) for k in vae_keys()]
        self.device = ts[0].device
        self.nfeats = nfeats
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().seeme_vae_create(C.byref(self.h), _ptr_array(ts), len(ts), nfeats, max_batch, max_frames),
                       "seeme_vae_create")

    def encode(self, features: torch.Tensor, lengths: torch.Tensor, eps: torch.Tensor):
        features = _dev_f32(features, "features")
        eps = _dev_f32(eps.reshape(-1, 256), "eps")
        B, T, _ = features.shape
        lengths = lengths.to(device=features.device, dtype=torch.int32).contiguous()
        z = torch.empty(B, 256, device=features.device)
        mu, std = torch.empty_like(z), torch.empty_like(z)
        with torch.cuda.device(features.device):
            _lib.check(_lib.lib().seeme_vae_encode(self.h, features.data_ptr(), lengths.data_ptr(), eps.data_ptr(), B, T,
                                                   z.data_ptr(), mu.data_ptr(), std.data_ptr(), _stream()), "seeme_vae_encode")
        return z, mu, std

    def decode(self, z: torch.Tensor, lengths: torch.Tensor, T: int):
        z = _dev_f32(z.reshape(-1, 256), "z")
        B = z.shape[0]
        lengths = lengths.to(device=z.device, dtype=torch.int32).contiguous()
        out = torch.empty(B, T, self.nfeats, device=z.device)
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().seeme_vae_decode(self.h, z.data_ptr(), lengths.data_ptr(), B, T, out.data_ptr(), _stream()),
                       "seeme_vae_decode")
        return out


class DenoiserOp(_Handle):
    """``MldDenoiser.forward`` (mld_denoiser.py:151-244) and the fused ``_diffusion_reverse`` loop (mld.py:432-511)"""
    _destroy = "seeme_denoiser_destroy"

    def __init__(self, sd: Dict[str, torch.Tensor], max_rows: int):
        super().__init__()
        ts = [_dev_f32(sd[k], k) for k in denoiser_keys()]
        self.device = ts[0].device
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().seeme_denoiser_create(C.byref(self.h), _ptr_array(ts), len(ts), max_rows), "seeme_denoiser_create")
        self.max_rows = max_rows
        self.table_key = None          # timesteps the C-side tables hold, when they were built from a host sinusoid

    def set_backend(self, backend: str):
        """"persistent": one launch of the cluster kernel per run (lowest latency); "graph": the CUDA graph of small
        kernels (least SM time; for many concurrent chains next to the scene encoder)"""
        code = {"persistent": 0, "graph": 1, "tile": 2}[backend]
        if getattr(self, "_backend", None) != code:
            _lib.check(_lib.lib().seeme_denoiser_set_backend(self.h, code), "seeme_denoiser_set_backend")
            self._backend = code

    def set_time_table(self, timesteps: Sequence[int], sinusoid: Optional[torch.Tensor] = None):
        n = len(timesteps)
        self.table_key = tuple(int(t) for t in timesteps) if sinusoid is not None else None
        ts = (C.c_int32 * n)(*[int(t) for t in timesteps])
        sp = None
        if sinusoid is not None:
            s = sinusoid.detach().to("cpu", torch.float32).contiguous()
            assert tuple(s.shape) == (n, 256)
            sp = C.cast(s.data_ptr(), C.POINTER(C.c_float))
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().seeme_denoiser_set_time_table(self.h, ts, n, sp, _stream()), "seeme_denoiser_set_time_table")

    def forward(self, sample: torch.Tensor, timestep: int, cond: torch.Tensor):
        """sample [R,256], cond [Nc,R,256] -> [R,256]"""
        sample, cond = _dev_f32(sample, "sample"), _dev_f32(cond, "cond")
        R = sample.shape[0]
        Nc = cond.shape[0]
        if self.table_key != (int(timestep),):
            self.table_key = None      # the library rebuilds its table for [timestep] with its own sinusoid
        out = torch.empty_like(sample)
        with torch.cuda.device(sample.device):
            _lib.check(_lib.lib().seeme_denoiser_forward(self.h, sample.data_ptr(), int(timestep), cond.data_ptr(), Nc, R,
                                                         out.data_ptr(), _stream()), "seeme_denoiser_forward")
        return out

    def sample(self, x_T: torch.Tensor, cond: torch.Tensor, guidance_scale: float, timesteps: Sequence[int],
               coef: torch.Tensor):
        """x_T [B,256], cond [Nc,R,256], coef CPU fp32 [n,4] -> z [B,256]"""
        x_T, cond = _dev_f32(x_T, "x_T"), _dev_f32(cond, "cond")
        B, Nc, n = x_T.shape[0], cond.shape[0], len(timesteps)
        ts = (C.c_int32 * n)(*[int(t) for t in timesteps])
        coef = coef.detach().to("cpu", torch.float32).contiguous()
        assert tuple(coef.shape) == (n, 4)
        z = torch.empty_like(x_T)
        with torch.cuda.device(x_T.device):
            _lib.check(_lib.lib().seeme_sampler_run(self.h, x_T.data_ptr(), cond.data_ptr(), Nc, B, float(guidance_scale), n, ts,
                                                    C.cast(coef.data_ptr(), C.POINTER(C.c_float)), z.data_ptr(), _stream()),
                       "seeme_sampler_run")
        return z


def ddim_step(eps: torch.Tensor, sample: torch.Tensor, c: Sequence[float]) -> torch.Tensor:
    eps, sample = _dev_f32(eps, "eps"), _dev_f32(sample, "sample")
    out = torch.empty_like(sample)
    with torch.cuda.device(sample.device):
        _lib.check(_lib.lib().seeme_ddim_step(eps.data_ptr(), sample.data_ptr(), out.data_ptr(), sample.numel(),
                                              float(c[0]), float(c[1]), float(c[2]), float(c[3]), _stream()), "seeme_ddim_step")
    return out


class SmplOp(_Handle):
    """``smplx.SMPL.forward`` (+ ``aa_to_quat``, + fused ``renorm``) -- SURVEY App. C"""
    _destroy = "seeme_smpl_destroy"

    def __init__(self, buffers: Dict[str, torch.Tensor], max_frames: int):
        super().__init__()
        b = {k: _dev_f32(buffers[k], k) for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights")}
        self.device = b["v_template"].device
        par = [int(p) for p in buffers["parents"].tolist()]
        par[0] = -1
        parents = (C.c_int32 * 24)(*par)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().seeme_smpl_create(C.byref(self.h), b["v_template"].data_ptr(), b["shapedirs"].data_ptr(),
                                                    b["posedirs"].data_ptr(), b["J_regressor"].data_ptr(),
                                                    b["lbs_weights"].data_ptr(), parents, max_frames), "seeme_smpl_create")
        self.max_frames = max_frames

    def forward(self, betas, body_pose, global_orient, transl=None, want_vertices=True, want_quat=True):
        betas, body_pose, global_orient = _dev_f32(betas, "betas"), _dev_f32(body_pose, "body_pose"), _dev_f32(global_orient, "global_orient")
        F = betas.shape[0]
        dev = betas.device
        if transl is not None:
            transl = _dev_f32(transl, "transl")
        verts = torch.empty(F, 6890, 3, device=dev) if want_vertices else None
        joints = torch.empty(F, 24, 3, device=dev)
        quat = torch.empty(F, 4, device=dev) if want_quat else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().seeme_smpl_forward(self.h, betas.data_ptr(), body_pose.data_ptr(), global_orient.data_ptr(),
                                                     transl.data_ptr() if transl is not None else None, F,
                                                     verts.data_ptr() if want_vertices else None, joints.data_ptr(),
                                                     quat.data_ptr() if want_quat else None, _stream()), "seeme_smpl_forward")
        return verts, joints, quat

    def forward_feats(self, feats, mean, std, n_body, betas, want_vertices=True, want_m=True):
        """feats [F,Dn] normalised; mean/std float64 [Dn] CUDA -> (m float64 [F,Dn], verts, joints, quat)"""
        feats, betas = _dev_f32(feats, "feats"), _dev_f32(betas, "betas")
        F, Dn = feats.shape
        dev = feats.device
        mean = mean.to(dev, torch.float64).contiguous()
        std = std.to(dev, torch.float64).contiguous()
        m = torch.empty(F, Dn, device=dev, dtype=torch.float64) if want_m else None
        verts = torch.empty(F, 6890, 3, device=dev) if want_vertices else None
        joints = torch.empty(F, 24, 3, device=dev)
        quat = torch.empty(F, 4, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().seeme_smpl_forward_feats(self.h, feats.data_ptr(), Dn, mean.data_ptr(), std.data_ptr(), n_body,
                                                           betas.data_ptr(), F, m.data_ptr() if want_m else None,
                                                           verts.data_ptr() if want_vertices else None, joints.data_ptr(),
                                                           quat.data_ptr(), _stream()), "seeme_smpl_forward_feats")
        return m, verts, joints, quat
