"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference modules from /root/reference.

Only usable inside the build container (``/root/reference`` does not exist on the GPU box).
It is used (a) by ``oracle/make_golden.py`` to generate the committed fixtures under
``tests/golden/`` and (b) by the ``-m "not gpu"`` tests that pin ``oracle/restate.py`` against
the reference's own code.  Nothing under ``seeme_b200/`` may import this file.

Missing third-party packages the reference imports at module import time are replaced by
empty stub modules (they are never *called* on the hot path):
  * ``clip``  (``mld/models/architectures/mdiff_transformer.py:10``)
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

REF_ROOT = os.environ.get("SEEME_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "mld"))


def _stub(name: str) -> None:
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)


def _prepare() -> None:
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _stub("clip")


def ablation():
    """TRAIN.ABLATION at the north-star configs (configs/config_mld_egobody.yaml:45-50 over
    configs/base.yaml:18-27)."""
    return SimpleNamespace(SKIP_CONNECT=True, PE_TYPE="mld", DIFF_PE_TYPE="mld", MD_TRANS=True,
                           VAE_TYPE="actor", MLP_DIST=False, PREDICT_EPSILON=True, PREDICT_TRANSL=True)


def build_denoiser(condition=("text", "scene", "interactee")):
    """MldDenoiser with configs/modules/denoiser.yaml params."""
    _prepare()
    from mld.models.architectures.mld_denoiser import MldDenoiser
    m = MldDenoiser(ablation=ablation(), nfeats=75, condition=list(condition), latent_dim=[1, 256],
                    ff_size=128, num_layers=5, num_heads=1, dropout=0.1, normalize_before=False,
                    activation="gelu", flip_sin_to_cos=True, return_intermediate_dec=False,
                    position_embedding="learned", arch="trans_enc", freq_shift=0,
                    guidance_scale=7.5, guidance_uncondp=0.1, text_encoded_dim=256, nclasses=10)
    return m.eval()


def build_vae(nfeats=75):
    """MldVae with configs/modules/motion_vae.yaml params (depth/heads/ff overridden in the ctor)."""
    _prepare()
    from mld.models.architectures.mld_vae import MldVae
    m = MldVae(ablation=ablation(), nfeats=nfeats, latent_dim=[1, 256], ff_size=1024, num_layers=9,
               num_heads=4, dropout=0.1, arch="encoder_decoder", normalize_before=False,
               activation="gelu", position_embedding="learned")
    return m.eval()


def build_pointnet():
    """ResnetPointnet(out_dim=512, hidden_dim=256) as built at EgoHMR/models/prohmr/prohmr_scene.py:51."""
    _prepare()
    from EgoHMR.models.respointnet import ResnetPointnet
    return ResnetPointnet(out_dim=512, hidden_dim=256).eval()


def aa_to_quat(theta):
    _prepare()
    from mld.utils.geometry2 import aa_to_quat as f
    return f(theta)


def lengths_to_mask(lengths, device):
    _prepare()
    from mld.utils.temos_utils import lengths_to_mask as f
    return f(lengths, device)


# --------------------------------------------------------------------------------------------
# importing mld/models/modeltype/mld.py itself (for the unmodified MLD._diffusion_reverse / ego_eval)
# --------------------------------------------------------------------------------------------
class _AnyStub(types.ModuleType):
    """A module whose every attribute is a permissive dummy class -- enough for ``import x`` /
    ``from x import y`` / ``class Foo(x.Bar)`` at import time of code we never call."""
    __path__: list = []
    __all__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = f"{self.__name__}.{name}"
        if sub in sys.modules:
            return sys.modules[sub]
        import torch.nn as nn
        if name in ("LightningModule",):
            return nn.Module
        if name == "Metric":
            return _metric_shim()
        cls = type(name, (), {"__init__": lambda self, *a, **k: None,
                              "__call__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


def _metric_shim():
    import torch.nn as nn

    class Metric(nn.Module):
        """10-line torchmetrics.Metric shim: add_state/reset."""

        def __init__(self, *a, **k):
            super().__init__()
            self._defaults = {}

        def add_state(self, name, default, dist_reduce_fx=None):
            self._defaults[name] = default
            setattr(self, name, default.clone() if hasattr(default, "clone") else list(default))

        def reset(self):
            for k, v in self._defaults.items():
                setattr(self, k, v.clone() if hasattr(v, "clone") else list(v))
    return Metric


class _StubFinder:
    ROOTS = ("smplx", "torchmetrics", "yacs", "matplotlib", "omegaconf", "pytorch_lightning", "UMNN",
             "clip", "trimesh", "pytorch3d", "kornia", "diffusers", "spacy", "wandb", "pyrender",
             "open3d", "chumpy", "loguru", "rich", "tensorboard", "natsort", "shortuuid", "bert_score",
             "moviepy", "imageio", "smplpytorch", "human_body_prior", "PIL", "skimage", "tensorboardX")

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        import importlib.util
        root = fullname.split(".")[0]
        if root not in self.ROOTS:
            return None
        try:  # a real installed package wins
            for f in sys.meta_path:
                if f is self:
                    continue
                spec = f.find_spec(fullname, path, target) if hasattr(f, "find_spec") else None
                if spec is not None:
                    return spec if root not in _FORCED else None
        except Exception:
            pass
        return importlib.machinery.ModuleSpec(fullname, self, is_package=True)

    def create_module(self, spec):
        return _AnyStub(spec.name)

    def exec_module(self, module):
        pass


_FORCED: set = set()


def import_mld():
    """Returns the reference's ``mld.models.modeltype.mld`` module (unmodified source), imported with
    stub packages for everything that is not installed offline."""
    _prepare()
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    nf = os.path.join(REF_ROOT, "nflows")
    if nf not in sys.path:
        sys.path.insert(1, nf)
    import importlib
    return importlib.import_module("mld.models.modeltype.mld")


class _SMPLOut:
    def __init__(self, vertices, joints):
        self.vertices, self.joints = vertices, joints


def make_carrier(weights, smpl_buffers, stats, condition=("text", "scene", "interactee"), guidance_scale=7.5,
                 dataset="egobody", n_steps=50, estimate="wearer", pred_global_orient=True):
    """An ``nn.Module`` carrying exactly the attributes the unmodified ``MLD.ego_eval`` /
    ``MLD._diffusion_reverse`` read, built WITHOUT ``MLD.__init__`` (which ``torch.load``s an absent
    EgoHMR checkpoint, mld.py:193-196, and constructs smplx).  Real reference networks; DDIM and SMPL
    come from the restatements in ``oracle/restate.py`` (neither package exists offline).
    ``save_for_edo`` is False."""
    import torch
    import torch.nn as nn
    from oracle import restate as O
    mld_mod = import_mld()

    class SMPLRef(nn.Module):
        def forward(self, betas=None, body_pose=None, global_orient=None, transl=None, pose2rot=True, **kw):
            v, j = O.smpl_forward(smpl_buffers, betas, body_pose, global_orient, transl)
            j45 = torch.cat([j, torch.zeros(j.shape[0], 21, 3)], dim=1)
            return _SMPLOut(v, j45)

    class Scene(nn.Module):
        def __init__(self):
            super().__init__()
            self.scene_enc = build_pointnet()

        def encode_scene(self, p):
            return self.scene_enc(p)

    class Carrier(nn.Module):
        # the reference's own, unmodified functions bound as methods
        _diffusion_reverse = mld_mod.MLD._diffusion_reverse
        ego_eval = mld_mod.MLD.ego_eval

    c = Carrier()
    c.denoiser = build_denoiser(condition)
    c.denoiser.load_state_dict(weights["denoiser"])
    c.vae = build_vae()
    c.vae.load_state_dict(weights["vae"])
    c.proscene = Scene()
    c.proscene.scene_enc.load_state_dict(weights["pointnet"])
    c.output_scene = nn.Sequential(nn.ReLU(), nn.Linear(512, 256))
    c.output_scene.load_state_dict(weights["output_scene"])
    c.smpl_model = SMPLRef()
    c.scheduler = O.DDIMRef()
    mean, std = stats
    c.renorm = lambda x: O.renorm(x, mean, std)
    sched_cfg = SimpleNamespace(num_inference_timesteps=n_steps, eta=0.0)
    c.cfg = SimpleNamespace(model=SimpleNamespace(scheduler=sched_cfg), DATASET=SimpleNamespace(NFEATS=75))
    c.condition = list(condition)
    c.guidance_scale = guidance_scale
    c.do_classifier_free_guidance = guidance_scale > 1.0
    c.vae_type, c.stage, c.latent_dim = "mld", "diffusion", [1, 256]
    c.estimate, c.predict_transl, c.data_type, c.name_dataset = estimate, True, "angle", dataset   # cfg.ESTIMATE (mld.py:111)
    c.save_for_edo = False
    c.save_cnt = 0
    c.pred_betas = c.pose_estimation_task = c.see_future = False
    c.global_orient_egoego = c.transl_egoego = c.pred_transl_egohmr = False
    c.pred_global_orient = pred_global_orient   # TEST.GLOBAL_ORIENT_PRED (config_mld_egobody.yaml:73: True; mld.py:112,1501-1505)
    c.times = []
    c.eval()
    c._ref_ego_eval = c.ego_eval
    c._ref_diffusion_reverse = c._diffusion_reverse
    return c


class noise_queue:
    """Context manager feeding pre-generated noise to the reference in its own draw order:
    ``Normal.rsample`` -> ``torch.distributions.utils._standard_normal`` (cond, then uncond under CFG),
    then ``torch.randn`` in ``_diffusion_reverse`` (mld.py:449-453)."""

    def __init__(self, normal_draws, randn_draws):
        self.normal, self.randn = list(normal_draws), list(randn_draws)

    def __enter__(self):
        import unittest.mock as M
        import torch
        import torch.distributions.normal as N
        self._p = [M.patch.object(N, "_standard_normal", lambda shape, dtype, device: self.normal.pop(0).reshape(shape).clone()),
                   M.patch.object(torch, "randn", lambda *a, **k: self.randn.pop(0).clone())]
        for p in self._p:
            p.start()
        return self

    def __exit__(self, *exc):
        for p in self._p:
            p.stop()
