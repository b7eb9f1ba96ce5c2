"""CPU tests of the boundary and host logic: the C-ABI library loads and exports every symbol that
include/seeme_b200.h declares (no compute without a GPU), the host-side scheduler tables equal the
oracle's, the YAML/config surface resolves, and the product path fails loudly without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "seeme_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(seeme_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from seeme_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build(verbose=False)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/seeme_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    l = _lib.lib()
    assert l.seeme_abi_version() == 1
    assert isinstance(_lib.launch_count(), int)


def test_key_orders_cover_state_dicts(weights):
    from seeme_b200 import ops
    assert sorted(ops.denoiser_keys() + ["mem_pos.pe"]) == sorted(weights["denoiser"].keys())
    assert sorted(ops.vae_keys()) == sorted(weights["vae"].keys())
    assert sorted(ops.pointnet_keys()) == sorted(weights["pointnet"].keys())
    assert len(ops.denoiser_keys()) == 201 and len(ops.vae_keys()) == 169 and len(ops.pointnet_keys()) + 2 == 26


def test_scheduler_tables_match_oracle():
    from oracle import restate as O
    from seeme_b200.scheduler import DDIMScheduler
    s = DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                      clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    s.set_timesteps(50)
    o = O.DDIMRef()
    o.set_timesteps(50)
    assert s.timesteps.tolist() == o.timesteps.tolist() == list(range(981, 0, -20))
    assert torch.equal(s.alphas_cumprod, o.alphas_cumprod)
    c = s.step_coefficients()
    assert c.shape == (50, 4) and c.dtype == torch.float32
    # a step with the coefficient table == the oracle's step (same fp32 op order, on the CPU)
    x, e = torch.randn(3, 256), torch.randn(3, 256)
    for i in (0, 17, 49):
        t = int(s.timesteps[i])
        ref = o.step(e, t, x).prev_sample
        mine = c[i, 2] * ((x - c[i, 0] * e) / c[i, 1]) + c[i, 3] * e
        assert torch.allclose(ref, mine, atol=1e-6, rtol=1e-6)
    with pytest.raises(NotImplementedError):
        s.step(e, 1, x, eta=0.5)


def test_time_sinusoid_matches_oracle():
    from oracle import restate as O
    from seeme_b200.modules import time_sinusoid
    t = torch.tensor([981, 481, 21, 1, 0])
    assert torch.equal(time_sinusoid(t), O.timestep_embedding(t))


def test_config_surface_and_model_builds_on_cpu():
    import seeme_b200
    from seeme_b200.config import load_config
    cfg = load_config(os.path.join(seeme_b200.CONFIG_DIR, "config_mld_egobody.yaml"))
    assert cfg.model.denoiser.params.latent_dim == [1, 256]          # ${model.latent_dim}
    assert cfg.model.denoiser.params.ablation.MD_TRANS is True       # ${TRAIN.ABLATION}
    assert cfg.model.scheduler.num_inference_timesteps == 50
    assert cfg.TRAIN.ABLATION.PREDICT_EPSILON is True                # inherited from base.yaml
    for name, cond in (("config_mld_egobody.yaml", ["text", "scene", "interactee"]), ("config_mld_gimo.yaml", ["text", "scene"])):
        m = seeme_b200.build_model(name, device="cpu", max_batch=2)
        assert m.condition == cond
        keys = set(m.state_dict())
        assert {"denoiser.encoder.middle_block.sa_block.self_attn.in_proj_weight", "vae.final_layer.weight",
                "proscene.scene_enc.block_3.shortcut.weight", "output_scene.1.bias", "smpl_model.lbs_weights"} <= keys
    m = seeme_b200.build_model("config_mld_interactee.yaml", device="cpu", max_batch=2)
    assert m.estimate == "interactee" and m.cfg.TEST.REPLICATION_TIMES == 10


def test_unsupported_configs_raise():
    import seeme_b200
    with pytest.raises(NotImplementedError):     # the reference's image branches cannot run under CFG (mld.py:1299)
        seeme_b200.build_model(device="cpu", condition=["text", "scene", "image"], guidance_scale=7.5)
    m = seeme_b200.build_model(device="cpu", condition=["text", "scene", "image"], guidance_scale=1.0, max_batch=2, n_points=64)
    keys = m.state_dict().keys()
    assert "proscene.backbone.layer4.2.bn3.running_var" in keys and "output_images.1.weight" in keys
    assert "proscene.backbone.layer1.0.downsample.1.num_batches_tracked" in keys
    from seeme_b200.modules import MldDenoiser
    abl = {"SKIP_CONNECT": True, "MD_TRANS": True, "DIFF_PE_TYPE": "mld", "VAE_TYPE": "actor"}
    with pytest.raises(TypeError):
        MldDenoiser(abl, condition=["scene"], latent_dim=[1, 256], num_layers=5, num_heads=1, text_encoded_dim=256)
    with pytest.raises(NotImplementedError):
        MldDenoiser(abl, condition=["text"], latent_dim=[1, 256], num_layers=9, num_heads=4, text_encoded_dim=256)


def test_pipeline_defaults_and_handle_signature():
    """host logic of the batch pipeline that needs no GPU: the default depth by batch capacity, the work-queue variable, the
    handle signature (detects in-place updates and replaced parameters without walking the module tree), and the loud
    failure of the async entry points on the CPU"""
    import torch
    import seeme_b200
    assert os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS") is not None      # set (if unset) by importing the package
    for mb, depth in ((2, 8), (64, 8), (128, 8), (256, 32), (512, 16), (1024, 8), (4096, 4)):
        m = seeme_b200.build_model(device="cpu", max_batch=mb, n_points=64)
        assert m.pipeline_depth == depth, (mb, m.pipeline_depth)
    assert seeme_b200.build_model(device="cpu", max_batch=256, n_points=64, pipeline_depth=5).pipeline_depth == 5
    assert m.sampler_backend == "auto" and m.encoder_handles == 4 and m.persistent_sm_budget == 64
    with pytest.raises(RuntimeError, match="CUDA"):
        m.prepare_pipeline()
    from seeme_b200 import synthetic as S
    with pytest.raises(RuntimeError, match="CUDA"):
        m.ego_eval_async(S.make_batch(2, n_points=64))
    den = m.denoiser
    s0 = den._signature()
    assert s0 == den._signature() and len(s0) == len(list(den.parameters()))
    p = next(den.parameters())
    with torch.no_grad():
        p.add_(1.0)                                                        # in-place update: version counter
    s1 = den._signature()
    assert s1 != s0
    name, mod = None, None
    for mod_ in den.modules():
        for n_ in mod_._parameters:
            name, mod = n_, mod_
    setattr(mod, name, torch.nn.Parameter(mod._parameters[name].detach().clone(), requires_grad=False))   # replaced object
    assert den._signature() != s1


def test_no_cpu_fallback():
    """the product path must fail loudly when asked to run without CUDA"""
    import seeme_b200
    from seeme_b200 import synthetic as S
    m = seeme_b200.build_model(device="cpu", max_batch=2, n_points=64)
    batch = S.make_batch(2, n_points=64)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.ego_eval(batch)
    src = ""
    for root, _, files in os.walk(os.path.join(ROOT, "seeme_b200")):
        for f in files:
            if f.endswith(".py"):
                src += open(os.path.join(root, f)).read()
    assert "import oracle" not in src and "from oracle" not in src


def test_metric_matches_reference_computemetrics():
    """batched EgoMetric vs the reference's per-frame numpy ComputeMetrics.update (build container only)"""
    from oracle import ref_modules as R
    if not R.available():
        pytest.skip("/root/reference not present")
    R.import_mld()
    from mld.models.metrics.compute import ComputeMetrics
    from seeme_b200.metrics import EgoMetric
    g = torch.Generator().manual_seed(11)
    B, T_ = 5, 60
    jr = torch.randn(B, T_, 24, 3, generator=g) * 0.3
    jp = jr + 0.02 * torch.randn(B, T_, 24, 3, generator=g)
    jp[3] += 1.0 * torch.randn(T_, 1, 3, generator=g)               # a sequence the root-error gate rejects
    qr = torch.nn.functional.normalize(torch.randn(B * T_, 4, generator=g), dim=1)
    qp = torch.nn.functional.normalize(qr + 0.05 * torch.randn(B * T_, 4, generator=g), dim=1)
    qp[T_:2 * T_] = torch.nn.functional.normalize(torch.randn(T_, 4, generator=g), dim=1)   # head-orientation gate
    ji = torch.randn(B, T_, 24, 3, generator=g)
    qi = torch.nn.functional.normalize(torch.randn(B * T_, 4, generator=g), dim=1)
    lengths = [60, 45, 60, 60, 20]
    ref = ComputeMetrics(njoints=23, jointstype="humanml3d", dist_sync_on_step=False)
    mine = EgoMetric(njoints=23)
    for split in ("test", "val"):
        ref.reset(); mine.reset()
        ref.update(split, jp, jr, qp, qr, ji[:, :, [0]], ji, qi, None, lengths, {})
        mine.update(split, jp, jr, qp, qr, ji[:, :, [0]], ji, qi, None, lengths, {})
        for k in ("MPJPE", "ROOT_ERROR", "ACCL", "HEAD_ORIENTATION_ERROR"):
            assert float(getattr(ref, k)) == pytest.approx(mine.state[k], rel=1e-5, abs=1e-6), (split, k)
        for k in ("count", "count_seq", "count_seq_root", "count_seq_accl", "count_seq_head_orientation"):
            assert float(getattr(ref, k)) == mine.state[k], (split, k)


def test_mr_metric_and_pve_match_reference():
    """batched MRMetric (MPJPE / PA-MPJPE / ACCEL) and vertice_pve vs the reference's per-sequence CPU code (build container only)"""
    from oracle import ref_modules as R
    if not R.available():
        pytest.skip("/root/reference not present")
    R.import_mld()
    from mld.models.metrics.mr import MRMetrics
    from mld.models.metrics import metrics_utils_egobody as U
    from seeme_b200.metrics import MRMetric, vertice_pve
    g = torch.Generator().manual_seed(5)
    B, T_, J = 4, 60, 22
    ref_j = torch.randn(B, T_, J, 3, generator=g) * 0.4
    rst_j = 1.1 * ref_j @ torch.linalg.qr(torch.randn(3, 3, generator=g))[0] + 0.03 * torch.randn(B, T_, J, 3, generator=g) + 0.2
    lengths = [60, 33, 60, 48]
    a, b = MRMetrics(njoints=J, jointstype="humanml3d", dist_sync_on_step=False), MRMetric(njoints=J, jointstype="humanml3d")
    for _ in range(2):
        a.update(rst_j, ref_j, lengths)
        b.update(rst_j, ref_j, lengths)
    ra, rb = a.compute(sanity_flag=False), b.compute()
    for k in ("MPJPE", "PAMPJPE", "ACCEL"):
        assert float(ra[k]) == pytest.approx(rb[k], rel=2e-5), k
    assert float(a.count) == b.state["count"] and float(a.count_seq) == b.state["count_seq"]
    with pytest.raises(NotImplementedError):
        MRMetric(njoints=J, jointstype="smpl")
    pv = torch.randn(6, 500, 3, generator=g)
    tv = 0.9 * pv @ torch.linalg.qr(torch.randn(3, 3, generator=g))[0] + 0.01 * torch.randn(6, 500, 3, generator=g)
    for al in ("none", "scale", "procrustes"):
        assert float(vertice_pve(pv, tv, al)) == pytest.approx(float(U.vertice_pve(pv.numpy(), tv.numpy(), alignment=al)), rel=1e-4), al
