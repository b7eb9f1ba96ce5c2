"""CPU tests pinning the oracle (oracle/restate.py): against the committed golden vectors produced by
the UNMODIFIED reference modules (tests/golden, generator oracle/make_golden.py), against the
reference modules themselves when /root/reference is present (build container), and against the
analytic known-answer tests listed in SURVEY.md section 4."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import restate as O
from oracle import ref_modules as R
from seeme_b200 import synthetic as S

from conftest import GOLDEN

T = lambda a: torch.from_numpy(np.asarray(a))
needs_ref = pytest.mark.skipif(not R.available(), reason="/root/reference not present (GPU box)")


def test_weight_generator_matches_fixture_checksums(weights, golden_stages):
    for k, sd in weights.items():
        got = float(sum(v.double().abs().sum() for v in sd.values()))
        assert got == pytest.approx(float(golden_stages["checksum_" + k]), rel=1e-12), k


@pytest.mark.parametrize("key,t,enc", [("den_out_t481_nc2", 481, "den_enc2"), ("den_out_t1_nc2", 1, "den_enc2"),
                                       ("den_out_t981_nc1", 981, "den_enc1")])
def test_denoiser_restatement_vs_reference_golden(weights, golden_stages, key, t, enc):
    g = golden_stages
    with torch.no_grad():
        out = O.denoiser_forward(weights["denoiser"], T(g["den_x"]), torch.tensor(t), T(g[enc]))
    assert torch.allclose(out, T(g[key]), atol=2e-5, rtol=1e-5)


def test_vae_restatement_vs_reference_golden(weights, golden_stages):
    g = golden_stages
    lens = g["vae_lens"].tolist()
    with torch.no_grad():
        z, mu, std = O.vae_encode(weights["vae"], T(g["vae_f"]), lens, T(g["vae_eps"]))
        dec = O.vae_decode(weights["vae"], T(g["vae_z"]), lens)
    assert torch.allclose(z, T(g["vae_z"]), atol=3e-5, rtol=1e-5)
    assert torch.allclose(mu, T(g["vae_mu"]), atol=3e-5, rtol=1e-5)
    assert torch.allclose(std, T(g["vae_std"]), atol=3e-5, rtol=1e-5)
    assert torch.allclose(dec, T(g["vae_dec"]), atol=3e-5, rtol=1e-5)


def test_pointnet_restatement_vs_reference_golden(weights, golden_stages):
    g = golden_stages
    with torch.no_grad():
        out = O.pointnet_forward(weights["pointnet"], T(g["pn_p"]))
        zero = O.pointnet_forward(weights["pointnet"], torch.zeros(1, 16, 3))
        zero2 = O.pointnet_forward(weights["pointnet"], torch.zeros(1, 5, 3))
    assert torch.allclose(out, T(g["pn_out"]), atol=1e-5)
    assert torch.allclose(zero, T(g["pn_out_zero"]), atol=1e-6)
    assert torch.allclose(zero, zero2, atol=1e-6)     # App. H8: the all-zero cloud embedding is independent of N


def test_aa_to_quat_vs_reference_golden(golden_stages):
    g = golden_stages
    assert torch.allclose(O.aa_to_quat(T(g["quat_theta"])), T(g["quat_out"]), atol=1e-7)


@pytest.mark.parametrize("name,dataset,cond,gs", [("egobody_cfg", "egobody", ("text", "scene", "interactee"), 7.5),
                                                  ("egobody_nocfg", "egobody", ("text", "scene", "interactee"), 1.0),
                                                  ("gimo_cfg", "gimo", ("text", "scene"), 7.5),
                                                  # config_mld_interactee.yaml protocol: ESTIMATE interactee, MOTION_LENGTH 1
                                                  ("interactee_T1_scene", "egobody", ("text", "scene"), 1.0),
                                                  ("interactee_T1_scene_int_cfg", "egobody", ("text", "scene", "interactee"), 7.5),
                                                  # TEST.GLOBAL_ORIENT_PRED: False (mld.py:1501-1505)
                                                  ("egobody_gt_orient", "egobody", ("text", "scene", "interactee"), 7.5)])
def test_ego_eval_restatement_vs_reference_golden(weights, smpl_buffers, name, dataset, cond, gs):
    g = dict(np.load(os.path.join(GOLDEN, f"ego_eval_{name}.npz")))
    B = g["joints_rst"].shape[0]
    Tm, est, pgo = int(g["cfg_T"]), str(g["cfg_estimate"]), bool(g["cfg_pred_global_orient"])
    batch = S.make_batch(B, n_points=1000, T=Tm, ragged=Tm > 1, dataset=dataset)
    noise = {k[6:]: T(v) for k, v in g.items() if k.startswith("noise_")}
    with torch.no_grad():
        rs = O.ego_eval(weights, smpl_buffers, S.norm_stats(), batch, noise, condition=cond, guidance_scale=gs, dataset=dataset,
                        estimate=est, pred_global_orient=pgo)
    assert rs["lengths"] == g["lengths"].tolist()
    assert rs["m_rst"].dtype == torch.float64 and g["m_rst"].dtype == np.float64     # App. D7
    for k in ("m_ref", "m_rst", "joints_ref", "joints_rst", "orientation_quat_rst", "orientation_quat_ref"):
        assert torch.allclose(rs[k].double(), T(g[k]).double(), atol=5e-5), k
    if "interactee" in cond:
        for k in ("joints_interactee", "root_interactee", "orientation_quat_int"):
            assert torch.allclose(rs[k], T(g[k]), atol=1e-5), k


# ---- live checks against the reference's own modules (build container only) -------------------
@needs_ref
def test_state_dict_specs_match_reference_modules(weights):
    for m, sd in ((R.build_denoiser(), weights["denoiser"]), (R.build_vae(), weights["vae"]), (R.build_pointnet(), weights["pointnet"])):
        ref = m.state_dict()
        assert set(ref) == set(sd)
        assert all(tuple(ref[k].shape) == tuple(sd[k].shape) for k in ref)


@needs_ref
def test_diffusion_reverse_restatement_vs_unmodified_reference(weights, smpl_buffers):
    c = R.make_carrier(weights, smpl_buffers, S.norm_stats(), guidance_scale=7.5, n_steps=10)
    g = torch.Generator().manual_seed(3)
    enc = torch.randn(6, 2, 256, generator=g)
    xT = torch.randn(3, 1, 256, generator=g)
    with torch.no_grad(), R.noise_queue([], [xT]):
        ref = c._ref_diffusion_reverse(enc, [60] * 3)
    with torch.no_grad():
        mine = O.diffusion_reverse(weights["denoiser"], enc, xT, 7.5, n_steps=10)
    assert torch.allclose(ref, mine, atol=1e-4, rtol=1e-5)


# ---- analytic known-answer tests ------------------------------------------------
def test_ddim_kats():
    s = O.DDIMRef()
    s.set_timesteps(50)
    ts = s.timesteps.tolist()
    assert ts[0] == 981 and ts[-1] == 1 and len(ts) == 50 and ts[1] == 961
    assert float(s.alphas_cumprod[0]) == pytest.approx(1 - 0.00085, rel=1e-6)
    assert bool((s.alphas_cumprod[1:] < s.alphas_cumprod[:-1]).all())
    # t = 1 -> prev = -19 -> final_alpha_cumprod = abar[0] (set_alpha_to_one False)
    x, e = torch.randn(4, 8), torch.randn(4, 8)
    out = s.step(e, 1, x).prev_sample
    a, ap = s.alphas_cumprod[1], s.alphas_cumprod[0]
    ref = ap.sqrt() * (x - (1 - a).sqrt() * e) / a.sqrt() + (1 - ap).sqrt() * e
    assert torch.allclose(out, ref, atol=1e-6)


def test_quaternion_matrix_kats():
    """docstring KATs of compute.py:40-48 on the batched metric helper"""
    from seeme_b200.metrics import quaternion_rotmat
    q = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0.99810947, 0.06146124, 0, 0]], dtype=torch.float64)
    R_ = quaternion_rotmat(q)
    assert torch.allclose(R_[0], torch.eye(3, dtype=torch.float64))
    assert torch.allclose(R_[1], torch.diag(torch.tensor([1.0, -1, -1], dtype=torch.float64)))
    c, s_ = math.cos(0.123), math.sin(0.123)
    assert torch.allclose(R_[2], torch.tensor([[1, 0, 0], [0, c, -s_], [0, s_, c]], dtype=torch.float64), atol=1e-7)


def test_smpl_kats(smpl_buffers):
    b = smpl_buffers
    F = 3
    betas = 0.5 * torch.randn(F, 10)
    zeros = torch.zeros(F, 69)
    v, j = O.smpl_forward(b, betas, zeros, torch.zeros(F, 3), None)
    v_shaped = b["v_template"][None] + torch.einsum("bl,mkl->bmk", betas, b["shapedirs"])
    assert torch.allclose(v, v_shaped, atol=1e-5)                                   # zero pose => verts == v_shaped
    assert torch.allclose(j, torch.einsum("bik,ji->bjk", v_shaped, b["J_regressor"]), atol=1e-5)
    # rows of lbs_weights sum to one => a pure translation moves everything rigidly
    tr = torch.randn(F, 3)
    pose = 0.3 * torch.randn(F, 69)
    go = 0.3 * torch.randn(F, 3)
    v0, j0 = O.smpl_forward(b, betas, pose, go, None)
    v1, j1 = O.smpl_forward(b, betas, pose, go, tr)
    assert torch.allclose(v1, v0 + tr[:, None], atol=1e-5) and torch.allclose(j1, j0 + tr[:, None], atol=1e-5)
    # rotating only the root rotates all joints rigidly about the root joint
    v2, j2 = O.smpl_forward(b, betas, zeros, go, None)
    Rm = O.batch_rodrigues(go)
    _, jrest = O.smpl_forward(b, betas, zeros, torch.zeros(F, 3), None)
    expect = torch.einsum("fab,fjb->fja", Rm, jrest - jrest[:, :1]) + jrest[:, :1]
    assert torch.allclose(j2, expect, atol=1e-5)
    assert float((b["lbs_weights"] > 0).sum(1).max()) <= 4


def test_image_backbone_restatement_vs_reference_golden():
    """oracle/restate.image_backbone_forward against the UNMODIFIED reference ResNet-50 (EgoHMR/models/resnet.py),
    golden written by oracle/make_golden_image.py from the same seeded state_dict / crops"""
    from seeme_b200 import synthetic as S
    g = np.load(os.path.join(GOLDEN, "resnet50_image.npz"))
    sd = S.resnet50_state(int(g["seed"]))
    x = S.images(int(g["batch"]), int(g["seed"]))
    with torch.no_grad():
        got = O.image_backbone_forward(sd, x)
    assert torch.allclose(got, torch.from_numpy(g["feat"]), atol=1e-5, rtol=1e-5)
