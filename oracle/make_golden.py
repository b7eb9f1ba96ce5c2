"""TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/*.npz`` by running the UNMODIFIED reference
modules from /root/reference (``oracle/ref_modules.py``) on seeded synthetic weights and inputs.

    python -m oracle.make_golden          # build container only; /root/reference must exist

The fixtures store the small inputs next to the reference outputs, so the GPU box (which has no
reference tree) can check both the CPU restatement and the CUDA path against the reference's own
results.  Weights are NOT stored (60 MB): they are regenerated from ``seeme_b200.synthetic`` seeds; a
checksum of every state_dict is stored so a drift in the generator is detected.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_modules as R  # noqa: E402
from oracle import restate as O  # noqa: E402
from seeme_b200 import synthetic as S  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def weights(seed=0):
    return {"denoiser": S.denoiser_state(seed), "vae": S.vae_state(seed), "pointnet": S.pointnet_state(seed),
            "output_scene": S.output_scene_state(seed)}


def checksum(sd) -> float:
    return float(sum(v.double().abs().sum() for v in sd.values()))


def npify(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items() if v is not None}


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    W = weights()
    g = torch.Generator().manual_seed(99)
    out = {"checksum_" + k: np.float64(checksum(v)) for k, v in W.items()}

    # --- denoiser (mld_denoiser.py:151-244), Nc = 2 and Nc = 1, two timesteps ------------------
    den = R.build_denoiser()
    den.load_state_dict(W["denoiser"])
    x = torch.randn(6, 1, 256, generator=g) * 3.0
    enc2 = torch.randn(2, 6, 256, generator=g)
    enc1 = torch.randn(1, 6, 256, generator=g)
    with torch.no_grad():
        out.update(npify({
            "den_x": x, "den_enc2": enc2, "den_enc1": enc1,
            "den_out_t481_nc2": den(sample=x, timestep=torch.tensor(481), encoder_hidden_states=enc2, lengths=[60] * 6)[0],
            "den_out_t1_nc2": den(sample=x, timestep=torch.tensor(1), encoder_hidden_states=enc2, lengths=[60] * 6)[0],
            "den_out_t981_nc1": den(sample=x, timestep=torch.tensor(981), encoder_hidden_states=enc1, lengths=[60] * 6)[0],
        }))

    # --- VAE (mld_vae.py:128-256), ragged lengths -------------------------------------------
    vae = R.build_vae()
    vae.load_state_dict(W["vae"])
    f = torch.randn(4, 60, 75, generator=g)
    lens = [60, 33, 47, 20]
    eps = torch.randn(1, 4, 256, generator=g)
    with torch.no_grad(), R.noise_queue([eps], []):
        z, dist = vae.encode(f, None, lens)
    with torch.no_grad():
        dec = vae.decode(z, lens)
    out.update(npify({"vae_f": f, "vae_lens": torch.tensor(lens), "vae_eps": eps, "vae_z": z, "vae_mu": dist.loc,
                      "vae_std": dist.scale, "vae_dec": dec}))

    # --- scene encoder (respointnet.py:33-59) -------------------------------------------------
    pn = R.build_pointnet()
    pn.load_state_dict(W["pointnet"])
    p = S.egobody_scene(2, 2000, torch.Generator().manual_seed(1))
    with torch.no_grad():
        out.update(npify({"pn_p": p, "pn_out": pn(p), "pn_out_zero": pn(torch.zeros(1, 16, 3))}))

    # --- aa_to_quat (geometry2.py:33-54) and the metric helper KATs (compute.py:40-48) ---------------
    th = torch.randn(16, 3, generator=g)
    out.update(npify({"quat_theta": th, "quat_out": R.aa_to_quat(th)}))
    np.savez_compressed(os.path.join(OUT, "stages.npz"), **out)

    # --- full ego_eval (mld.py:1076-1905) through the unmodified reference function -----------------
    smpl, stats = S.smpl_buffers(), S.norm_stats()
    # name, dataset, condition, guidance, B, T (MOTION_LENGTH), ESTIMATE, TEST.GLOBAL_ORIENT_PRED.  The last three rows are the
    # config_mld_interactee.yaml protocol (ESTIMATE: interactee, MOTION_LENGTH: 1, guidance 1.0; lines 18,20,73) through the
    # working scene / scene+interactee branches, and the GLOBAL_ORIENT_PRED: False switch (mld.py:1501-1505)
    for name, dataset, cond, gs, B, T, est, pgo in (
            ("egobody_cfg", "egobody", ("text", "scene", "interactee"), 7.5, 2, 60, "wearer", True),
            ("egobody_nocfg", "egobody", ("text", "scene", "interactee"), 1.0, 2, 60, "wearer", True),
            ("gimo_cfg", "gimo", ("text", "scene"), 7.5, 2, 60, "wearer", True),
            ("interactee_T1_scene", "egobody", ("text", "scene"), 1.0, 3, 1, "interactee", True),
            ("interactee_T1_scene_int_cfg", "egobody", ("text", "scene", "interactee"), 7.5, 3, 1, "interactee", True),
            ("egobody_gt_orient", "egobody", ("text", "scene", "interactee"), 7.5, 2, 60, "wearer", False)):
        batch = S.make_batch(B, n_points=1000, T=T, ragged=T > 1, dataset=dataset)
        gg = torch.Generator().manual_seed(7)
        noise = {"eps_int": torch.randn(1, B, 256, generator=gg), "eps_unc": torch.randn(1, B, 256, generator=gg),
                 "x_T": torch.randn(B, 1, 256, generator=gg)}
        c = R.make_carrier(W, smpl, stats, condition=cond, guidance_scale=gs, dataset=dataset, estimate=est, pred_global_orient=pgo)
        normal = []
        if "interactee" in cond:
            normal = [noise["eps_int"], noise["eps_unc"]] if gs > 1 else [noise["eps_int"]]
        torch.manual_seed(5)    # the no-interactee branch draws torch.rand_like after sampling
        with torch.no_grad(), R.noise_queue(normal, [noise["x_T"]]):
            ref = c._ref_ego_eval(batch)
        keep = {k: ref[k] for k in ("m_ref", "m_rst", "joints_ref", "joints_rst", "orientation_quat_rst", "orientation_quat_ref")}
        if "interactee" in cond:
            keep.update({k: ref[k] for k in ("joints_interactee", "root_interactee", "orientation_quat_int")})
        keep["lengths"] = torch.tensor(ref["lengths"])
        keep.update({"noise_" + k: v for k, v in noise.items()})
        keep.update({"cfg_T": np.int64(T), "cfg_estimate": np.str_(est), "cfg_pred_global_orient": np.bool_(pgo)})
        np.savez_compressed(os.path.join(OUT, f"ego_eval_{name}.npz"), **npify(keep))
        print(name, "joints_rst absmax", float(ref["joints_rst"].abs().max()))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
