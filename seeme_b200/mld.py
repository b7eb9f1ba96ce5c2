"""``MLD`` -- the reference's LightningModule surface (``mld/models/modeltype/mld.py``,
``mld/models/modeltype/base.py``) for the inference hot path, running on sm_100a kernels.

Kept from the reference: constructor ``MLD(cfg, datamodule, **kw)``, attribute names, ``state_dict``
prefixes (``denoiser.*``, ``vae.*``, ``proscene.scene_enc.*``, ``output_scene.1.*``, ``smpl_model.*``),
``forward(batch)``, ``test_step`` / ``allsplit_step`` / ``ego_eval(batch) -> rs_set`` and
``_diffusion_reverse(encoder_hidden_states, lengths)``, including the quirks that are part of the
contract: CFG half ordering (scene: cond first, interactee: uncond first),
GT betas for the predicted mesh, ``TEST.GLOBAL_ORIENT_PRED`` switch, GIMO's 63-d body pose with two
zeroed hand joints, un-masked padded frames, float64 ``m_ref``/``m_rst``.  Deviations: ``save_for_edo``
is off (App. D1), ``rs_set`` always carries ``list_names`` (App. D2), training paths are not built.

It is a plain ``nn.Module`` (pytorch_lightning is not required); when Lightning is installed it can be
wrapped or used as the ``LightningModule`` base by passing ``base=`` to ``make_lightning_class``.
"""
from __future__ import annotations

import os
import time
import weakref
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from .config import instantiate_from_config
from .modules import MldDenoiser, MldVae, ProHMRScene, SMPL, time_sinusoid


class PendingEval:
    """Result of ``MLD.ego_eval_async``: the ``rs_set`` tensors are being produced on ``stream``.

    With sampler coalescing (``MLD.sampler_group`` > 1) a batch's decode stage is only enqueued once its group's shared
    sampler chain has been launched; every accessor below first closes the group if it is still open."""

    def __init__(self, model, stream, slot):
        self._model, self.stream, self.slot = model, stream, slot
        self._rs_set = self.last_vertices = self.last_latent = self.event = None
        self._ctx = self._enc_event = self._error = None
        self._callbacks = []

    def _ready(self):
        if self._error is not None:
            raise self._error
        if self._rs_set is None:
            self._model._flush_group()
        return self

    @property
    def rs_set(self):
        return self._ready()._rs_set

    def then(self, fn):
        """``fn(rs_set)`` runs on the slot's stream right after the batch's last kernel has been enqueued (e.g. a
        device-to-host copy of a result); ``event`` / ``synchronize`` cover what it enqueues."""
        if self._rs_set is not None:
            with torch.cuda.stream(self.stream):
                fn(self._rs_set)
                self.event = torch.cuda.Event()
                self.event.record(self.stream)
        else:
            self._callbacks.append(fn)
        return self

    def result(self):
        """make the caller's current stream wait for the slot, then hand out the rs_set (device tensors)"""
        self._ready()
        torch.cuda.current_stream().wait_event(self.event)
        return self._rs_set

    def synchronize(self):
        """block the host until the slot's work (including any device-to-host copies enqueued on ``stream``) is done"""
        self._ready()
        self.event.synchronize()
        return self._rs_set


class MLD(nn.Module):
    def __init__(self, cfg, datamodule, smpl_buffers: Optional[Dict[str, torch.Tensor]] = None, **kwargs):
        super().__init__()
        self.cfg = cfg
        self.times: List[float] = []                                     # base.py:26
        self.stage = cfg.TRAIN.STAGE
        self.condition = list(cfg.model.condition)
        self.is_vae = cfg.model.vae
        self.predict_epsilon = cfg.TRAIN.ABLATION.get("PREDICT_EPSILON", True)
        self.name_dataset = cfg.DATASET_NAME
        self.njoints = cfg.model.njoints
        self.latent_dim = list(cfg.model.latent_dim)
        self.guidance_scale = float(cfg.model.guidance_scale)
        self.guidance_uncodp = cfg.model.guidance_uncondp
        self.datamodule = datamodule
        self.estimate = cfg.ESTIMATE
        self.pred_global_orient = cfg.TEST.GLOBAL_ORIENT_PRED
        self.pred_betas = cfg.TEST.BETAS_PRED
        self.global_orient_egoego = cfg.TEST.GLOBAL_ORIENT_EGOEGO
        self.transl_egoego = cfg.TEST.TRANSL_EGOEGO
        self.pose_estimation_task = cfg.TEST.POSE_ESTIMATION_TASK
        self.see_future = cfg.TEST.SEE_FUTURE
        self.predict_transl = cfg.TRAIN.ABLATION.PREDICT_TRANSL
        if self.name_dataset == "egobody":
            self.nfeats = 75 if self.predict_transl else 72              # mld.py:119-123
        elif self.name_dataset == "gimo":
            self.nfeats = 69 if self.predict_transl else 66
        else:
            raise NotImplementedError(f"dataset {self.name_dataset!r}: only egobody and gimo are on the SEE-ME path")
        self.data_type = cfg.DATA_TYPE
        self.save_for_edo = False                                        # App. D1
        unsupported = []
        if self.data_type != "angle":
            unsupported.append("DATA_TYPE must be 'angle'")
        if self.stage != "diffusion":
            unsupported.append("only TRAIN.STAGE=diffusion (stage-2 inference) is built")
        if not self.predict_transl:
            unsupported.append("PREDICT_TRANSL must be True")
        if self.pred_betas or self.global_orient_egoego or self.transl_egoego or self.pose_estimation_task or self.see_future:
            unsupported.append("BETAS_PRED / *_EGOEGO / POSE_ESTIMATION_TASK / SEE_FUTURE variants are not built")
        if "image" in self.condition and self.guidance_scale > 1.0:
            # the image branches build no unconditional scene / image embedding (mld.py:1078-1099), so under CFG the
            # reference's torch.cat of [1,2B,256] and [1,B,256] tokens fails (:1299); config_mld_interactee.yaml:120 runs 1.0
            unsupported.append("image conditioning runs without classifier-free guidance (guidance_scale <= 1.0), as in the reference")
        if unsupported:
            raise NotImplementedError("MLD (sm_100a): " + "; ".join(unsupported))

        max_batch = int(kwargs.get("max_batch", max(int(cfg.TEST.BATCH_SIZE), 1)))
        n_points = int(kwargs.get("max_points", 20000))
        max_frames = int(cfg.get("MOTION_LENGTH", 60))
        # body model: smplx.SMPL(model_path=cfg.model.smpl_path, ...) in the reference (mld.py:151-153)
        self.smpl_model = SMPL(smpl_buffers if smpl_buffers is not None else cfg.model.smpl_path,
                               batch_size=cfg.TRAIN.BATCH_SIZE, gender="neutral", max_frames=max_batch * max_frames)
        try:
            self.vae_type = cfg.model.vae_type
        except AttributeError:
            self.vae_type = cfg.model.motion_vae.target.split(".")[-1].lower().replace("vae", "")   # -> "mld"
        if "scene" in self.condition:
            # scene_precision: operand format of the scene encoder's per-point GEMMs ("fp16-fused" = 16, the
            # default; "split-bf16" = 3 tracks the fp32 reference to ~1e-5 at 3x the MMA count)
            sp = kwargs.get("scene_precision", cfg.model.get("scene_precision", "fp16-fused"))
            sp = {"fp16-fused": 16, "fp16-fused-smem": 17, "split-bf16": 3, "bf16": 1, "fp32": 0}.get(sp, sp)
            self.scene_precision = int(sp)
            self.proscene = ProHMRScene(cfg.get("PROSCENE"), max_batch=max_batch, max_points=n_points, precision=int(sp),
                                        with_backbone="image" in self.condition)
            self.output_scene = nn.Sequential(nn.ReLU(), nn.Linear(512, 256))                        # mld.py:257-261
        elif "image" in self.condition:
            self.proscene = ProHMRScene(cfg.get("PROSCENE"), max_batch=max_batch, max_points=128, with_backbone=True)
        if "image" in self.condition:
            self.output_images = nn.Sequential(nn.ReLU(), nn.Linear(2048, 256))                      # mld.py:251-255
        mv = cfg.model.motion_vae
        mv_params = dict(mv.get("params", {}))
        mv_params.setdefault("max_batch", max_batch)
        mv_params.setdefault("max_frames", max_frames)
        self.vae = instantiate_from_config({"target": mv["target"], "params": mv_params})
        dn = cfg.model.denoiser
        dn_params = dict(dn.get("params", {}))
        dn_params.setdefault("max_rows", 2 * max_batch)
        self.denoiser = instantiate_from_config({"target": dn["target"], "params": dn_params})
        self.scheduler = instantiate_from_config(cfg.model.scheduler)
        self.noise_scheduler = instantiate_from_config(cfg.model.noise_scheduler) if "noise_scheduler" in cfg.model else None
        for p in self.parameters():
            p.requires_grad = False
        self.metrics_dict = list(cfg.METRIC.TYPE)
        self.configure_metrics(kwargs.get("metrics"))
        self.do_classifier_free_guidance = self.guidance_scale > 1.0       # mld.py:331
        self.renorm = datamodule.renorm                                     # mld.py:333
        # which bodies get the 6890-vertex skinning: "rst" (predicted, default), "all", or "none";
        # the reference skins all three and discards the vertices (only joints reach rs_set)
        self.compute_vertices = kwargs.get("compute_vertices", "rst")
        # concurrent sub-batches (see ego_eval): at most `lanes` lanes of at least `min_lane_batch` sequences each
        self.lanes = int(kwargs.get("lanes", cfg.model.get("lanes", 1)))
        self.min_lane_batch = int(kwargs.get("min_lane_batch", cfg.model.get("min_lane_batch", 32)))
        # batches in flight for ego_eval_async / run_test_batches
        # (default: 32 slots for batches of 256 sequences; fewer for larger ones -- a slot's handles and its in-flight result,
        # 6 890 x 60 vertices per sequence, take ~10 MB per sequence of capacity: 32 x 256 sequences = 76 GB of the 180 GB -- and
        # 8 for batches below 256, whose GPU time per batch is close to the host's submission time: measured at 64 sequences per
        # batch, 8 slots with the cluster sampler give 4.0 ms per batch, 32 slots 4.4-6.3 ms with host stalls of up to 260 ms
        # whenever a slot's stream pool misses in the caching allocator while 32 samplers are in flight)
        default_depth = 8 if max_batch < 256 else min(32, max(4, 8192 // max_batch))
        self.pipeline_depth = int(kwargs.get("pipeline_depth", cfg.model.get("pipeline_depth", os.environ.get("SEEME_PIPELINE_DEPTH", default_depth))))
        # ego_eval_async / run_test_batches: consecutive batches whose 50-step sampler runs as ONE chain over all their rows
        # (the chain is latency-bound: 3 750 dependent kernels take the same ~27 ms for 512 or 2 048 rows)
        # sampler back-end (include/seeme_b200.h: seeme_denoiser_set_backend): "persistent" = one launch of the 8-CTA-cluster kernel
        # per run, "tile" = one launch of the one-CTA-per-128-rows kernel, "graph" = the CUDA graph of small kernels, "auto" =
        # persistent for a single batch (ego_eval: lowest latency), tile inside the batch pipeline (ego_eval_async with several
        # batches in flight: least SM time next to the scene encoder)
        self.sampler_backend = str(kwargs.get("sampler_backend", cfg.model.get("sampler_backend", os.environ.get("SEEME_SAMPLER_BACKEND", "auto"))))
        self.sampler_group = int(kwargs.get("sampler_group", cfg.model.get("sampler_group", os.environ.get("SEEME_SAMPLER_GROUP", 1))))
        # "auto" back-end: the cluster kernel while (batches in flight) x (8 SMs per 128-row tile) stays within this many SMs
        self.persistent_sm_budget = int(kwargs.get("persistent_sm_budget", cfg.model.get("persistent_sm_budget", os.environ.get("SEEME_PERSISTENT_SM_BUDGET", 64))))
        # pipeline slots share this many scene-encoder handles (2.6 GB of workspace each at 128 clouds x 20 000 points): an encoder
        # fills every SM, so two of them never overlap anyway; 0 = one handle per slot
        self.encoder_handles = int(kwargs.get("encoder_handles", cfg.model.get("encoder_handles", os.environ.get("SEEME_ENCODER_HANDLES", 4))))
        self.last_vertices: Dict[str, torch.Tensor] = {}
        self._uncond_scene = None
        self.eval()

    # ------------------------------------------------------------------------------------------
    def configure_metrics(self, metrics=None):
        """base.py:160-173.  ``metrics`` may carry ready metric objects ({"EgoMetric": obj}); otherwise the
        in-repo batched EgoMetric is used (same state sums as mld/models/metrics/compute.py)."""
        from .metrics import EgoMetric, MRMetric
        for m in self.metrics_dict:
            if metrics and m in metrics:
                obj = metrics[m]
            elif m == "EgoMetric":
                obj = EgoMetric(njoints=self.njoints, dist_sync_on_step=self.cfg.METRIC.get("DIST_SYNC_ON_STEP", True))
            elif m == "MRMetrics":                                              # base.py:180-185
                obj = MRMetric(njoints=self.njoints, jointstype=self.cfg.get("DATASET", {}).get("JOINT_TYPE", "humanml3d"),
                               dist_sync_on_step=self.cfg.METRIC.get("DIST_SYNC_ON_STEP", True))
            else:
                raise NotImplementedError(f"Do not support Metric Type {m}")
            object.__setattr__(self, m, obj)     # metrics hold no persistent state

    def _stats(self, device):
        cache = self.__dict__.setdefault("_stats_cache", {})
        if device not in cache:
            mean, std = np.asarray(self.datamodule.mean), np.asarray(self.datamodule.std)
            mean = mean[0] if mean.ndim == 2 else mean               # renorm reads row 0 (EgoBody.py:152-154)
            std = std[0] if std.ndim == 2 else std
            cache[device] = (torch.as_tensor(mean, dtype=torch.float64).to(device), torch.as_tensor(std, dtype=torch.float64).to(device))
        return cache[device]

    # ------------------------------------------------------------------------------------------
    def _diffusion_reverse(self, encoder_hidden_states, lengths=None, latents=None, op=None):
        """mld.py:432-511.  encoder_hidden_states [B',Nc,256] (B' = 2B under CFG) -> [1,B,256].
        ``latents`` ([B,1,256]) injects the initial noise; when None it is drawn with ``torch.randn``
        exactly where the reference draws it (:449-453)."""
        bsz = encoder_hidden_states.shape[0]
        if self.do_classifier_free_guidance:
            bsz = bsz // 2
        dev = encoder_hidden_states.device
        if latents is None:
            latents = torch.randn((bsz, self.latent_dim[0], self.latent_dim[-1]), device=dev, dtype=torch.float)
        latents = latents * self.scheduler.init_noise_sigma
        n_steps = self.cfg.model.scheduler.num_inference_timesteps
        if self.scheduler.num_inference_steps != n_steps:
            self.scheduler.set_timesteps(n_steps)
            self.__dict__["_coef"] = None
        ts = [int(t) for t in self.scheduler.timesteps]
        if self.__dict__.get("_coef") is None:
            self.__dict__["_coef"] = self.scheduler.step_coefficients()
            self.__dict__["_sinus"] = time_sinusoid(self.scheduler.timesteps)
        op = self.denoiser.op if op is None else op
        backend = self.sampler_backend
        if backend == "auto":
            # measured on B200: a single batch is fastest on the cluster kernel (16.8 ms per 512 rows on 32
            # SMs); in a FULL batch pipeline the SM time counts, not the latency, and the one-CTA-per-tile kernel holds 4 SMs per
            # 512 rows (20.2k sequences/s at depth 32 against 18.0k with the kernel graph and 16.0k with the cluster kernel; the
            # same order at 128 and 64 sequences per batch).  While few batches are in flight (pipeline filling, or an epoch of
            # a few small batches as in the interactee protocol) the clusters of all of them still fit next to each other
            # (8 SMs per 128-row tile) and the lower latency wins; a shallow pipeline (<= 8 slots: the small-batch default) always
            # takes the cluster kernel.
            in_pipe = self.__dict__.get("_in_pipeline", False) and int(self.pipeline_depth) > 1
            tiles = -(-encoder_hidden_states.shape[0] // 128)
            crowded = int(self.pipeline_depth) > 8 and (self.__dict__.get("_n_inflight", 0) + 1) * tiles * 8 > int(self.persistent_sm_budget)
            backend = "tile" if in_pipe and crowded else "persistent"
        if os.environ.get("SEEME_SAMPLER") != "graph":
            op.set_backend(backend)
        # the key lives on the kernel-side handle object (one per lane / slot), not in an id()-keyed dict: a rebuilt handle
        # starts without it, and DenoiserOp.forward / set_time_table keep it in step with the C-side table
        if getattr(op, "table_key", None) != tuple(ts):
            op.set_time_table(ts, self.__dict__["_sinus"])
        cond = encoder_hidden_states.permute(1, 0, 2).contiguous()         # [Nc,B',256] as the denoiser receives it
        z = op.sample(latents.reshape(bsz, 256), cond, self.guidance_scale, ts, self.__dict__["_coef"])
        return z.view(bsz, 1, 256).permute(1, 0, 2)

    # ------------------------------------------------------------------------------------------
    def _length_index(self) -> int:
        """position of ``length`` in the dataset's item tuple (dataset.py:1778-1794)"""
        if "image" in self.condition:
            return 6 if "scene" in self.condition else 5
        return 5 if "scene" in self.condition else 4

    def _encode_scene(self, scene):
        from . import modules as _m
        lane = _m._LANE[0]
        shared = 1000 <= lane < 2000 and int(self.encoder_handles) > 0      # a pipeline slot (ego_eval_async)
        if shared:
            # the handle's workspace is used inside the call only (the embedding is a fresh tensor), so slots k, k + E, ... take
            # turns on handle k % E: the slot's stream waits for the previous user's encoder
            k = (lane - 1000) % int(self.encoder_handles)
            _m._LANE[0] = 3000 + k
            try:
                op = self.proscene.scene_enc.op(self.output_scene)
            finally:
                _m._LANE[0] = lane
            events = self.__dict__.setdefault("_enc_events", {})
            if k in events:
                torch.cuda.current_stream(scene.device).wait_event(events[k])
        else:
            op = self.proscene.scene_enc.op(self.output_scene)
        emb = op(scene.float())                                           # output_scene(encode_scene(.)) fused
        unc = None
        if self.do_classifier_free_guidance:
            # encode_scene(zeros) is input independent (every point identical -> the max-pool is that point,
            # SURVEY App. H8): computed once on a tiny all-zero cloud and cached per packed-weights handle
            cache = self._uncond_scene if isinstance(self._uncond_scene, dict) else {}
            self._uncond_scene = cache
            if id(op) not in cache or cache[id(op)][0] is not op:
                cache[id(op)] = (op, op(torch.zeros(1, 8, 3, device=scene.device)))
            unc = cache[id(op)][1].expand(emb.shape[0], -1)
        if shared:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(scene.device))
            events[k] = ev
        if unc is None:
            return emb[None]
        return torch.cat([emb, unc], dim=0)[None]                          # mld.py:1157-1158 (COND first)

    def ego_eval(self, batch, noise: Optional[Dict[str, torch.Tensor]] = None):
        """mld.py:1076-1905 (scene / scene+interactee / interactee-only branches).
        ``noise`` = {"eps_int", "eps_unc" [1,B,256], "x_T" [B,1,256]} injects the three draws; missing
        entries are drawn with torch.randn in the reference's order.

        Large batches are split into ``self.lanes`` contiguous sub-batches that run concurrently on their own CUDA
        streams with their own kernel-side handles: the 50-step sampler is a latency-bound chain of small kernels that
        leaves most SMs idle, so one lane's sampler overlaps the other lanes' scene encoder / VAE / SMPL work.  Every
        stage is per-sample, so the results are those of the unsplit batch."""
        noise = dict(noise or {})
        B = batch[0].shape[0]
        lanes = max(1, min(int(self.lanes), B // max(int(self.min_lane_batch), 1)))
        if lanes <= 1 or not batch[0].is_cuda:
            return self._ego_eval_one(batch, noise)
        dev = batch[0].device
        # the draws happen once, for the whole batch and in the reference's order (mld_vae.py:190-192, mld.py:449-453)
        if "interactee" in self.condition:
            if noise.get("eps_int") is None:
                noise["eps_int"] = torch.randn(1, B, 256, device=dev, dtype=torch.float32)
            if self.do_classifier_free_guidance and noise.get("eps_unc") is None:
                noise["eps_unc"] = torch.randn(1, B, 256, device=dev, dtype=torch.float32)
        if noise.get("x_T") is None:
            noise["x_T"] = torch.randn((B, self.latent_dim[0], self.latent_dim[-1]), device=dev, dtype=torch.float)
        from . import modules as _m
        from .dist import shard_range
        main = torch.cuda.current_stream(dev)
        # one device->host read of the lengths for the whole batch (mld.py:1264); every lane decodes to max(lengths)
        lengths_all = batch[self._length_index()].long().reshape(-1).tolist()
        t_max = int(max(lengths_all))
        streams = self.__dict__.setdefault("_lane_streams", {})
        outs = []
        for k in range(lanes):
            lo, hi = shard_range(B, k, lanes)
            st = streams.get((dev, k))
            if st is None:
                st = streams[(dev, k)] = torch.cuda.Stream(device=dev)
            sub = tuple((x[lo:hi] if torch.is_tensor(x) else x) for x in batch)
            sub_noise = {"eps_int": None, "eps_unc": None, "x_T": noise["x_T"][lo:hi]}
            for key in ("eps_int", "eps_unc"):
                if noise.get(key) is not None:
                    sub_noise[key] = noise[key][:, lo:hi].contiguous()
            st.wait_stream(main)
            _m._LANE[0] = k
            try:
                with torch.cuda.stream(st):
                    outs.append(self._ego_eval_one(sub, sub_noise, defer_random=True, t_max=t_max, lengths=lengths_all[lo:hi]))
            finally:
                _m._LANE[0] = 0
        # the lanes' outputs are consumed on the caller's stream after this join; the next call makes every lane stream
        # wait for the caller's stream again before reusing memory, so no record_stream bookkeeping is needed
        for k in range(lanes):
            main.wait_stream(streams[(dev, k)])
        rs_set = {}
        for key, v0 in outs[0].items():
            if torch.is_tensor(v0):
                rs_set[key] = torch.cat([o[key] for o in outs], dim=0)
            elif key == "lengths":
                rs_set[key] = [x for o in outs for x in o[key]]
            else:
                rs_set[key] = v0
        if "interactee" not in self.condition:                                # mld.py:1572-1574 (RNG side effect kept)
            joints_int = torch.rand_like(rs_set["joints_rst"])
            rs_set["joints_interactee"] = joints_int
            rs_set["root_interactee"] = joints_int[:, :, 0:1, :]
            rs_set["orientation_quat_int"] = torch.rand_like(rs_set["orientation_quat_rst"])
        lv = [o.pop("_last_vertices") for o in outs]
        self.last_vertices = {k: (None if lv[0][k] is None else torch.cat([d[k] for d in lv], dim=0)) for k in lv[0]}
        self.last_latent = torch.cat([o.pop("_last_latent") for o in outs], dim=1)
        rs_set.pop("_last_vertices", None)
        rs_set.pop("_last_latent", None)
        return rs_set

    # ------------------------------------------------------------------------------------------
    def ego_eval_async(self, batch, noise: Optional[Dict[str, torch.Tensor]] = None) -> "PendingEval":
        """Enqueue ``ego_eval(batch)`` on the next pipeline slot and return immediately.

        The 50-step sampler is a ~26 ms latency-bound chain of small kernels that leaves most SMs idle, while the scene
        encoder / VAE / SMPL stages are throughput-bound; with ``pipeline_depth`` batches in flight -- each slot has its
        own CUDA stream and its own kernel-side handles (workspaces, sampler graph) -- batch k+1's scene encoder runs
        under batch k's sampler.  CPU tensors in ``batch`` / ``noise`` (pinned host memory) are copied on the slot's
        stream, so the copies overlap other slots' compute as well.  ``PendingEval.result()`` makes the caller's current
        stream wait for the slot and returns the ``rs_set``; a batch takes the first idle slot (round robin when all are busy), so
        at most ``pipeline_depth`` results should be outstanding."""
        from . import modules as _m
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("MLD (sm_100a): parameters are on the CPU; seeme_b200 runs on CUDA only")
        depth = max(1, int(self.pipeline_depth))
        # slot choice: the lowest-numbered slot whose previous batch has finished (its handles and workspaces exist already: an
        # epoch of a few batches keeps reusing the same warm slots), else a slot never used, else round robin
        last = self.__dict__.setdefault("_slot_last", {})
        slot = None
        for k in range(depth):
            ref = last.get(k)
            if ref is None:
                if slot is None:
                    slot = k                                         # never used: taken unless a warm idle slot follows
                continue
            p = ref()
            if p is None or p._error is not None or (p.event is not None and p.event.query()):
                slot = k
                break
        if slot is None:
            slot = self.__dict__.get("_next_slot", 0) % depth
            self.__dict__["_next_slot"] = slot + 1
        streams = self.__dict__.setdefault("_slot_streams", {})
        st = streams.get((dev, slot))
        if st is None:
            st = streams[(dev, slot)] = torch.cuda.Stream(device=dev)
        st.wait_stream(torch.cuda.current_stream(dev))          # device inputs were produced on the caller's stream
        lane = _m._LANE[0]
        _m._LANE[0] = 1000 + slot                                # handles of this slot (distinct from the lanes' handles)
        # the Python list of lengths (mld.py:1264) is read BEFORE anything is enqueued: from the host tensor when the batch
        # is host-resident, else once per distinct device tensor (a .tolist() after the scene encoder has been enqueued
        # would block the host until that slot's encoder has finished)
        length_t = batch[self._length_index()]
        if torch.is_tensor(length_t) and length_t.is_cuda:
            # keyed by the live tensor OBJECT (weak reference) and its version counter -- not by address: the caching allocator
            # hands the address of a freed batch to the next one
            cache = self.__dict__.setdefault("_length_cache", {})
            ent = cache.get(id(length_t))
            if ent is None or ent[0]() is not length_t or ent[1] != length_t._version:
                if len(cache) > 64:
                    cache.clear()
                ent = cache[id(length_t)] = (weakref.ref(length_t), length_t._version, length_t.long().reshape(-1).tolist())
            lengths_host = ent[2]
        else:
            lengths_host = torch.as_tensor(length_t).long().reshape(-1).tolist()
        pend = PendingEval(self, st, slot)
        # batches submitted earlier and not yet finished on the device (the sampler back-end policy looks at it)
        # (weak references: a result the caller has dropped must not be kept alive -- its vertices are 1.3 GB)
        def _busy(ref):
            p = ref()
            return p is not None and p._error is None and (p.event is None or not p.event.query())
        outstanding = [r for r in self.__dict__.get("_outstanding", []) if _busy(r)]
        self.__dict__["_n_inflight"] = len(outstanding)
        outstanding.append(weakref.ref(pend))
        self.__dict__["_outstanding"] = outstanding
        last[slot] = weakref.ref(pend)
        try:
            with torch.cuda.stream(st):
                b = tuple((x.to(dev, non_blocking=True) if torch.is_tensor(x) and not x.is_cuda else x) for x in batch)
                n = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) and not v.is_cuda else v) for k, v in (noise or {}).items()}
                # device-resident inputs were allocated on the caller's stream but are read here, on the slot's stream, up to
                # tens of ms later (the decode stage reads feats / transl / beta after the sampler): tell the caching allocator,
                # or a caller that drops the batch could see its memory handed out while the slot still reads it
                for x in list(b) + list(n.values()):
                    if torch.is_tensor(x) and x.is_cuda:
                        x.record_stream(st)
                pend._ctx = self._stage_encode(b, n, lengths_host)
                pend._enc_event = torch.cuda.Event()
                pend._enc_event.record(st)
        finally:
            _m._LANE[0] = lane
        group = self.__dict__.setdefault("_group_open", [])
        group.append(pend)
        if len(group) >= max(1, min(int(self.sampler_group), depth)):
            self._flush_group()
        return pend

    def prepare_pipeline(self, n_slots: Optional[int] = None):
        """Create the kernel-side handles (packed weights, workspaces, sampler tapes) of ``n_slots`` pipeline slots (default:
        ``pipeline_depth``) up front.  A handle set costs ~150 ms of host time to build; without this call a slot is built
        the first time a batch finds every existing slot busy, i.e. somewhere inside the first epochs."""
        from . import modules as _m
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("MLD (sm_100a): parameters are on the CPU; seeme_b200 runs on CUDA only")
        n = max(1, min(int(self.pipeline_depth), int(n_slots if n_slots is not None else self.pipeline_depth)))
        lane = _m._LANE[0]
        streams = self.__dict__.setdefault("_slot_streams", {})
        last = self.__dict__.setdefault("_slot_last", {})
        try:
            with torch.cuda.device(dev):
                for k in range(n):
                    _m._LANE[0] = 1000 + k
                    if (dev, k) not in streams:
                        streams[(dev, k)] = torch.cuda.Stream(device=dev)
                    self.vae.op, self.denoiser.op, self.smpl_model.op
                    if "image" in self.condition:
                        self.proscene.backbone.op(self.output_images)
                    last.setdefault(k, lambda: None)                     # "used and idle" for the slot choice
                if "scene" in self.condition:
                    for k in range(min(n, max(1, int(self.encoder_handles))) if int(self.encoder_handles) > 0 else n):
                        _m._LANE[0] = (3000 if int(self.encoder_handles) > 0 else 1000) + k
                        self.proscene.scene_enc.op(self.output_scene)
        finally:
            _m._LANE[0] = lane
        torch.cuda.synchronize(dev)
        return self

    def _flush_group(self):
        """Launch the sampler of the open group -- ONE chain over the rows of all its batches, on a group stream with its
        own denoiser handle -- and enqueue every member's decode stage (VAE decode, SMPL) on the member's slot stream."""
        from . import modules as _m
        members = self.__dict__.get("_group_open") or []
        if not members:
            return
        self.__dict__["_group_open"] = []
        dev = members[0]._ctx["cond_emb"].device
        lane = _m._LANE[0]
        try:
            self._flush_members(members, dev)
        except BaseException as exc:
            # members whose decode stage was never enqueued would otherwise fail later with an AttributeError on their
            # missing event, hiding this error: they re-raise it from result() / synchronize()
            for m in members:
                if m._rs_set is None:
                    m._error = exc
            raise
        finally:
            _m._LANE[0] = lane

    def _flush_members(self, members, dev):
        from . import modules as _m
        self.__dict__["_in_pipeline"] = True
        try:
            self._flush_members_inner(members, dev)
        finally:
            self.__dict__["_in_pipeline"] = False

    def _flush_members_inner(self, members, dev):
        from . import modules as _m
        if True:
            if len(members) == 1:
                m = members[0]
                _m._LANE[0] = 1000 + m.slot
                with torch.cuda.stream(m.stream):
                    z = self._diffusion_reverse(m._ctx["cond_emb"].permute(1, 0, 2), m._ctx["lengths"], latents=m._ctx["x_T"])
                zs = [z]
            else:
                k = self.__dict__.get("_group_count", 0)
                self.__dict__["_group_count"] = k + 1
                n_gs = max(2, -(-max(1, int(self.pipeline_depth)) // len(members)) + 1)
                gstreams = self.__dict__.setdefault("_group_streams", {})
                gs = gstreams.get((dev, k % n_gs))
                if gs is None:
                    gs = gstreams[(dev, k % n_gs)] = torch.cuda.Stream(device=dev)
                sizes = [m._ctx["x_T"].shape[0] for m in members]
                for m in members:
                    gs.wait_event(m._enc_event)
                _m._LANE[0] = 2000 + (k % n_gs)
                with torch.cuda.stream(gs):
                    conds = [m._ctx["cond_emb"] for m in members]                     # [Nc, B or 2B, 256] each
                    for c in conds:
                        c.record_stream(gs)
                    if self.do_classifier_free_guidance:                               # the kernel pairs row i with row i + B_total
                        cond_all = torch.cat([c[:, :b] for c, b in zip(conds, sizes)] + [c[:, b:] for c, b in zip(conds, sizes)], dim=1)
                    else:
                        cond_all = torch.cat(conds, dim=1)
                    for m in members:
                        m._ctx["x_T"].record_stream(gs)
                    lat_all = torch.cat([m._ctx["x_T"] for m in members], dim=0)
                    rows = cond_all.shape[1]
                    cap = -(-rows // 512) * 512
                    z_all = self._diffusion_reverse(cond_all.permute(1, 0, 2), None, latents=lat_all, op=self.denoiser.op_rows(cap))
                    ev_z = torch.cuda.Event()
                    ev_z.record(gs)
                zs, o = [], 0
                for m, b in zip(members, sizes):
                    m.stream.wait_event(ev_z)
                    z_all.record_stream(m.stream)
                    zs.append(z_all[:, o:o + b])
                    o += b
            for m, z in zip(members, zs):
                _m._LANE[0] = 1000 + m.slot
                with torch.cuda.stream(m.stream):
                    rs = self._stage_decode(m._ctx, z.contiguous(), defer_random=True)
                    if "interactee" not in self.condition:            # mld.py:1572-1574 (RNG side effect kept)
                        joints_int = torch.rand_like(rs["joints_rst"])
                        rs["joints_interactee"] = joints_int
                        rs["root_interactee"] = joints_int[:, :, 0:1, :]
                        rs["orientation_quat_int"] = torch.rand_like(rs["orientation_quat_rst"])
                    m.last_vertices, m.last_latent = rs.pop("_last_vertices"), rs.pop("_last_latent")
                    for fn in m._callbacks:
                        fn(rs)
                    m._callbacks = []
                    m.event = torch.cuda.Event()
                    m.event.record(m.stream)
                    m._rs_set, m._ctx = rs, None

    def _encode_uncond(self, B: int, T: int, nfeats: int, lengths, eps, dev):
        """``vae.encode(zeros_like(feats), lengths)`` of the CFG branch (mld.py:1280-1290).  The input is all zeros, so the
        posterior (mu, std) of a row depends on its length only: one zero sequence per DISTINCT length is encoded and the
        rows are gathered -- every row goes through the same arithmetic as in the full batch -- and the reparameterised
        sample ``mu + eps * std`` (torch.distributions.Normal.rsample) uses the caller's / a fresh ``eps`` [1,B,256]."""
        uniq = sorted(set(int(x) for x in lengths))
        if eps is None:
            eps = torch.randn(1, B, self.latent_dim[-1], device=dev, dtype=torch.float32)
        U = len(uniq)
        ulen = torch.tensor(uniq, dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
        _, dist = self.vae.encode(torch.zeros(U, T, nfeats, device=dev), None, uniq,
                                  eps=torch.zeros(1, U, self.latent_dim[-1], device=dev), lengths_dev=ulen)
        if U == 1:
            mu, std = dist.loc, dist.scale                                   # broadcast over the batch
        else:
            pos = {l: i for i, l in enumerate(uniq)}
            idx = torch.tensor([pos[int(x)] for x in lengths], dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
            mu, std = dist.loc.index_select(1, idx), dist.scale.index_select(1, idx)
        return mu + eps * std

    def _ego_eval_one(self, batch, noise, defer_random: bool = False, t_max: Optional[int] = None, lengths=None):
        """one (sub-)batch on the current stream with the current lane's handles"""
        ctx = self._stage_encode(batch, noise, lengths)
        z = self._diffusion_reverse(ctx["cond_emb"].permute(1, 0, 2), ctx["lengths"], latents=ctx["x_T"])
        return self._stage_decode(ctx, z, defer_random=defer_random, t_max=t_max)

    def _stage_encode(self, batch, noise, lengths=None):
        """scene encoder + interactee VAE encode (cond and CFG-uncond) -> the denoiser's conditioning and the initial noise"""
        image_emb = None
        if "image" in self.condition:
            # the tuple the dataset yields and train_diffusion_forward unpacks (dataset.py:1788-1792, mld.py:889-909);
            # ego_eval's own unpacking drops `length` and then reads it (App. D4)
            if "scene" in self.condition:
                feats_ref, transl, beta, utils_, scene, images, length = batch
                scene_emb = self._encode_scene(scene)
            else:
                feats_ref, transl, beta, utils_, images, length = batch
                scene_emb = None
            image_emb = self.proscene.backbone.op(self.output_images)(images.float())[None]   # mld.py:1083-1086
        elif "scene" in self.condition:
            feats_ref, transl, beta, utils_, scene, length, dict_images = batch
            scene_emb = self._encode_scene(scene)                          # [1,B or 2B,256]
        else:
            feats_ref, transl, beta, utils_, length = batch[:5]
            scene_emb = None
        feats_ref, transl, beta = feats_ref.float(), transl.float(), beta.float()
        dev = feats_ref.device
        if lengths is None:
            lengths = length.long().reshape(-1).tolist()                   # mld.py:1264
        len_dev = length.reshape(-1).to(torch.int32)
        start = time.time()
        f_ref_int = None
        if "interactee" in self.condition:
            f_ref_int = torch.cat([feats_ref[:, :, 1, :], transl[:, 1, :, :]], dim=-1).contiguous()
            B, T, _ = f_ref_int.shape
            text_emb, _ = self.vae.encode(f_ref_int, None, lengths, eps=noise.get("eps_int"), lengths_dev=len_dev)
            if self.do_classifier_free_guidance:
                unc = self._encode_uncond(B, T, f_ref_int.shape[-1], lengths, noise.get("eps_unc"), dev)
                text_emb = torch.cat([unc, text_emb], dim=1)               # mld.py:1290 (UNCOND first)
            tokens = [text_emb, scene_emb, image_emb]                       # mld.py:1297-1313 (token order)
        else:
            tokens = [scene_emb, image_emb]
        tokens = [t for t in tokens if t is not None]
        cond_emb = (tokens[0] if len(tokens) == 1 else torch.cat(tokens, dim=0)) if tokens else None
        if cond_emb is None:
            raise NotImplementedError("MLD (sm_100a): at least one of scene / image / interactee conditioning is required")
        x_T = noise.get("x_T")
        if x_T is None:                                                    # drawn where the reference draws it (mld.py:449-453)
            x_T = torch.randn((feats_ref.shape[0], self.latent_dim[0], self.latent_dim[-1]), device=dev, dtype=torch.float)
        return {"feats_ref": feats_ref, "transl": transl, "beta": beta, "lengths": lengths, "len_dev": len_dev,
                "f_ref_int": f_ref_int, "cond_emb": cond_emb, "x_T": x_T, "start": start}

    def _stage_decode(self, ctx, z, defer_random: bool = False, t_max: Optional[int] = None):
        """VAE decode of the sampled latents [1,B,256], renorm + SMPL of the three bodies -> rs_set (mld.py:1355-1905)"""
        feats_ref, transl, beta, lengths, len_dev, f_ref_int = (ctx[k] for k in ("feats_ref", "transl", "beta", "lengths", "len_dev", "f_ref_int"))
        dev = feats_ref.device
        feats_rst = self.vae.decode(z, lengths, T=t_max, lengths_dev=len_dev)   # [B,max(lengths),nfeats_net]
        self.times.append(time.time() - ctx["start"])                      # mld.py:1367-1368 (no device sync, like the reference)

        min_len = min(feats_ref.shape[1], feats_rst.shape[1])
        idx_ref = 0 if self.estimate == "wearer" else 1
        mean, std = self._stats(dev)
        Bsz = feats_ref.shape[0]
        n_body = 69 if self.name_dataset == "egobody" else 63             # gimo: 21 joints + 2 zeroed hands (mld.py:1659-1665)
        want_all = self.compute_vertices == "all"
        # ground-truth body: cat(feats_ref, transl) -> renorm -> SMPL (mld.py:1451-1488 / 1656-1690)
        f_ref = torch.cat([feats_ref[:, :min_len, idx_ref, :], transl[:, idx_ref, :min_len, :]], dim=-1).contiguous()
        betas_w = beta[:, idx_ref, :min_len, :].contiguous()
        m_ref, v_ref, joints_ref, quat_ref = self._body(f_ref, betas_w, mean, std, n_body, want_all)
        # predicted body, GT betas (mld.py:1490-1534 / 1692-1741)
        fr = feats_rst[:, :min_len]
        if self.name_dataset == "gimo":
            fr = torch.cat([fr[:, :, :66], fr[:, :, -3:]], dim=-1)          # mld.py:1692-1694
        fr_pred = fr
        if not self.pred_global_orient:
            fr = torch.cat([f_ref[:, :, :3], fr[:, :, 3:]], dim=-1)         # mld.py:1501-1505 (same stats on both sides)
        m_rst, v_rst, joints_rst, quat_rst = self._body(fr.contiguous(), betas_w, mean, std, n_body,
                                                        self.compute_vertices in ("rst", "all"))
        if not self.pred_global_orient:
            # only the SMPL input takes the ground-truth orientation; rs_set["m_rst"] stays renorm(feats_rst) (mld.py:1490-1491)
            m_rst[:, :, :3] = fr_pred[:, :, :3].double() * std[:3] + mean[:3]
        rs_set = {
            "m_ref": m_ref, "m_rst": m_rst, "joints_ref": joints_ref, "joints_rst": joints_rst,
            "orientation_quat_rst": quat_rst, "orientation_quat_ref": quat_ref,
            "joints_interactee_gt": None, "lengths": lengths, "list_names": {},
        }
        v_int = None
        if f_ref_int is not None:
            # the interactee is always a 75-d row with a 69-d body pose (mld.py:1542-1570 / 1744-1766)
            _, v_int, joints_int, quat_int = self._body(f_ref_int[:, :min_len].contiguous(), beta[:, 1, :min_len, :].contiguous(),
                                                        mean, std, 69, want_all, want_m=False)
            rs_set["joints_interactee"] = joints_int
            rs_set["root_interactee"] = joints_int[:, :, 0:1, :]
            rs_set["orientation_quat_int"] = quat_int
        elif not defer_random:
            joints_int = torch.rand_like(joints_rst)                        # mld.py:1572-1574 (RNG side effect kept)
            rs_set["joints_interactee"] = joints_int
            rs_set["root_interactee"] = joints_int[:, :, 0:1, :]
            rs_set["orientation_quat_int"] = torch.rand_like(quat_rst)
        last_vertices = {k: (None if v is None else v.view(Bsz, min_len, 6890, 3))
                         for k, v in (("rst", v_rst), ("ref", v_ref), ("int", v_int))}
        if defer_random:       # lane mode: the caller merges these
            rs_set["_last_vertices"] = last_vertices
            rs_set["_last_latent"] = z
        else:
            self.last_vertices = last_vertices
            self.last_latent = z
        return rs_set

    def _body(self, feats, betas, mean, std, n_body, want_v, want_m=True):
        """renorm (float64) + slicing + SMPL forward + aa_to_quat for a [B,T,Dn] normalised feature tensor."""
        B, T, Dn = feats.shape
        if mean.numel() < Dn:
            raise ValueError(f"dataset statistics have {mean.numel()} dims, features have {Dn}")
        m, v, j, q = self.smpl_model.op.forward_feats(feats.reshape(B * T, Dn), mean[:Dn], std[:Dn], n_body,
                                                      betas.reshape(B * T, 10), want_vertices=want_v, want_m=want_m)
        return (None if m is None else m.view(B, T, Dn)), v, j.view(B, T, 24, 3), q

    # ------------------------------------------------------------------------------------------
    def forward(self, batch):
        """The reference's ``forward`` is dead code calling a removed text encoder (App. D3); here tuple
        batches are routed to ``ego_eval`` and the un-padded predicted joints are returned."""
        rs_set = self.ego_eval(batch)
        return [j[:n] for j, n in zip(rs_set["joints_rst"], rs_set["lengths"])]

    def test_step(self, batch, batch_idx):                                  # base.py:44-53
        return self.allsplit_step("test", batch, batch_idx)

    def run_test_batches(self, batches, split: str = "test"):
        """The test loop of ``trainer.test`` (base.py:44-53 per batch) with ``pipeline_depth`` batches in flight:
        batch k's metric update runs while batches k+1.. are on the GPU.  Yields ``joints_rst`` per batch, in order."""
        from collections import deque
        pending = deque()

        def retire():
            rs_set = pending.popleft().result()
            for metric in self.metrics_dict:
                if metric == "MRMetrics":
                    getattr(self, metric).update(rs_set["joints_rst"], rs_set["joints_ref"], rs_set["lengths"])
                    continue
                getattr(self, metric).update(
                    split, rs_set["joints_rst"], rs_set["joints_ref"], rs_set["orientation_quat_rst"],
                    rs_set["orientation_quat_ref"], rs_set["root_interactee"], rs_set["joints_interactee"],
                    rs_set["orientation_quat_int"], rs_set["joints_interactee_gt"], rs_set["lengths"], rs_set["list_names"])
            return rs_set["joints_rst"]

        for batch in batches:
            pending.append(self.ego_eval_async(batch))
            if len(pending) >= max(1, int(self.pipeline_depth)):
                yield retire()
        while pending:
            yield retire()

    def validation_step(self, batch, batch_idx):
        return self.allsplit_step("val", batch, batch_idx)

    def training_step(self, batch, batch_idx):
        raise NotImplementedError("training is out of scope for the B200 inference path")

    def allsplit_step(self, split: str, batch, batch_idx):                  # mld.py:2037-2130
        if split not in ("val", "test"):
            raise NotImplementedError("training is out of scope for the B200 inference path")
        rs_set = self.ego_eval(batch)
        for metric in self.metrics_dict:
            if metric == "EgoMetric":
                getattr(self, metric).update(
                    split, rs_set["joints_rst"], rs_set["joints_ref"], rs_set["orientation_quat_rst"],
                    rs_set["orientation_quat_ref"], rs_set["root_interactee"], rs_set["joints_interactee"],
                    rs_set["orientation_quat_int"], rs_set["joints_interactee_gt"], rs_set["lengths"], rs_set["list_names"])
            elif metric == "MRMetrics":                                         # mld.py:2116-2119
                getattr(self, metric).update(rs_set["joints_rst"], rs_set["joints_ref"], rs_set["lengths"])
            else:
                raise TypeError(f"Not support this metric {metric}")
        if split == "test":
            return rs_set["joints_rst"]
        return None

    def allsplit_epoch_end(self, split: str, outputs=None):                 # base.py:58-99
        dico = {}
        for metric in self.metrics_dict:
            metrics_dict = getattr(self, metric).compute(sanity_flag=False)
            getattr(self, metric).reset()
            dico.update({f"Metrics/{m}": float(v) for m, v in metrics_dict.items()})
        return dico

    def on_test_epoch_end(self, outputs=None):
        self.cfg.TEST.REP_I = self.cfg.TEST.get("REP_I", 0) + 1
        return self.allsplit_epoch_end("test", outputs)
