"""Duck type of ``diffusers.DDIMScheduler`` for the reference's scheduler.yaml
(``configs/modules/scheduler.yaml:1-14``; call sites ``mld/models/modeltype/mld.py:456-464,495-497``).

The schedule tables live on the host in fp32 and are built with the same tensor ops diffusers uses
(``linspace(sqrt(b0), sqrt(b1)) ** 2`` -> ``cumprod(1 - beta)``), so the coefficients the CUDA kernels
receive are bit-identical to what the reference's scheduler would compute on the CPU.
``step`` runs the fused elementwise CUDA kernel ``seeme_ddim_step``; the sampler
(``seeme_sampler_run``) consumes the per-step coefficient table from ``step_coefficients``.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List

import torch


class DDIMScheduler:
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", clip_sample: bool = True, set_alpha_to_one: bool = True,
                 steps_offset: int = 0, prediction_type: str = "epsilon", **kwargs):
        if beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__.__name__}")
        if clip_sample:
            raise NotImplementedError("clip_sample=True is not on the SEE-ME path (scheduler.yaml:10)")
        if prediction_type != "epsilon":
            raise NotImplementedError("only epsilon prediction is supported (PREDICT_EPSILON: True, base.yaml:27)")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, steps_offset=steps_offset,
                                      beta_start=beta_start, beta_end=beta_end, beta_schedule=beta_schedule,
                                      clip_sample=clip_sample, set_alpha_to_one=set_alpha_to_one,
                                      prediction_type=prediction_type)
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.long)

    def set_timesteps(self, num_inference_steps: int, device=None):
        n_train = self.config.num_train_timesteps
        if num_inference_steps > n_train:
            raise ValueError(f"num_inference_steps {num_inference_steps} > num_train_timesteps {n_train}")
        self.num_inference_steps = num_inference_steps
        ratio = n_train // num_inference_steps          # "leading" spacing
        ts = (torch.arange(0, num_inference_steps) * ratio).flip(0).long() + self.config.steps_offset
        self.timesteps = ts if device is None else ts.to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _coef(self, t: int):
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return torch.stack([(1 - a_t) ** 0.5, a_t ** 0.5, a_p ** 0.5, (1 - a_p) ** 0.5])

    def step_coefficients(self) -> torch.Tensor:
        """fp32 [n,4] = {sqrt(1-abar_t), sqrt(abar_t), sqrt(abar_prev), sqrt(1-abar_prev)} per inference step
        (eta = 0: x0 = (x - c0 eps)/c1; x_prev = c2 x0 + c3 eps)."""
        return torch.stack([self._coef(int(t)) for t in self.timesteps]).float()

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0, **kwargs):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if eta != 0.0:
            raise NotImplementedError("eta != 0 is not on the SEE-ME path (scheduler.yaml:4)")
        from . import ops
        c = self._coef(int(timestep)).tolist()
        prev = ops.ddim_step(model_output, sample, c)
        return SimpleNamespace(prev_sample=prev.view_as(sample))

    def add_noise(self, original_samples, noise, timesteps):
        raise NotImplementedError("add_noise is training-only (mld.py:604-606) and out of scope")

    def __len__(self):
        return self.config.num_train_timesteps


class DDPMScheduler:
    """``noise_scheduler`` is constructed by MLD.__init__ (mld.py:287) but only used in training
    (``_diffusion_process``); kept as an inert placeholder so reference YAMLs instantiate."""

    def __init__(self, **kwargs):
        self.config = SimpleNamespace(**kwargs)

    def add_noise(self, *a, **k):
        raise NotImplementedError("DDPM add_noise is training-only and out of scope")
