"""``SDDM`` — host mirror of reference model/model.py:7-124 (the arch wrapper whose ``infer`` IS the hot loop).

``infer(condition)`` keeps the reference signature and semantics (x_T initialisation per ``p_transition``, T reverse
steps, clamp every step, returns x_0 of shape [B,1,L]) but enqueues the whole loop through one C-ABI call
(``sddm_sample``).  Two extensions, both keyword-only and off by default:
  * ``noises=[T,B,1,L]`` injects the Gaussian draws (verification mode; order as the reference consumes them);
  * ``seed=...`` fixes the in-kernel Philox stream (default: drawn from torch's CPU generator).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
from torch import nn

from .diffusion import GaussianDiffusion
from .diffwave import DiffWave
from .wavegrad import WaveGrad
from .unet_modified2 import UNetModified2


class BaseModel(nn.Module):
    """reference base/base_model.py: abstract forward + a __str__ that reports the trainable parameter count."""

    def forward(self, *inputs):
        raise NotImplementedError

    def __str__(self):
        n = sum(int(np.prod(p.size())) for p in self.parameters() if p.requires_grad)
        return super().__str__() + "\nTrainable parameters: {}".format(n)


class SDDM(BaseModel):
    def __init__(self, diffusion: GaussianDiffusion, noise_estimate_model: nn.Module, noise_condition="sqrt_alpha_bar",
                 p_transition="original", q_transition="original"):
        super().__init__()
        self.diffusion = diffusion
        self.noise_estimate_model = noise_estimate_model
        self.num_timesteps = self.diffusion.num_timesteps
        self.noise_condition = noise_condition
        self.p_transition = p_transition
        self.q_transition = q_transition
        if noise_condition not in ("sqrt_alpha_bar", "time_step"):
            raise NotImplementedError
        if p_transition not in ("original", "supportive", "sr3", "conditional", "condition_in"):
            raise NotImplementedError
        if q_transition not in ("original", "conditional"):
            raise NotImplementedError

    # train step, forward half (reference :29-48): forward-diffusion draw + eps_hat; returns (predicted, noise) for the loss.
    # Inference-mode only: there are no backward kernels, so no gradients flow.
    @torch.no_grad()
    def forward(self, target, condition, *, seed: Optional[int] = None):
        if not target.is_cuda:
            raise RuntimeError("SDDM.forward (sddm_b200) needs CUDA tensors: there is no CPU fallback")
        if self.q_transition == "original":
            x_t, noise_level, t, noise = self.diffusion.q_stochastic(target, None, seed=seed, return_noise=True)
            level = noise_level if self.noise_condition == "sqrt_alpha_bar" else t
            predicted = self.noise_estimate_model(condition, x_t, level)
        else:
            x_t, noise, noise_level = self.diffusion.q_stochastic_conditional(target, condition, None, seed=seed)
            predicted = self.noise_estimate_model(condition, x_t, noise_level)
        return predicted, noise

    def _fused_ok(self) -> bool:
        return isinstance(self.noise_estimate_model, UNetModified2) and self.noise_condition == "sqrt_alpha_bar"

    @torch.no_grad()
    def infer(self, condition, continuous=False, *, noises: Optional[torch.Tensor] = None, seed: Optional[int] = None,
              row0: int = 0, return_trace: bool = False):
        if not condition.is_cuda:
            raise RuntimeError("SDDM.infer (sddm_b200) needs CUDA tensors: there is no CPU fallback")
        T = self.num_timesteps
        if seed is None and noises is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        seed = 0 if seed is None else int(seed)
        if continuous:
            assert condition.shape[0] == 1, "Batch size must be 1 to do continuous sampling"
        if self._fused_ok():
            plan = self.noise_estimate_model.get_plan(self.diffusion)
            want_trace = continuous or return_trace
            res = plan.sample(condition, self.p_transition, noises=noises, seed=seed, row0=row0, trace=want_trace)
            if not want_trace:
                return res.reshape(condition.shape)
            out, eps_tr, x_tr = res
            if return_trace:
                return out.reshape(condition.shape), eps_tr, x_tr
            every = 1 | (T // 100)                                   # reference model.py:73
            samples = [condition]
            samples += [x_tr[T - t].reshape(condition.shape) for t in range(T, 0, -1) if t % every == 0]
            return samples
        return self._infer_stepwise(condition, continuous, noises, seed)

    def _infer_stepwise(self, condition, continuous, noises, seed):
        """Generic loop (time_step conditioning / foreign denoisers): one eps call + one update kernel per step."""
        d, T, B = self.diffusion, self.num_timesteps, condition.shape[0]
        z = (lambda k: None) if noises is None else (lambda k: noises[k])
        if self.p_transition == "conditional":
            x_t = d.get_x_T_conditional(condition, noise=z(0), seed=seed)
        elif self.p_transition == "condition_in":
            x_t = d.get_x_T(condition, noise=z(0), seed=seed)
        elif self.p_transition == "supportive":
            x_t = condition
        else:
            x_t = d._x_T("original", condition, z(0), seed)           # pure-noise start (reference model.py:68-70)
        shape = [B] + [1] * (condition.ndim - 1)
        every = 1 | (T // 100)
        samples = [condition]
        for t in range(T, 0, -1):
            if self.noise_condition == "sqrt_alpha_bar":
                level = d.get_noise_level(t) * torch.ones(shape, device=condition.device)
            else:
                level = t * torch.ones(shape, device=condition.device)
            predicted = self.noise_estimate_model(condition, x_t, level)
            zt = z(T + 1 - t) if t > 1 else None
            if self.p_transition in ("original", "condition_in"):
                x_t = d.p_transition(x_t, t, predicted, noise=zt, seed=seed + t)
            elif self.p_transition == "sr3":
                x_t = d.p_transition_sr3(x_t, t, predicted, noise=zt, seed=seed + t)
            elif self.p_transition == "supportive":
                x_t = d.p_transition_supportive(x_t, t, predicted, condition, noise=zt, seed=seed + t)
            else:
                x_t = d.p_transition_conditional(x_t, t, predicted, condition, noise=zt, seed=seed + t)
            if continuous and t % every == 0:
                samples.append(x_t)
        return samples if continuous else x_t


class SDDM_spectrogram(SDDM):
    """reference model/model.py:206-257: spectrogram-conditioned sampling (config_diffwave.json / config_wavegrad.json):
    pure-noise start of length ``hop_samples * frames``, then T x (eps_hat, p_transition 'original').  With WaveGrad the
    shipped reference wrapper raises (it hands [B,1,T] to WaveGrad.forward, which needs [B,T]): here the loop
    runs, with that one squeeze."""

    def __init__(self, diffusion: GaussianDiffusion, noise_estimate_model: nn.Module, hop_samples: int,
                 noise_condition="sqrt_alpha_bar"):
        super().__init__(diffusion, noise_estimate_model, noise_condition)
        self.hop_samples = hop_samples

    @torch.no_grad()
    def infer(self, condition, continuous=False, *, noises: Optional[torch.Tensor] = None, seed: Optional[int] = None,
              row0: int = 0, return_trace: bool = False):
        if not condition.is_cuda:
            raise RuntimeError("SDDM_spectrogram.infer (sddm_b200) needs CUDA tensors: there is no CPU fallback")
        net = self.noise_estimate_model
        if not isinstance(net, (DiffWave, WaveGrad)):
            raise NotImplementedError("SDDM_spectrogram (sddm_b200) drives the DiffWave and WaveGrad denoisers")
        want_hop = 256 if isinstance(net, DiffWave) else 300
        if self.hop_samples != want_hop:
            raise ValueError("%s upsamples a spectrogram frame to %d samples: hop_samples must be %d, got %d"
                             % (type(net).__name__, want_hop, want_hop, self.hop_samples))
        if seed is None and noises is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        seed = 0 if seed is None else int(seed)
        plan = net.get_plan(self.diffusion, self.noise_condition)
        if not continuous:
            return plan.sample(condition, noises=noises, seed=seed, row0=row0, trace=return_trace)
        # continuous=True (reference :230-244): the intermediate x_t every `sample_inter` steps, one eps call + one update kernel per step
        assert condition.shape[0] == 1, "Batch size must be 1 to do continuous sampling"
        d, T, B = self.diffusion, self.num_timesteps, condition.shape[0]
        every = 1 | (T // 100)
        Ls = self.hop_samples * condition.shape[-1]
        z = (lambda k: None) if noises is None else (lambda k: noises[k].reshape(B, 1, Ls))
        x_t = d._x_T("original", torch.empty(B, 1, Ls, device=condition.device), z(0), seed)      # pure-noise start (:216)
        samples = [condition]
        if isinstance(net, DiffWave):
            plan.condition(condition)          # step independent: evaluated once for this spectrogram, reused by every step below
        for t in range(T, 0, -1):
            predicted = (plan.eps(condition, x_t, t=t, reuse_condition=True) if isinstance(net, DiffWave)
                         else plan.eps(condition, x_t, t=t)).reshape(x_t.shape)
            x_t = d.p_transition(x_t, t, predicted, noise=z(T + 1 - t) if t > 1 else None, seed=seed + t)
            if t % every == 0:
                samples.append(x_t)
        return samples
