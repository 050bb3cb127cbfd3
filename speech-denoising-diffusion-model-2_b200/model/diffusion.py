"""``GaussianDiffusion`` — host mirror of reference model/diffusion.py:49-326.

Same constructor kwargs, the same 14 registered buffers (so reference checkpoints load), the same method names.
The schedule tables are built once on the CPU in fp32 with the reference's op order (bit-identical to the
reference built with ``device='cpu'``) and then moved to ``device``; every per-element update runs in the
hand-written CUDA kernels behind ``sddm_x_T_raw`` / ``sddm_p_step_raw`` (csrc/kernels_misc.cu).  CPU tensors are
rejected: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from .. import _lib

_BUFFERS = ("betas", "alphas", "alpha_bar", "sqrt_alpha_bar", "predicted_noise_coeff", "sigma", "supportive_gamma",
            "supportive_sigma_hat", "m", "sqrt_delta", "c_xt", "c_yt", "c_epst", "sqrt_delta_estimated")


def build_schedule(schedule: str, n_timestep: int, linear_start: float, linear_end: float):
    """fp32 CPU tables of length T+1 (reference diffusion.py:65-161)."""
    T, f32 = n_timestep, torch.float32
    betas = torch.zeros(T + 1, dtype=f32)
    if schedule == "linear":
        betas[1:] = torch.linspace(linear_start, linear_end, T, dtype=f32)
        alphas = 1 - betas
        alpha_bar = torch.cumprod(alphas, dim=0)
    elif schedule == "quad":
        betas[1:] = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, T, dtype=f32) ** 2
        alphas = 1 - betas
        alpha_bar = torch.cumprod(alphas, dim=0)
    elif schedule == "cosine":
        s = 0.008
        grid = torch.arange(T + 1, dtype=f32) / T + s
        f = torch.cos(grid / (1 + s) * (torch.pi / 2)).pow(2)
        alpha_bar = f / f[0]
        betas[1:] = 1 - alpha_bar[1:] / alpha_bar[:-1]
        betas = betas.clamp(max=0.999)
        alphas = 1 - betas
    else:
        raise NotImplementedError
    tab = dict(betas=betas, alphas=alphas, alpha_bar=alpha_bar, sqrt_alpha_bar=torch.sqrt(alpha_bar))
    one_m_ab = 1.0 - alpha_bar
    sigma = torch.zeros_like(betas)
    sigma[1:] = (one_m_ab[:-1] / one_m_ab[1:] * betas[1:]) ** 0.5
    coeff = torch.zeros_like(betas)
    coeff[1:] = betas[1:] / torch.sqrt(1 - alpha_bar[1:])
    gamma = torch.zeros_like(betas)
    gamma[1] = 0.2
    gamma[2:] = sigma[2:]
    sigma_hat = torch.zeros_like(betas)
    sigma_hat[1:] = sigma[1:] - gamma[1:] / torch.sqrt(alphas[1:])
    tab.update(predicted_noise_coeff=coeff, sigma=sigma, supportive_gamma=gamma, supportive_sigma_hat=sigma_hat)
    # conditional diffusion (Lu et al.) coefficients
    sab = tab["sqrt_alpha_bar"]
    m = torch.sqrt((1 - alpha_bar) / sab)
    delta = (1 - alpha_bar) - m ** 2 * alpha_bar
    r = (1 - m[1:]) / (1 - m[:-1])
    ad = alphas[1:] * delta[:-1]
    d_step = delta[1:] - r ** 2 * ad
    root_a = torch.sqrt(alphas[1:])
    c_xt, c_yt, c_epst, d_est = (torch.zeros_like(betas) for _ in range(4))
    c_xt[1:] = r * delta[:-1] / delta[1:] * root_a + (1 - m[:-1]) * (d_step / delta[1:]) * (1 / root_a)
    c_yt[1:] = (m[:-1] * delta[1:] - m[1:] * r * ad) * sab[:-1] / delta[1:]
    c_epst[1:] = (1 - m[:-1]) * d_step / delta[1:] * torch.sqrt(1 - alpha_bar[1:]) / root_a
    d_est[1:] = d_step * delta[:-1] / delta[1:]
    tab.update(m=m, sqrt_delta=torch.sqrt(delta), c_xt=c_xt, c_yt=c_yt, c_epst=c_epst,
               sqrt_delta_estimated=torch.sqrt(d_est))
    return tab


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sddm_b200 runs on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _prep(t, like=None):
    if t is None:
        return None
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class GaussianDiffusion(nn.Module):
    def __init__(self, schedule="linear", n_timestep=1000, linear_start=1e-4, linear_end=2e-2, device="cuda"):
        super().__init__()
        self.num_timesteps = n_timestep
        self.device = device
        self.schedule = schedule
        tab = build_schedule(schedule, n_timestep, linear_start, linear_end)
        for k in _BUFFERS:
            self.register_buffer(k, tab[k].to(device))
        self._host = None

    # -- host-side scalar access (no device sync inside the loop) -------------------------------------
    def host_tables(self):
        if self._host is None:
            self._host = {k: getattr(self, k).detach().cpu().numpy().astype(np.float32) for k in _BUFFERS}
        return self._host

    def tables_key(self) -> str:
        """Content hash of the schedule tables: the plan caches of the denoisers key on it (an id() can be reused by a rebuilt dict
        after a checkpoint load)."""
        import hashlib
        h = hashlib.sha1()
        for k in _BUFFERS:
            h.update(self.host_tables()[k].tobytes())
        return h.hexdigest()

    def _load_from_state_dict(self, *a, **k):
        self._host = None
        return super()._load_from_state_dict(*a, **k)

    def step_scalars(self, t: int, variant: str):
        """k8 of sddm_p_step_raw for step t, each scalar rounded exactly as the reference's eager ops do."""
        h, f = self.host_tables(), np.float32
        if variant == "sr3":
            std = np.sqrt(h["betas"][t])
        elif variant == "supportive":
            std = max(f(0), h["supportive_sigma_hat"][t])
        elif variant == "conditional":
            std = h["sqrt_delta_estimated"][t]
        else:
            std = h["sigma"][t]
        g = h["supportive_gamma"][t]
        vals = [h["predicted_noise_coeff"][t], np.sqrt(h["alphas"][t]), std, g, f(1) - g, h["c_xt"][t], h["c_yt"][t],
                h["c_epst"][t]]
        return (C.c_float * 8)(*[float(v) for v in vals])

    # -- reference API ---------------------------------------------------------------------------------
    def get_noise_level(self, t):
        return self.sqrt_alpha_bar[t]

    def _x_T(self, variant, condition, noise, seed):
        _need_cuda(condition, noise)
        cond = _prep(condition)
        z = _prep(noise)
        h, T = self.host_tables(), self.num_timesteps
        a = h["sqrt_alpha_bar"][T]
        b = h["sqrt_delta"][T] if variant == "conditional" else np.sqrt(np.float32(1) - a * a)
        out = torch.empty_like(cond)
        B, L = cond.shape[0], cond.numel() // cond.shape[0]
        with torch.cuda.device(cond.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_x_T_raw(_lib.VARIANTS[variant], float(a), float(b), _ptr(cond), _ptr(z),
                                              _seed(seed), 0, _ptr(out), B, L, C.c_void_p(st)))
        return out

    def get_x_T(self, condition, noise=None, seed=None):
        """x_T = sqrt_ab[T]*y + sqrt(1-sqrt_ab[T]^2)*z; z injected or Philox in-kernel (reference :281-300)."""
        return self._x_T("condition_in", condition, noise, seed)

    def get_x_T_conditional(self, condition, noise=None, seed=None):
        return self._x_T("conditional", condition, noise, seed)

    def _step(self, variant, x_t, t, predicted, condition=None, noise=None, seed=None):
        _need_cuda(x_t, predicted, condition, noise)
        x = _prep(x_t).clone()
        eps, cond, z = _prep(predicted), _prep(condition), _prep(noise)
        B, L = x.shape[0], x.numel() // x.shape[0]
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_p_step_raw(_lib.VARIANTS[variant], self.step_scalars(int(t), variant), _ptr(x),
                                                 _ptr(eps), _ptr(cond), _ptr(z), _seed(seed), 0, int(t),
                                                 self.num_timesteps, B, L, C.c_void_p(st)))
        return x

    @torch.no_grad()
    def p_transition(self, x_t, t, predicted, noise=None, seed=None):
        return self._step("original", x_t, t, predicted, None, noise, seed)

    @torch.no_grad()
    def p_transition_sr3(self, x_t, t, predicted, noise=None, seed=None):
        return self._step("sr3", x_t, t, predicted, None, noise, seed)

    @torch.no_grad()
    def p_transition_supportive(self, x_t, t, predicted_noise, condition, noise=None, seed=None):
        return self._step("supportive", x_t, t, predicted_noise, condition, noise, seed)

    @torch.no_grad()
    def p_transition_conditional(self, x_t, t, predicted_noise, condition, noise=None, seed=None):
        return self._step("conditional", x_t, t, predicted_noise, condition, noise, seed)

    # -- forward diffusion (the draw of the training step SDDM.forward) -------------------------------
    def _q(self, mode, coef, x_0, y, noise, seed, row0):
        _need_cuda(x_0, y, noise)
        x0, yy, z = _prep(x_0), _prep(y), _prep(noise)
        B, L = x0.shape[0], x0.numel() // x0.shape[0]
        x_t = torch.empty_like(x0)
        comb = torch.empty_like(x0) if mode == 1 else None
        z_out = torch.empty_like(x0) if z is None else None
        coef = coef.to(x0.device, torch.float32).contiguous()
        with torch.cuda.device(x0.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_q_sample_raw(mode, _ptr(coef), _ptr(x0), _ptr(yy), _ptr(z), _seed(seed), int(row0), _ptr(x_t), _ptr(comb),
                                                    _ptr(z_out), B, L, C.c_void_p(st)))
        return x_t, comb, (z if z is not None else z_out)

    @torch.no_grad()
    def q_stochastic(self, x_0, noise, t_is_integer=False, *, t=None, random_step=None, seed=None, row0=0, return_noise=False):
        """reference :225-251.  Returns (x_t, sqrt_alpha_bar_sample [B,1,..], t + random_step [B,1,..]).  The per-row scalars
        are drawn / gathered with the reference's own torch ops on B-element tensors; the [B,1,T] work is one CUDA kernel.
        Extensions (keyword-only): fixed ``t`` / ``random_step`` draws (tests), ``noise=None`` => in-kernel Philox."""
        b = x_0.shape[0]
        shape = [b] + [1] * (x_0.ndim - 1)
        dev = x_0.device
        if t is None:
            t = torch.randint(1, self.num_timesteps + 1, [b], device=dev)
        t = t.to(dev).reshape(b)
        if t_is_integer:
            sample = self.sqrt_alpha_bar[t]
            random_step = 0
        else:
            l_a, l_b = self.sqrt_alpha_bar[t - 1], self.sqrt_alpha_bar[t]
            if random_step is None:
                random_step = torch.rand(b, device=dev)
            random_step = random_step.to(dev).reshape(b)
            sample = l_a + random_step * (l_b - l_a)
        coef = torch.stack([sample, torch.sqrt(1.0 - torch.square(sample)), torch.zeros_like(sample), torch.zeros_like(sample)], dim=1)
        x_t, _, z = self._q(0, coef, x_0, None, noise, seed, row0)
        out = (x_t.reshape(x_0.shape), sample.view(shape), (t + random_step).view(shape))
        return out + (z.reshape(x_0.shape),) if return_noise else out

    @torch.no_grad()
    def q_stochastic_conditional(self, x_0, y, noise, *, t=None, seed=None, row0=0, return_noise=False):
        """reference :253-279.  Returns (x_t, combined_noise, sqrt_alpha_bar[t] [B,1,..])."""
        b = x_0.shape[0]
        shape = [b] + [1] * (x_0.ndim - 1)
        dev = x_0.device
        if t is None:
            t = torch.randint(1, self.num_timesteps + 1, tuple(shape), device=dev)
        t = t.to(dev).reshape(b)
        sab = self.sqrt_alpha_bar[t]
        coef = torch.stack([sab, self.m[t] * sab, self.sqrt_delta[t], 1.0 / torch.sqrt(1.0 - self.alpha_bar[t])], dim=1)
        x_t, comb, z = self._q(1, coef, x_0, y, noise, seed, row0)
        out = (x_t.reshape(x_0.shape), comb.reshape(x_0.shape), sab.view(shape))
        return out + (z.reshape(x_0.shape),) if return_noise else out


class VariableGaussianDiffusion(nn.Module):
    """Mirror of the reference's SNR-adaptive diffusion (model/diffusion.py:329-446): every frame n of a row carries its own linear
    beta schedule whose end value follows from the estimated SNR of that frame.  Same constructor arguments and method names; tensors
    are frames ``[B, 1, N, L]`` and ``snr_estimate [B, N]`` (dB).  The reference rebuilds the whole ``[B, 1, N, T + 1]`` schedule with
    numpy on the host inside every call; here each kernel recomputes the terms it needs per frame (csrc/var_diffusion.cu).
    Extensions (keyword-only): injected ``noise`` for tests, ``seed`` / ``row0`` for the in-kernel Philox stream.
    The two networks of that variant (SNREstimator, UNetModified2_withVariableNoiseLevel) are not built - SURVEY section 8f row 4."""

    def __init__(self, n_timestep=100, snr_estimate_scale=100, device="cuda"):
        super().__init__()
        self.num_timesteps = int(n_timestep)
        self.snr_estimate_scale = float(snr_estimate_scale)
        self.device = device
        self.linear_start = 1e-6

    def _frames(self, x, snr):
        x = _prep(x)
        snr = _prep(snr)
        _need_cuda(x, snr)
        if x.ndim != 4 or x.shape[1] != 1:
            raise ValueError("frames must be [B, 1, N, L], got %s" % (tuple(x.shape),))
        B, _, N, L = x.shape
        snr = snr.reshape(B, -1)
        if snr.shape[1] != N:
            raise ValueError("snr_estimate must hold one value per frame: [%d, %d], got %s" % (B, N, tuple(snr.shape)))
        return x, snr.contiguous(), B, N, L

    @torch.no_grad()
    def get_beta_schedule(self, snr_estimate):
        """reference :345-359 -> (betas, alpha_bar), both [B, 1, N, T + 1]"""
        snr = _prep(snr_estimate)
        _need_cuda(snr)
        B, N = snr.shape
        T = self.num_timesteps
        betas = torch.empty(B, N, T + 1, device=snr.device)
        ab = torch.empty_like(betas)
        with torch.cuda.device(snr.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_var_schedule(_ptr(snr), B, N, T, self.snr_estimate_scale, _ptr(betas), _ptr(ab), C.c_void_p(st)))
        return betas.unsqueeze(1), ab.unsqueeze(1)

    @torch.no_grad()
    def get_noise_level(self, t, snr_estimate):
        """reference :438-444 -> sqrt(alpha_bar_t) per frame, [B, 1, N, 1]"""
        snr = _prep(snr_estimate)
        _need_cuda(snr)
        B, N = snr.shape
        out = torch.empty(B, N, device=snr.device)
        with torch.cuda.device(snr.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_var_noise_level(_ptr(snr), B, N, self.num_timesteps, self.snr_estimate_scale, int(t), _ptr(out), C.c_void_p(st)))
        return out.reshape(B, 1, N, 1)

    def _mix(self, x, snr_estimate, t, noise, seed, row0):
        x, snr, B, N, L = self._frames(x, snr_estimate)
        z = _prep(noise)
        _need_cuda(z)
        out = torch.empty_like(x)
        level = torch.empty(B, N, device=x.device)
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_var_mix(_ptr(x), _ptr(snr), _ptr(z), _seed(seed), int(row0), B, N, L, self.num_timesteps,
                                               self.snr_estimate_scale, int(t), _ptr(out), _ptr(level), C.c_void_p(st)))
        return out, level.reshape(B, 1, N, 1)

    @torch.no_grad()
    def get_x_T(self, condition, snr_estimate, *, noise=None, seed=None, row0=0):
        """reference :417-435"""
        return self._mix(condition, snr_estimate, self.num_timesteps, noise, seed, row0)[0]

    @torch.no_grad()
    def q_stochastic(self, x_0, noise, snr_estimate, t_is_integer=True, *, t=None, seed=None, row0=0):
        """reference :394-415 -> (x_t, sqrt_alpha_bar_sample [B, 1, N, 1], t); one t for the whole batch, as there"""
        if not t_is_integer:
            raise NotImplementedError   # as the reference
        if t is None:
            t = torch.randint(1, self.num_timesteps + 1, [1])
        t = torch.as_tensor(t).reshape(1)
        x_t, level = self._mix(x_0, snr_estimate, int(t.item()), noise, seed, row0)
        return x_t, level, t.to(x_t.device)

    @torch.no_grad()
    def p_transition(self, x_t, t, snr_estimate, predicted, *, noise=None, seed=None, row0=0):
        """reference :373-391"""
        x, snr, B, N, L = self._frames(x_t, snr_estimate)
        eps = _prep(predicted)
        z = _prep(noise)
        _need_cuda(eps, z)
        if eps.shape != x.shape:
            raise ValueError("predicted must have the shape of x_t")
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sddm_var_posterior(_ptr(x), _ptr(eps), _ptr(snr), _ptr(z), _seed(seed), int(row0), B, N, L, self.num_timesteps,
                                                     self.snr_estimate_scale, int(t), _ptr(out), C.c_void_p(st)))
        return out


def _seed(seed):
    if seed is None:   # draw from torch's CPU generator so torch.manual_seed governs reproducibility
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return C.c_uint64(int(seed) & (2 ** 64 - 1))
