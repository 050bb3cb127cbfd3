"""CPU oracle (TEST INFRASTRUCTURE ONLY) of the SNR-adaptive diffusion of SURVEY section 8f row 4: a plain restatement of the
reference's VariableGaussianDiffusion (model/diffusion.py:329-446) - the per-frame schedule, start point, noise level, posterior
step and forward draw.  Every frame n of every row b carries its own linear beta schedule, whose end value follows from the
estimated SNR of that frame.  Pinned against outputs of the reference itself (tests/golden/vardiff.npz, written by
tests/golden/make_golden_vardiff.py in the build container).  The two networks of that variant (SNREstimator,
UNetModified2_withVariableNoiseLevel) are not built; this file and csrc/var_diffusion.cu cover the diffusion half only."""
import numpy as np
import torch

LINEAR_START = 1e-6      # diffusion.py:342 (a one-element tuple there; it multiplies a ones vector, so the value is what counts)


def beta_schedule(snr: torch.Tensor, T: int, scale: float):
    """snr [B, N] (dB) -> betas, alpha_bar [B, 1, N, T + 1]        (get_beta_schedule, diffusion.py:345-359).
    beta_end = (10^(-snr / 20) / scale)^2 in fp32; betas[1..T] = numpy linspace(start, end, T) evaluated in fp64 and rounded to fp32
    (the start vector is fp64, so numpy promotes); betas[0] = 0; alpha_bar = running product of the fp32 values 1 - beta."""
    snr = snr.float()
    ends = (torch.pow(torch.tensor(10.0), snr / -20) / scale) ** 2                     # fp32, as the reference's tensor expression
    start = np.float64(LINEAR_START)
    stop = ends.numpy().astype(np.float64)                                             # [B, N]
    step = (stop - start) / (T - 1)
    i = np.arange(T, dtype=np.float64).reshape(T, 1, 1)
    lin = i * step + start                                                             # numpy.linspace: arange * step + start ...
    lin[-1] = stop                                                                     # ... with the end point written exactly
    betas = np.zeros(snr.shape + (T + 1,), dtype=np.float32)
    betas[..., 1:] = np.moveaxis(lin, 0, -1).astype(np.float32)
    alphas = np.float32(1.0) - betas
    # torch.cumprod on the CPU (where the goldens were made) keeps the running product in fp64 and rounds every output to fp32;
    # on CUDA the reference's scan runs in fp32 in a tree order - both are within 1e-6 of each other, the goldens pin the former
    alpha_bar = np.cumprod(alphas.astype(np.float64), axis=-1).astype(np.float32)
    return torch.from_numpy(betas).unsqueeze(1), torch.from_numpy(alpha_bar).unsqueeze(1)


def noise_level(snr: torch.Tensor, t: int, T: int, scale: float) -> torch.Tensor:
    """sqrt(alpha_bar_t) per frame, [B, 1, N, 1]                   (get_noise_level, diffusion.py:438-444)"""
    _, ab = beta_schedule(snr, T, scale)
    return torch.sqrt(ab[..., [t]])


def x_T(cond: torch.Tensor, snr: torch.Tensor, z: torch.Tensor, T: int, scale: float) -> torch.Tensor:
    """cond, z [B, 1, N, L] -> sqrt(ab_T) cond + sqrt(1 - sqrt(ab_T)^2) z         (get_x_T, diffusion.py:417-435)"""
    s = noise_level(snr, T, T, scale)
    return s * cond + torch.sqrt(1.0 - torch.square(s)) * z


def p_transition(x_t: torch.Tensor, t: int, snr: torch.Tensor, predicted: torch.Tensor, z: torch.Tensor, T: int, scale: float) -> torch.Tensor:
    """One reverse step with the per-frame coefficients, clamped to [-1, 1]      (p_transition, diffusion.py:373-391)"""
    betas, ab = beta_schedule(snr, T, scale)
    b_t, ab_t = betas[..., [t]], ab[..., [t]]
    x = (x_t - b_t / torch.sqrt(1 - ab_t) * predicted) / (1 - b_t) ** 0.5
    if t > 1:
        sigma = ((1.0 - ab[..., [t - 1]]) / (1.0 - ab_t) * b_t) ** 0.5
        x = x + sigma * z
    return x.clamp(-1.0, 1.0)


def q_sample(x_0: torch.Tensor, noise: torch.Tensor, snr: torch.Tensor, t: int, T: int, scale: float):
    """forward draw at an integer step t: (x_t, sqrt(alpha_bar_t))               (q_stochastic, diffusion.py:394-415)"""
    s = noise_level(snr, t, T, scale)
    return s * x_0 + torch.sqrt(1.0 - torch.square(s)) * noise, s
