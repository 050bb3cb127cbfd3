"""Parity metric of the hot path (reference model/metric.py:5-34)."""
import torch


def sisnr(s_hat: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """Mean scale-invariant SNR in dB of estimate ``s_hat`` against ``s``; inputs [B,1,T] or [B,T]."""
    s_hat = s_hat.reshape(s_hat.shape[0], 1, -1)
    s = s.reshape(s.shape[0], 1, -1)
    s_hat = s_hat - s_hat.mean(dim=-1, keepdim=True)
    s = s - s.mean(dim=-1, keepdim=True)
    target = (s_hat * s).sum(-1, keepdim=True) * s / (s * s).sum(-1, keepdim=True)
    resid = s_hat - target
    ratio = (target * target).sum(-1, keepdim=True) / (resid * resid).sum(-1, keepdim=True)
    return torch.squeeze(torch.mean(10 * torch.log10(ratio)))
