"""CPU oracle for the reverse-diffusion speech-enhancement hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker / the CPU
baseline — never as the thing shipped.

It is a *functional restatement* (plain fp32 torch ops on CPU, driven directly by a
reference-layout ``state_dict``) of the reference algorithm:

* noise schedule ............ /root/reference/model/diffusion.py:65-161
* x_T initialisation ........ /root/reference/model/diffusion.py:281-300
* posterior updates ......... /root/reference/model/diffusion.py:164-222
* sampling loop ............. /root/reference/model/model.py:50-124
* forward diffusion ......... /root/reference/model/diffusion.py:225-279 (q_stochastic*, the draw of the training step)
* UNetModified2 denoiser .... /root/reference/model/UNetModified2.py:5-269
* SI-SNR .................... /root/reference/model/metric.py:5-34

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real reference
modules from /root/reference (possible only in the build container) and commits
golden vectors under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
restatement against them (and against the reference's only known-answer vectors: the
framing / overlap-add toy in model/tstnn.py:302-308 and the schedule constants).

The arithmetic below the oracle is PyTorch ATen (torch==2.11.0, the version the
reference itself would run on here; the reference pins nothing).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

DIFFUSION_BUFFERS = (
    "betas", "alphas", "alpha_bar", "sqrt_alpha_bar", "predicted_noise_coeff", "sigma",
    "supportive_gamma", "supportive_sigma_hat", "m", "sqrt_delta", "c_xt", "c_yt", "c_epst",
    "sqrt_delta_estimated",
)

UNET_DEFAULT_CFG = dict(
    num_samples=16448, in_channel=2, out_channel=1, inner_channel=32, norm_groups=32,
    channel_mults=(1, 2, 3, 4, 5), res_blocks=1, dropout=0, segment_len=128, segment_stride=64,
)


# --------------------------------------------------------------------------------------
# schedule  (diffusion.py:65-161)
# --------------------------------------------------------------------------------------
def make_schedule(schedule: str = "linear", n_timestep: int = 100, linear_start: float = 1e-6,
                  linear_end: float = 1e-3) -> Dict[str, Tensor]:
    """All 14 fp32 buffers of length T+1 (index 0 is the 'no noise' slot)."""
    T = n_timestep
    f32 = torch.float32
    betas = torch.zeros(T + 1, dtype=f32)
    if schedule == "linear":                                   # diffusion.py:65-69
        betas[1:] = torch.linspace(linear_start, linear_end, T, dtype=f32)
        alphas = 1 - betas
        alpha_bar = torch.cumprod(alphas, dim=0)
    elif schedule == "quad":                                   # diffusion.py:70-73
        betas[1:] = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, T, dtype=f32) ** 2
        alphas = 1 - betas
        alpha_bar = torch.cumprod(alphas, dim=0)
    elif schedule == "cosine":                                 # diffusion.py:74-82
        s = 0.008
        steps = torch.arange(T + 1, dtype=f32) / T + s
        f = torch.cos(steps / (1 + s) * (torch.pi / 2)).pow(2)
        alpha_bar = f / f[0]
        betas[1:] = 1 - alpha_bar[1:] / alpha_bar[:-1]
        betas = betas.clamp(max=0.999)
        alphas = 1 - betas
    else:                                                      # diffusion.py:83-84
        raise NotImplementedError(schedule)
    out = {"betas": betas, "alphas": alphas, "alpha_bar": alpha_bar,
           "sqrt_alpha_bar": torch.sqrt(alpha_bar)}

    # posterior coefficients (diffusion.py:98-107)
    sigma = torch.zeros_like(betas)
    sigma[1:] = ((1.0 - alpha_bar[:-1]) / (1.0 - alpha_bar[1:]) * betas[1:]) ** 0.5
    pnc = torch.zeros_like(betas)
    pnc[1:] = betas[1:] / torch.sqrt(1 - alpha_bar[1:])
    out["predicted_noise_coeff"], out["sigma"] = pnc, sigma

    # 'supportive' variant (diffusion.py:109-118)
    gam = torch.zeros_like(betas)
    gam[1] = 0.2
    gam[2:] = sigma[2:]
    shat = torch.zeros_like(betas)
    shat[1:] = sigma[1:] - gam[1:] / torch.sqrt(alphas[1:])
    out["supportive_gamma"], out["supportive_sigma_hat"] = gam, shat

    # 'conditional' variant (diffusion.py:120-161)
    sab = out["sqrt_alpha_bar"]
    m = torch.sqrt((1 - alpha_bar) / sab)
    delta = (1 - alpha_bar) - m ** 2 * alpha_bar
    ratio = (1 - m[1:]) / (1 - m[:-1])
    a_d = alphas[1:] * delta[:-1]
    d_cond = delta[1:] - ratio ** 2 * a_d
    sa = torch.sqrt(alphas[1:])
    c_xt = torch.zeros_like(betas)
    c_xt[1:] = ratio * delta[:-1] / delta[1:] * sa + (1 - m[:-1]) * (d_cond / delta[1:]) * (1 / sa)
    c_yt = torch.zeros_like(betas)
    c_yt[1:] = (m[:-1] * delta[1:] - m[1:] * ratio * a_d) * sab[:-1] / delta[1:]
    c_eps = torch.zeros_like(betas)
    c_eps[1:] = (1 - m[:-1]) * d_cond / delta[1:] * torch.sqrt(1 - alpha_bar[1:]) / sa
    d_est = torch.zeros_like(betas)
    d_est[1:] = d_cond * delta[:-1] / delta[1:]
    out.update(m=m, sqrt_delta=torch.sqrt(delta), c_xt=c_xt, c_yt=c_yt, c_epst=c_eps,
               sqrt_delta_estimated=torch.sqrt(d_est))
    return out


# --------------------------------------------------------------------------------------
# framing / overlap-add  (UNetModified2.py:5-41)
# --------------------------------------------------------------------------------------
def signal_to_frames(sig: Tensor, frame_len: int, stride: int) -> Tensor:
    """[B,1,n] -> [B,1,n_frames,frame_len]; frame i = samples [i*stride, i*stride+frame_len)."""
    n = sig.shape[-1]
    assert (n - frame_len) % stride == 0                      # UNetModified2.py:13
    return sig.unfold(-1, frame_len, stride).contiguous()


def overlap_add(frames: Tensor, n_samples: int, stride: int) -> Tensor:
    """[B,1,n_frames,F] -> [B,1,n_samples]; un-windowed, un-normalised '+=' of every frame."""
    B, C, nf, Fl = frames.shape
    out = torch.zeros(B, C, n_samples, dtype=frames.dtype)
    for i in range(nf):                                        # UNetModified2.py:37-39
        out[:, :, i * stride:i * stride + Fl] += frames[:, :, i, :]
    return out


# --------------------------------------------------------------------------------------
# UNetModified2 forward  (UNetModified2.py:146-269), functional over a state_dict
# --------------------------------------------------------------------------------------
def swish(x: Tensor) -> Tensor:                                # UNetModified2.py:44-46
    return x * torch.sigmoid(x)


def positional_encoding(noise_level: Tensor, dim: int) -> Tensor:
    """UNetModified2.py:49-68 — [B,...] -> [B, dim] = [sin(nl*v), cos(nl*v)], v_k = 1e4 * 10^(-4k/half)."""
    half = dim // 2
    step = torch.arange(half)
    v = 1e4 * 10.0 ** (-step * 4.0 / half)
    enc = noise_level.reshape(-1, 1) * v
    return torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)


def _lin(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"])


def noise_level_embedding(sd: Dict[str, Tensor], noise_level: Tensor, inner: int, p: str = "") -> Tensor:
    """UNetModified2.py:168-174 — PE -> Linear -> Swish -> Linear -> Swish; returns [B, inner]."""
    e = positional_encoding(noise_level, inner)
    e = swish(_lin(sd, p + "noise_level_mlp.1", e))
    return swish(_lin(sd, p + "noise_level_mlp.3", e))


def _block(sd, key, x, groups):
    """GroupNorm -> Swish -> (dropout=0) -> conv3x3 pad 1   (UNetModified2.py:113-124)."""
    h = F.group_norm(x, groups, sd[key + ".block.0.weight"], sd[key + ".block.0.bias"], eps=1e-5)
    return F.conv2d(swish(h), sd[key + ".block.3.weight"], sd[key + ".block.3.bias"], padding=1)


def _resnet_block(sd, key, x, temb, groups):
    """UNetModified2.py:127-142 — additive noise embedding between the two blocks; 1x1 skip if Cin != Cout."""
    h = _block(sd, key + ".block1", x, groups)
    h = h + _lin(sd, key + ".noise_func.noise_func.0", temb)[:, :, None, None]
    h = _block(sd, key + ".block2", h, groups)
    if key + ".res_conv.weight" in sd:
        x = F.conv2d(x, sd[key + ".res_conv.weight"], sd[key + ".res_conv.bias"])
    return h + x


def unet_layout(cfg) -> Dict[str, List]:
    """Which index of downs / ups is a ResnetBlock / Downsample / Upsample (UNetModified2.py:177-232)."""
    nm, rb = len(cfg["channel_mults"]), cfg["res_blocks"]
    downs = ["stem"]
    for _ in range(nm):
        downs += ["res"] * rb + ["down"]
    ups = []
    for _ in range(nm):
        ups += ["res", "up"] + ["res"] * rb
    return {"downs": downs, "ups": ups}


def unet_forward(sd: Dict[str, Tensor], cfg: dict, x: Tensor, y_t: Tensor, noise_level: Tensor,
                 prefix: str = "noise_estimate_model.", taps: Optional[dict] = None) -> Tensor:
    """eps_hat = UNetModified2(x=condition, y_t, noise_level)   (UNetModified2.py:237-269).

    ``taps`` (optional dict) receives named intermediate activations for kernel-level parity tests.
    """
    p, g = prefix, cfg["norm_groups"]
    Fl, st, n = cfg["segment_len"], cfg["segment_stride"], cfg["num_samples"]
    h = torch.cat([signal_to_frames(x, Fl, st), signal_to_frames(y_t, Fl, st)], dim=1)
    temb = noise_level_embedding(sd, noise_level, cfg["inner_channel"], p)
    lay = unet_layout(cfg)
    feats = []
    for i, kind in enumerate(lay["downs"]):
        if kind == "stem":
            h = F.conv2d(h, sd[f"{p}downs.0.weight"], sd[f"{p}downs.0.bias"], padding=1)
        elif kind == "res":
            h = _resnet_block(sd, f"{p}downs.{i}", h, temb, g)
        else:                                                  # Downsample: conv3x3 stride 2 pad 1
            h = F.conv2d(h, sd[f"{p}downs.{i}.conv.weight"], sd[f"{p}downs.{i}.conv.bias"], stride=2, padding=1)
        feats.append(h)
        if taps is not None:
            taps[f"downs.{i}"] = h
    h = _resnet_block(sd, f"{p}mid.0", h, temb, g)
    if taps is not None:
        taps["mid.0"] = h
    for i, kind in enumerate(lay["ups"]):
        if kind == "res":
            h = _resnet_block(sd, f"{p}ups.{i}", torch.cat((h, feats.pop()), dim=1), temb, g)
        else:                                                  # Upsample: nearest x2 then conv3x3
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[f"{p}ups.{i}.conv.weight"], sd[f"{p}ups.{i}.conv.bias"], padding=1)
        if taps is not None:
            taps[f"ups.{i}"] = h
    h = _block(sd, f"{p}final_conv", h, g)
    if taps is not None:
        taps["final_conv"] = h
    return overlap_add(h, n, st)


# --------------------------------------------------------------------------------------
# diffusion steps  (diffusion.py:164-222, 281-320) — noise is always INJECTED here
# --------------------------------------------------------------------------------------
def get_x_T(sch, T: int, condition: Tensor, noise: Tensor) -> Tensor:          # diffusion.py:281-300
    s = sch["sqrt_alpha_bar"][T]
    return s * condition + torch.sqrt(1.0 - torch.square(s)) * noise


def get_x_T_conditional(sch, T: int, condition: Tensor, noise: Tensor) -> Tensor:   # diffusion.py:302-320
    return sch["sqrt_alpha_bar"][T] * condition + sch["sqrt_delta"][T] * noise


def p_transition(sch, x_t: Tensor, t: int, eps: Tensor, noise: Optional[Tensor], variant: str = "original",
                 condition: Optional[Tensor] = None) -> Tensor:
    """One reverse step incl. the in-place clamp to [-1,1]; ``noise`` is ignored for t == 1."""
    if variant in ("original", "condition_in"):                # diffusion.py:177-190
        x = (x_t - sch["predicted_noise_coeff"][t] * eps) / (sch["alphas"][t]) ** 0.5
        if t > 1:
            x = x + sch["sigma"][t] * noise
    elif variant == "sr3":                                     # diffusion.py:164-175
        x = (x_t - sch["predicted_noise_coeff"][t] * eps) / (sch["alphas"][t]) ** 0.5
        if t > 1:
            x = x + torch.sqrt(sch["betas"][t]) * noise
    elif variant == "supportive":                              # diffusion.py:192-209
        mu = x_t - sch["predicted_noise_coeff"][t] * eps
        g = sch["supportive_gamma"][t]
        x = ((1 - g) * mu + g * condition) / (sch["alphas"][t]) ** 0.5
        if t > 1:
            x = x + max(0, sch["supportive_sigma_hat"][t]) * noise
    elif variant == "conditional":                             # diffusion.py:211-222
        x = sch["c_xt"][t] * x_t + sch["c_yt"][t] * condition - sch["c_epst"][t] * eps
        if t > 1:
            x = x + sch["sqrt_delta_estimated"][t] * noise
    else:
        raise NotImplementedError(variant)
    return x.clamp(-1.0, 1.0)


def sample(sd: Dict[str, Tensor], cfg: dict, sch: Dict[str, Tensor], condition: Tensor, noises: Tensor,
           variant: str = "condition_in", prefix: str = "noise_estimate_model.",
           trace: Optional[dict] = None) -> Tensor:
    """SDDM.infer (model.py:50-124, non-continuous branch) with injected noise.

    noises[0] feeds the x_T draw; noises[k], k=1..T-1, feeds step t = T+1-k (t = T..2); t = 1 draws nothing.
    For 'supportive' x_T = condition (model.py:65-67) and noises[0] is unused; for 'original'/'sr3' x_T = noises[0].
    """
    T = sch["betas"].numel() - 1
    if variant == "condition_in":
        x = get_x_T(sch, T, condition, noises[0])
    elif variant == "conditional":
        x = get_x_T_conditional(sch, T, condition, noises[0])
    elif variant == "supportive":
        x = condition
    else:
        x = noises[0]
    B = condition.shape[0]
    for t in range(T, 0, -1):
        nl = sch["sqrt_alpha_bar"][t] * torch.ones(B, 1, 1)
        eps = unet_forward(sd, cfg, condition, x, nl, prefix)
        z = noises[T + 1 - t] if t > 1 else None
        if trace is not None:
            trace.setdefault("eps", {})[t] = eps
        x = p_transition(sch, x, t, eps, z, variant, condition)
        if trace is not None:
            trace.setdefault("x", {})[t - 1] = x
    return x


# --------------------------------------------------------------------------------------
# forward diffusion of the training step  (diffusion.py:225-279; the random draws are arguments)
# --------------------------------------------------------------------------------------
def q_stochastic(sch, x_0: Tensor, noise: Tensor, t: Tensor, random_step: Optional[Tensor]):
    """t: int64 [B] in [1, T]; random_step: float [B] in [0, 1) or None (= t_is_integer).  Returns (x_t, sample, t + step)."""
    shape = [x_0.shape[0]] + [1] * (x_0.ndim - 1)
    if random_step is None:
        sample, random_step = sch["sqrt_alpha_bar"][t], 0
    else:
        l_a, l_b = sch["sqrt_alpha_bar"][t - 1], sch["sqrt_alpha_bar"][t]
        sample = l_a + random_step * (l_b - l_a)
    sample = sample.view(shape)
    x_t = sample * x_0 + torch.sqrt(1.0 - torch.square(sample)) * noise
    return x_t, sample, (t + random_step).view(shape)


def q_stochastic_conditional(sch, x_0: Tensor, y: Tensor, noise: Tensor, t: Tensor):
    """t: int64 [B,1,1].  Returns (x_t, combined_noise, sqrt_alpha_bar[t])."""
    gaussian_noise = sch["sqrt_delta"][t] * noise
    noise_from_condition = sch["m"][t] * sch["sqrt_alpha_bar"][t] * (y - x_0)
    x_t = sch["sqrt_alpha_bar"][t] * x_0 + noise_from_condition + gaussian_noise
    combined = 1.0 / (torch.sqrt(1.0 - sch["alpha_bar"][t])) * (noise_from_condition + gaussian_noise)
    return x_t, combined, sch["sqrt_alpha_bar"][t]


def sisnr(s_hat: Tensor, s: Tensor) -> Tensor:                                  # model/metric.py:5-34
    s_hat = s_hat.reshape(s_hat.shape[0], 1, -1)
    s = s.reshape(s.shape[0], 1, -1)
    s_hat = s_hat - s_hat.mean(-1, keepdim=True)
    s = s - s.mean(-1, keepdim=True)
    proj = (s_hat * s).sum(-1, keepdim=True) * s / (s ** 2).sum(-1, keepdim=True)
    noise = s_hat - proj
    return (10 * torch.log10((proj ** 2).sum(-1, keepdim=True) / (noise ** 2).sum(-1, keepdim=True))).mean().squeeze()


# --------------------------------------------------------------------------------------
# synthetic, config-shaped weights (no reference needed): default torch init under a seed
# --------------------------------------------------------------------------------------
def random_state_dict(cfg: dict, seed: int = 0, prefix: str = "noise_estimate_model.") -> Dict[str, Tensor]:
    """Random, config-shaped UNetModified2 weights with non-trivial GroupNorm affine terms.

    NOT the reference initialisation (the host package reproduces that one; see
    sddm_b200.model.network).  Used by tests that only need *some* weights on both sides.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def conv(key, co, ci, k):
        bound = 1.0 / math.sqrt(ci * k * k)
        sd[key + ".weight"] = (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound
        sd[key + ".bias"] = (torch.rand(co, generator=g) * 2 - 1) * bound

    def lin(key, co, ci):
        bound = 1.0 / math.sqrt(ci)
        sd[key + ".weight"] = (torch.rand(co, ci, generator=g) * 2 - 1) * bound
        sd[key + ".bias"] = (torch.rand(co, generator=g) * 2 - 1) * bound

    def gn(key, c):
        sd[key + ".weight"] = 1.0 + 0.2 * torch.randn(c, generator=g)
        sd[key + ".bias"] = 0.1 * torch.randn(c, generator=g)

    def res(key, ci, co, emb):
        lin(key + ".noise_func.noise_func.0", co, emb)
        gn(key + ".block1.block.0", ci)
        conv(key + ".block1.block.3", co, ci, 3)
        gn(key + ".block2.block.0", co)
        conv(key + ".block2.block.3", co, co, 3)
        if ci != co:
            conv(key + ".res_conv", co, ci, 1)

    p, inner, mults, rb = prefix, cfg["inner_channel"], list(cfg["channel_mults"]), cfg["res_blocks"]
    lin(p + "noise_level_mlp.1", inner * 4, inner)
    lin(p + "noise_level_mlp.3", inner, inner * 4)
    conv(p + "downs.0", inner, cfg["in_channel"], 3)
    feat, cin, idx = [inner], inner, 1
    for mlt in mults:
        cout = inner * mlt
        for _ in range(rb):
            res(f"{p}downs.{idx}", cin, cout, inner); idx += 1
            feat.append(cout); cin = cout
        conv(f"{p}downs.{idx}.conv", cout, cout, 3); idx += 1
        feat.append(cout)
    res(f"{p}mid.0", cin, cin, inner)
    idx = 0
    cout = cin
    for lvl in reversed(range(len(mults))):
        cin = inner * mults[lvl]
        cout = cin
        res(f"{p}ups.{idx}", cin + feat.pop(), cout, inner); idx += 1
        conv(f"{p}ups.{idx}.conv", cout, cout, 3); idx += 1
        cout = inner if lvl == 0 else inner * mults[lvl - 1]
        for _ in range(rb):
            res(f"{p}ups.{idx}", cin + feat.pop(), cout, inner); idx += 1
            cin = cout
    gn(p + "final_conv.block.0", cout)
    conv(p + "final_conv.block.3", cfg["out_channel"], cout, 3)
    return sd
