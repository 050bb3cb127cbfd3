"""Generates the golden vectors under tests/golden/ by running the REAL reference modules.

Only runnable in the build container (needs /root/reference, read-only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
The GPU box has no /root/reference; tests there use the committed .npz/.json files produced here.

What is pinned (reference file:line):
  * schedule tables ........ GaussianDiffusion.__init__        model/diffusion.py:49-161
  * framing / overlap-add .. SignalToFrames                    model/UNetModified2.py:5-41 (toy of model/tstnn.py:302-308)
  * default-init weights ... UNetModified2.__init__ under torch.manual_seed(0) (checksums only)
  * eps_hat ................ UNetModified2.forward             model/UNetModified2.py:237-269
  * elementwise steps ...... get_x_T / p_transition*           model/diffusion.py:164-222,281-320
  * full sampling .......... SDDM.infer                        model/model.py:50-124 (noise injected via randn_like)
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import model.diffusion as ref_diffusion  # noqa: E402
import model.model as ref_model          # noqa: E402
import model.network as ref_network      # noqa: E402
from model.UNetModified2 import SignalToFrames  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
UNET_ARGS = dict(in_channel=2, out_channel=1, inner_channel=32, norm_groups=32, channel_mults=[1, 2, 3, 4, 5],
                 res_blocks=1, dropout=0, segment_len=128, segment_stride=64)
L = 16448
BUFS = ("betas", "alphas", "alpha_bar", "sqrt_alpha_bar", "predicted_noise_coeff", "sigma", "supportive_gamma",
        "supportive_sigma_hat", "m", "sqrt_delta", "c_xt", "c_yt", "c_epst", "sqrt_delta_estimated")


def digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


class InjectedNoise:
    """Replaces torch.randn_like inside the reference by an iterator over a pre-drawn tensor."""

    def __init__(self, noises):
        self.noises, self.k = noises, 0

    def __call__(self, like, **kw):
        z = self.noises[self.k].reshape(like.shape)
        self.k += 1
        return z


def with_injected(noises, fn):
    orig = torch.randn_like
    inj = InjectedNoise(noises)
    torch.randn_like = inj
    try:
        return fn(), inj.k
    finally:
        torch.randn_like = orig


def main():
    torch.set_num_threads(8)
    meta = {}

    # 1. schedules ---------------------------------------------------------------------------------------
    sched = {}
    for tag, kw in {"linear100": dict(schedule="linear", n_timestep=100, linear_start=1e-6, linear_end=1e-3),
                    "quad50": dict(schedule="quad", n_timestep=50, linear_start=1e-4, linear_end=2e-2),
                    "cosine20": dict(schedule="cosine", n_timestep=20, linear_start=1e-4, linear_end=2e-2),
                    "linear6": dict(schedule="linear", n_timestep=6, linear_start=1e-4, linear_end=5e-2)}.items():
        d = ref_diffusion.GaussianDiffusion(device="cpu", **kw)
        for b in BUFS:
            sched[f"{tag}.{b}"] = getattr(d, b).numpy()
        meta[f"schedule.{tag}"] = {k: (v if not isinstance(v, float) else v) for k, v in kw.items()}
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **sched)

    # 2. framing / OLA known answer --------------------------------------------------------------------------
    seg = SignalToFrames(10, 4, 2)
    sig = torch.arange(1, 11, dtype=torch.float32).reshape(1, 1, 10)
    fr = seg(sig)
    ola = seg.overlapAdd(fr)
    meta["framing_toy"] = {"signal": sig.flatten().tolist(), "frames": fr[0, 0].tolist(), "overlap_add": ola.flatten().tolist()}

    # 3. default-init weights under seed 0 (checksums) -----------------------------------------------------------
    torch.manual_seed(0)
    diff = ref_diffusion.GaussianDiffusion(schedule="linear", n_timestep=100, linear_start=1e-6, linear_end=1e-3, device="cpu")
    net = ref_network.UNetModified2(num_samples=L, **UNET_ARGS)
    model = ref_model.SDDM(diff, net, p_transition="condition_in").eval()
    sd = model.state_dict()
    meta["weights_seed0"] = {k: {"shape": list(v.shape), "sha256": digest(v)} for k, v in sd.items()}
    meta["pe_vector_sha256"] = digest(net.noise_level_mlp[0].embedding_vector)

    # 4. eps_hat of one forward -------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(7)
    x = (0.1 * torch.randn(2, 1, L, generator=g)).clamp(-1, 1)
    y = (0.3 * torch.randn(2, 1, L, generator=g)).clamp(-1, 1)
    nl = torch.tensor([diff.sqrt_alpha_bar[100].item(), 0.99]).reshape(2, 1, 1)
    taps = {}
    hooks = []
    for name in ("downs.0", "downs.1", "downs.2", "downs.5", "downs.10", "mid.0", "ups.0", "ups.1", "ups.7", "ups.14"):
        mod = net.get_submodule(name)
        hooks.append(mod.register_forward_hook(lambda m, i, o, name=name: taps.__setitem__(name, o.detach())))
    with torch.no_grad():
        eps = net(x, y, nl)
    for h in hooks:
        h.remove()
    meta["eps_inputs"] = {"seed": 7, "x_sha256": digest(x), "y_sha256": digest(y), "noise_level": nl.flatten().tolist()}
    tap_stats = {k: {"mean": float(v.mean()), "std": float(v.std()), "absmax": float(v.abs().max()),
                     "probe": v[:, : min(4, v.shape[1]), :2, :2].flatten().tolist()} for k, v in taps.items()}
    meta["eps_taps"] = tap_stats
    np.savez_compressed(os.path.join(OUT, "unet_eps.npz"), eps=eps.numpy())

    # 5. elementwise steps (all variants) -------------------------------------------------------------------------
    g = torch.Generator().manual_seed(11)
    Ls = 2048
    xs = torch.randn(2, 1, Ls, generator=g).clamp(-1, 1) * 0.9
    es = torch.randn(2, 1, Ls, generator=g)
    cs = (0.2 * torch.randn(2, 1, Ls, generator=g)).clamp(-1, 1)
    zs = torch.randn(2, 1, Ls, generator=g)
    steps = {"x": xs.numpy(), "eps": es.numpy(), "cond": cs.numpy(), "z": zs.numpy()}
    for t in (1, 2, 50, 100):
        for variant, fn in {"original": lambda: diff.p_transition(xs.clone(), t, es),
                            "sr3": lambda: diff.p_transition_sr3(xs.clone(), t, es),
                            "supportive": lambda: diff.p_transition_supportive(xs.clone(), t, es, cs),
                            "conditional": lambda: diff.p_transition_conditional(xs.clone(), t, es, cs)}.items():
            out, _ = with_injected([zs], fn)
            steps[f"{variant}.t{t}"] = out.numpy()
    out, _ = with_injected([zs], lambda: diff.get_x_T(cs))
    steps["get_x_T"] = out.numpy()
    out, _ = with_injected([zs], lambda: diff.get_x_T_conditional(cs))
    steps["get_x_T_conditional"] = out.numpy()
    np.savez_compressed(os.path.join(OUT, "steps.npz"), **steps)

    # 6. full sampling, config_unet.json shape, cfg-1 input ------------------------------------------
    clip = 0.05 * torch.randn(1, 32000, generator=torch.Generator().manual_seed(1))
    cond = torch.nn.functional.pad(clip, (0, 2 * L - 32000)).view(2, 1, L)
    noises = torch.randn(100, 2, 1, L, generator=torch.Generator().manual_seed(1234))
    trace = {}
    orig_fwd = net.forward

    def traced(xc, yt, lvl):
        e = orig_fwd(xc, yt, lvl)
        trace.setdefault("eps", []).append(e.detach().clone())
        return e

    net.forward = traced
    (out, used) = with_injected(noises, lambda: model.infer(cond))
    net.forward = orig_fwd
    assert used == 100, used
    meta["sample_cfg1"] = {"clip_seed": 1, "noise_seed": 1234, "cond_sha256": digest(cond), "noises_sha256": digest(noises),
                           "noise_draws": used}
    np.savez_compressed(os.path.join(OUT, "sample_cfg1.npz"), out=out.numpy(), eps_t100=trace["eps"][0].numpy(),
                        eps_t50=trace["eps"][50].numpy(), eps_t1=trace["eps"][99].numpy())

    # 7. short schedules: every p_transition variant through SDDM.infer (T = 6, one chunk) ---------------------------
    short = {}
    d6 = ref_diffusion.GaussianDiffusion(schedule="linear", n_timestep=6, linear_start=1e-4, linear_end=5e-2, device="cpu")
    c1 = cond[:1]
    n6 = torch.randn(6, 1, 1, L, generator=torch.Generator().manual_seed(99))
    for variant in ("original", "condition_in", "sr3", "supportive", "conditional"):
        m6 = ref_model.SDDM(d6, net, p_transition=variant).eval()
        (o6, used) = with_injected(n6, lambda: m6.infer(c1))
        short[variant] = o6.numpy()
        meta.setdefault("sample_T6", {})[variant] = {"noise_draws": used}
    meta["sample_T6"]["noise_seed"] = 99
    np.savez_compressed(os.path.join(OUT, "sample_T6.npz"), **short)

    meta["torch"] = torch.__version__
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
