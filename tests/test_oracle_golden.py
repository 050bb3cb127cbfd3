"""Pins the CPU oracle (oracle/sddm_oracle.py) against golden vectors produced by the real reference
(tests/golden/make_golden.py) — the oracle must be right before any CUDA result is compared with it."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import L, UNET_CFG, cfg1_condition, rel_err, seed0_state_dict

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import sddm_oracle as O  # noqa: E402

SCHEDULES = {"linear100": ("linear", 100, 1e-6, 1e-3), "quad50": ("quad", 50, 1e-4, 2e-2),
             "cosine20": ("cosine", 20, 1e-4, 2e-2), "linear6": ("linear", 6, 1e-4, 5e-2)}


def sha(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


@pytest.mark.parametrize("tag", list(SCHEDULES))
def test_schedule_tables_bit_exact(golden, tag):
    g = golden("schedules.npz")
    s = O.make_schedule(*SCHEDULES[tag])
    for name in O.DIFFUSION_BUFFERS:
        ref = g[f"{tag}.{name}"]
        assert s[name].shape == ref.shape
        assert torch.equal(s[name], ref) or (torch.isnan(ref) == torch.isnan(s[name])).all() and \
            torch.equal(torch.nan_to_num(s[name]), torch.nan_to_num(ref)), name


def test_schedule_known_answers():
    """SURVEY.md §8a row 1 constants (computed from the reference)."""
    s = O.make_schedule("linear", 100, 1e-6, 1e-3)
    assert abs(s["sqrt_alpha_bar"][100].item() - 0.975277364) < 1e-8
    assert abs(s["sqrt_alpha_bar"][50].item() - 0.993812501) < 1e-8
    assert abs(s["sigma"][100].item() - 0.031313002) < 1e-8
    assert s["sigma"][1].item() == 0.0
    assert abs(s["predicted_noise_coeff"][100].item() - 4.525207e-3) < 1e-8
    assert abs(s["alphas"][100].item() - 0.999000013) < 1e-8
    with pytest.raises(NotImplementedError):
        O.make_schedule("warmup10", 10)


def test_framing_known_answer(meta):
    """Reference's only executable check: model/tstnn.py:302-308."""
    toy = meta["framing_toy"]
    sig = torch.tensor(toy["signal"]).reshape(1, 1, -1)
    fr = O.signal_to_frames(sig, 4, 2)
    assert fr[0, 0].tolist() == toy["frames"] == [[1, 2, 3, 4], [3, 4, 5, 6], [5, 6, 7, 8], [7, 8, 9, 10]]
    assert O.overlap_add(fr, 10, 2).flatten().tolist() == toy["overlap_add"] == [1, 2, 6, 8, 10, 12, 14, 16, 9, 10]
    with pytest.raises(AssertionError):
        O.signal_to_frames(torch.zeros(1, 1, 11), 4, 2)


def test_steps_bit_exact(golden):
    g = golden("steps.npz")
    sch = O.make_schedule("linear", 100, 1e-6, 1e-3)
    for t in (1, 2, 50, 100):
        for variant in ("original", "sr3", "supportive", "conditional"):
            out = O.p_transition(sch, g["x"], t, g["eps"], g["z"], variant, g["cond"])
            assert torch.equal(out, g[f"{variant}.t{t}"]), (variant, t)
    assert torch.equal(O.get_x_T(sch, 100, g["cond"], g["z"]), g["get_x_T"])
    assert torch.equal(O.get_x_T_conditional(sch, 100, g["cond"], g["z"]), g["get_x_T_conditional"])


def test_seed0_weights_match_reference(meta):
    sd, _ = seed0_state_dict()
    ref = {k: v for k, v in meta["weights_seed0"].items() if k.startswith("noise_estimate_model.")}
    assert set(sd) == set(ref)
    for k, v in sd.items():
        assert list(v.shape) == ref[k]["shape"], k
        assert sha(v) == ref[k]["sha256"], k


def test_unet_eps_matches_reference(golden, meta):
    sd, _ = seed0_state_dict()
    g = torch.Generator().manual_seed(meta["eps_inputs"]["seed"])
    x = (0.1 * torch.randn(2, 1, L, generator=g)).clamp(-1, 1)
    y = (0.3 * torch.randn(2, 1, L, generator=g)).clamp(-1, 1)
    assert sha(x) == meta["eps_inputs"]["x_sha256"] and sha(y) == meta["eps_inputs"]["y_sha256"]
    nl = torch.tensor(meta["eps_inputs"]["noise_level"]).reshape(2, 1, 1)
    taps = {}
    eps = O.unet_forward(sd, dict(UNET_CFG), x, y, nl, taps=taps)
    assert rel_err(eps, golden("unet_eps.npz")["eps"]) < 2e-6      # fp32, same ATen kernels underneath
    for name, st in meta["eps_taps"].items():
        assert abs(float(taps[name].std()) - st["std"]) < 1e-5 * max(1.0, st["std"]), name


def test_sampling_T6_all_variants(golden):
    sd, _ = seed0_state_dict()
    g = golden("sample_T6.npz")
    sch = O.make_schedule("linear", 6, 1e-4, 5e-2)
    cond = cfg1_condition()[:1]
    noises = torch.randn(6, 1, 1, L, generator=torch.Generator().manual_seed(99))
    for variant in ("original", "condition_in", "sr3", "supportive", "conditional"):
        out = O.sample(sd, dict(UNET_CFG), sch, cond, noises, variant)
        assert rel_err(out, g[variant]) < 1e-5, variant


def test_sampling_cfg1_full_schedule(golden, meta):
    """Full 100-step config_unet.json sampling of the cfg-1 clip (2 chunks) against the reference's SDDM.infer."""
    sd, _ = seed0_state_dict()
    g = golden("sample_cfg1.npz")
    cond = cfg1_condition()
    noises = torch.randn(100, 2, 1, L, generator=torch.Generator().manual_seed(1234))
    assert sha(cond) == meta["sample_cfg1"]["cond_sha256"] and sha(noises) == meta["sample_cfg1"]["noises_sha256"]
    trace = {}
    out = O.sample(sd, dict(UNET_CFG), O.make_schedule("linear", 100, 1e-6, 1e-3), cond, noises, "condition_in", trace=trace)
    assert rel_err(trace["eps"][100], g["eps_t100"]) < 2e-6
    assert rel_err(trace["eps"][50], g["eps_t50"]) < 1e-4
    assert rel_err(out, g["out"]) < 1e-4
    assert float(O.sisnr(out, g["out"])) > 80.0


def test_random_state_dict_layout(meta):
    sd = O.random_state_dict(dict(UNET_CFG), seed=3)
    ref = {k: v["shape"] for k, v in meta["weights_seed0"].items() if k.startswith("noise_estimate_model.")}
    assert {k: list(v.shape) for k, v in sd.items()} == ref


def test_oracles_vs_reference_at_baseline_sizes(golden):
    """The three oracles against reference outputs at BASELINE's full sizes (tests/golden/make_golden_fullsize.py):
    cfg 2 rows 0/31/63 at t = 100 (x_T + one eps_hat), cfg 4 at F = 107, cfg 5 on one 10 s utterance."""
    import diffwave_oracle as DO
    import wavegrad_oracle as WO
    from conftest import (CFG2_GOLDEN_ROWS, DIFFWAVE_CASES, cfg2_inputs, cfg4_fullsize_inputs, cfg5_fullsize_inputs,
                          diffwave_test_module, wavegrad_test_module)
    g = golden("fullsize.npz")
    sd, _ = seed0_state_dict()
    cond, noises = cfg2_inputs()
    rows = list(CFG2_GOLDEN_ROWS)
    sch = O.make_schedule("linear", 100, 1e-6, 1e-3)
    with torch.no_grad():
        x = O.get_x_T(sch, 100, cond[rows], noises[0][rows])
        eps = O.unet_forward(sd, dict(UNET_CFG), cond[rows], x, sch["sqrt_alpha_bar"][100] * torch.ones(3, 1, 1))
    assert rel_err(eps, g["cfg2.eps_t100"]) < 1e-5
    del noises
    spec, audio, lv = cfg4_fullsize_inputs()
    wsd = {k: v.detach() for k, v in wavegrad_test_module().state_dict().items()}
    with torch.no_grad():
        e4 = WO.wavegrad_forward(wsd, spec[:2], audio[:2], lv[:2])
    assert rel_err(e4.reshape(g["cfg4.eps_rows01"].shape), g["cfg4.eps_rows01"]) < 1e-5
    spec, audio, step = cfg5_fullsize_inputs()
    dsd = {k: v.detach() for k, v in diffwave_test_module(DIFFWAVE_CASES["full"]).state_dict().items()}
    with torch.no_grad():
        e5 = DO.diffwave_forward(dsd, spec[:1], audio[:1], step[:1])
    assert rel_err(e5, g["cfg5.eps_row0"]) < 1e-5
