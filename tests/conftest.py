import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

UNET_CFG = dict(num_samples=16448, in_channel=2, out_channel=1, inner_channel=32, norm_groups=32,
                channel_mults=(1, 2, 3, 4, 5), res_blocks=1, dropout=0, segment_len=128, segment_stride=64)
L = 16448


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def meta():
    with open(os.path.join(GOLDEN, "meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, name)).items()}
    return load


@pytest.fixture(scope="session")
def built_lib():
    """Builds (if stale) and loads the extension; CPU-only hosts can still load it and call non-compute entry points."""
    import __graft_entry__ as ge
    ge.build()
    from sddm_b200 import _lib
    return _lib.lib()


def seed0_state_dict():
    """Reference-identical default init under torch.manual_seed(0), produced by the host mirror classes."""
    from sddm_b200.model.network import UNetModified2
    torch.manual_seed(0)
    net = UNetModified2(**UNET_CFG)
    sd = {"noise_estimate_model." + k: v.detach().clone() for k, v in net.state_dict().items()}
    return sd, net


def cfg1_condition():
    """SURVEY.md §8d cfg 1: one 32000-sample clip, seed 1, zero padded to 2 chunks."""
    clip = 0.05 * torch.randn(1, 32000, generator=torch.Generator().manual_seed(1))
    return torch.nn.functional.pad(clip, (0, 2 * L - 32000)).view(2, 1, L)


# cfg 5 (DiffWave) parity cases: "full" = config_diffwave.json's network on a short clip; "small" = fewer layers, odd frame
# count, per-row diffusion steps.  Shared by tests/golden/make_golden_diffwave.py (reference side) and the tests.
DIFFWAVE_CASES = {
    "full": dict(freq_bins=513, residual_layers=30, dilation_cycle_length=10, B=2, frames=8, steps=[37.0, 37.0], seed=5,
                 probe_layers=[0, 9, 29]),
    "small": dict(freq_bins=513, residual_layers=6, dilation_cycle_length=3, B=3, frames=5, steps=[3.0, 57.0, 200.0], seed=6,
                  probe_layers=[0, 5]),
}


def diffwave_test_module(case):
    """Host mirror with the reference's default init under torch.manual_seed(0); the zero-initialised output_projection
    weight (diffwave.py:131) is replaced by 0.1 * N(0,1) (seed 1) so that eps_hat depends on the whole network."""
    from sddm_b200.model.network import DiffWave
    torch.manual_seed(0)
    net = DiffWave(num_samples=-1, num_timesteps=200, freq_bins=case["freq_bins"], residual_channels=64,
                   residual_layers=case["residual_layers"], dilation_cycle_length=case["dilation_cycle_length"])
    with torch.no_grad():
        net.output_projection.weight.copy_(0.1 * torch.randn(net.output_projection.weight.shape, generator=torch.Generator().manual_seed(1)))
    return net


# cfg 4 (WaveGrad) parity cases (the network has no constructor arguments): spectrogram frames F -> 300 F samples
WAVEGRAD_CASES = {
    "b2f4": dict(B=2, frames=4, levels=[0.3, 0.9], seed=11),
    "b1f7": dict(B=1, frames=7, levels=[0.62], seed=12),
}


def wavegrad_test_module():
    """Host mirror with the reference's default init under torch.manual_seed(0); the zero-initialised biases
    (wavegrad.py:16,63-64) are replaced by 0.05 * N(0,1) (seed 1) so that every bias path is exercised."""
    from sddm_b200.model.network import WaveGrad
    torch.manual_seed(0)
    net = WaveGrad()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for k, v in net.state_dict().items():
            if k.endswith(".bias"):
                v.copy_(0.05 * torch.randn(v.shape, generator=g))
    return net


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| — the per-tensor error metric of BASELINE.md §3."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


# BASELINE-size parity inputs, shared by tests/golden/make_golden_fullsize.py (reference side) and the -m gpu tests
CFG2_GOLDEN_ROWS = (0, 31, 63)


def cfg2_inputs():
    """SURVEY.md §8d cfg 2: 64 chunks clip(0.1 N(0,1)), seed 0; injected noises [100, 64, 1, L], seed 1234."""
    cond = (0.1 * torch.randn(64, 1, L, generator=torch.Generator().manual_seed(0))).clamp(-1, 1)
    noises = torch.randn(100, 64, 1, L, generator=torch.Generator().manual_seed(1234))
    return cond, noises


def cfg5_fullsize_inputs():
    """cfg 5 size: 10 s utterances, spec [513, 626], 160 256 samples; per-row diffusion steps."""
    g = torch.Generator().manual_seed(12)
    B, frames = 2, 626
    spec = torch.rand(B, 513, frames, generator=g) * 0.7
    audio = torch.randn(B, 1, 256 * frames, generator=g)
    step = torch.tensor([150.0, 20.0]).reshape(B, 1, 1)
    return spec, audio, step


def cfg4_fullsize_inputs():
    """cfg 4 size: spec [128, 107], 32 100 samples; per-row noise levels."""
    g = torch.Generator().manual_seed(13)
    B, F = 3, 107
    spec = torch.rand(B, 128, F, generator=g)
    audio = torch.randn(B, 300 * F, generator=g)
    lv = torch.tensor([0.95, 0.5, 0.1])
    return spec, audio, lv
