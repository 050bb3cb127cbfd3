"""STFT feature front-end (prepare_spectrogram.py:20-55).

CPU: the oracle restatement (oracle/stft_oracle.py) against goldens written by torchaudio's own transforms
(tests/golden/make_golden_stft.py) and the host-side window / filterbank construction.
GPU: the CUDA kernel (through the C ABI / the host mirror classes) against the goldens and the oracle.
Tolerances: linear magnitudes max|d| <= 2e-5 * max|ref| (fp32 FFT vs fp32 FFT); log/clamp features max|d| <= 2e-3 (the log10
amplifies relative error of near-zero bins; features live in [0, 1])."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import stft_oracle as SO  # noqa: E402

CFG = {"sample_rate": 16000, "spectrogram": {"window_length": 1024, "hop_samples": 256}, "mel_spectrogram": {"n_mels": 80}}


def test_oracle_matches_torchaudio_goldens(golden):
    g = golden("stft.npz")
    spec = SO.spectrogram(g["wav"], 1024, 256, "hamming")
    mel = SO.mel_spectrogram(g["wav"], 1024, 256, 80, 16000)
    assert spec.shape == g["spec"].shape and mel.shape == g["mel"].shape
    assert float((spec - g["spec"]).abs().max()) <= 2e-5 * float(g["spec"].abs().max())
    assert float((mel - g["mel"]).abs().max()) <= 2e-5 * float(g["mel"].abs().max())
    assert float((SO.log_clamp(spec) - g["spec_feat"]).abs().max()) <= 2e-3
    assert float((SO.log_clamp(mel) - g["mel_feat"]).abs().max()) <= 2e-3
    assert torch.equal(SO.melscale_fbanks(513, 20.0, 8000.0, 80, 16000), g["mel_fb"])


def test_host_window_and_filterbank_match_torchaudio(golden):
    from sddm_b200 import prepare_spectrogram as PS
    g = golden("stft.npz")
    assert torch.equal(PS.melscale_fbanks(513, 20.0, 8000.0, 80, 16000), g["mel_fb"])
    assert float((SO.window("hamming", 1024) - torch.hamming_window(1024)).abs().max()) < 1e-6
    assert float((SO.window("hann", 1024) - torch.hann_window(1024)).abs().max()) < 1e-6
    with pytest.raises(NotImplementedError):
        PS.Spectrogram(n_fft=1024, power=2)
    with pytest.raises(RuntimeError):
        PS.Spectrogram(n_fft=1024, hop_length=256)(torch.zeros(1, 4000))          # CPU tensor: no fallback


@pytest.mark.gpu
def test_stft_kernel_vs_goldens_and_oracle(golden, built_lib):
    from sddm_b200 import prepare_spectrogram as PS
    dev = torch.device("cuda:0")
    g = golden("stft.npz")
    wav = g["wav"].to(dev)
    spec = PS.Spectrogram(n_fft=1024, hop_length=256, window_fn=torch.hamming_window, power=1, normalized=True)(wav).cpu()
    mel = PS.MelSpectrogram(n_fft=1024, hop_length=256, f_min=20.0, f_max=8000.0, n_mels=80, sample_rate=16000, power=1.0,
                            normalized=True)(wav).cpu()
    e_spec = float((spec - g["spec"]).abs().max() / g["spec"].abs().max())
    e_mel = float((mel - g["mel"]).abs().max() / g["mel"].abs().max())
    mel_f, spec_f = PS.features(wav, CFG)
    e_sf = float((spec_f.cpu() - g["spec_feat"]).abs().max())
    e_mf = float((mel_f.cpu() - g["mel_feat"]).abs().max())
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.txt"), "a") as f:
        f.write(f"stft vs torchaudio golden: spec {e_spec:.2e}, mel {e_mel:.2e} (rel to max); features spec {e_sf:.2e}, mel {e_mf:.2e} (abs)\n")
    assert e_spec <= 2e-5 and e_mel <= 2e-5 and e_sf <= 2e-3 and e_mf <= 2e-3
    # other lengths / hops / batch: ragged last frame block, L not a multiple of hop, reflect padding at both ends
    gen = torch.Generator().manual_seed(3)
    for B, L, hop in ((1, 513, 256), (3, 16448, 256), (2, 160000, 256), (2, 5000, 100)):
        x = 0.1 * torch.randn(B, L, generator=gen)
        got = PS.Spectrogram(n_fft=1024, hop_length=hop, window_fn=torch.hamming_window)(x.to(dev)).cpu()
        ref = SO.spectrogram(x, 1024, hop, "hamming")
        assert got.shape == ref.shape == (B, 513, 1 + L // hop)
        assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max()), (B, L, hop)
    # linearity and a pure tone landing in the right bin (size-independent properties)
    x = 0.1 * torch.randn(1, 20000, generator=gen)
    S = PS.Spectrogram(n_fft=1024, hop_length=256, window_fn=torch.hamming_window)
    assert torch.allclose(S((2.5 * x).to(dev)), 2.5 * S(x.to(dev)), rtol=1e-5, atol=1e-7)
    t = torch.arange(20000) / 16000.0
    tone = torch.sin(2 * np.pi * 1000.0 * t)[None]
    assert int(S(tone.to(dev))[0, :, 30].argmax()) == 64                              # 1000 Hz / (16000 / 1024) = 64
    with pytest.raises(Exception):
        PS.Spectrogram(n_fft=512, hop_length=128)(x.to(dev))


# ----------------------------------------------------------------------------------------------------
# inverse STFT (north_star names it next to the front-end; torch.istft is the definition)
# ----------------------------------------------------------------------------------------------------
def _istft_cases():
    g = np.load(os.path.join(GOLDEN, "istft.npz"))
    for tag in ("hamming256", "hann128", "hann256"):
        spec = torch.view_as_complex(torch.from_numpy(g[tag + ".spec"]).contiguous())
        yield tag, spec, int(g[tag + ".hop"]), torch.from_numpy(g[tag + ".window"]), torch.from_numpy(g[tag + ".istft"]), \
            torch.from_numpy(g[tag + ".wav"]), torch.view_as_complex(torch.from_numpy(g[tag + ".stft"]).contiguous())


def test_istft_oracle_vs_torch_golden():
    for tag, spec, hop, win, want, _, _ in _istft_cases():
        got = SO.istft(spec, 1024, hop, win, want.shape[-1])
        assert float((got - want).abs().max()) < 2e-6 * float(want.abs().max()) + 1e-7, tag


@pytest.mark.gpu
def test_istft_kernel_vs_torch_golden_and_round_trip(built_lib):
    from sddm_b200 import prepare_spectrogram as PS
    for tag, spec, hop, win, want, wav, stft in _istft_cases():
        got = PS.istft(spec.cuda(), 1024, hop, win, want.shape[-1]).cpu()
        err = float((got - want).abs().max() / want.abs().max())
        assert err < 5e-6, (tag, err)
        # perfect reconstruction: the inverse of a true STFT returns the signal (away from the last partial hop)
        n = hop * (stft.shape[-1] - 1)
        back = PS.istft(stft.cuda(), 1024, hop, win, n).cpu()
        assert float((back - wav[:, :n]).abs().max()) < 2e-5, tag
        # linearity
        a = PS.istft((2.0 * spec).cuda(), 1024, hop, win, want.shape[-1]).cpu()
        assert float((a - 2.0 * got).abs().max()) < 1e-5 * float(got.abs().max())
    with pytest.raises(RuntimeError):
        PS.istft(spec, 1024, hop, win)          # CPU tensor: no fallback
