"""Golden vectors of the STFT feature front-end, produced by the transforms the reference itself constructs
(prepare_spectrogram.py:20-35: torchaudio TT.Spectrogram / TT.MelSpectrogram).  Build container only (torchaudio CPU):
    python tests/golden/make_golden_stft.py
Writes tests/golden/stft.npz: input waveforms + linear magnitudes + the reference's log/clamp features."""
import os

import numpy as np
import torch
from torchaudio import transforms as TT

OUT = os.path.dirname(os.path.abspath(__file__))
sample_rate, window_length, hop_samples, n_mels = 16000, 1024, 256, 80          # config_diffwave.json
spectrogram = TT.Spectrogram(n_fft=window_length, hop_length=hop_samples, window_fn=torch.hamming_window, power=1, normalized=True)
mel_spec = TT.MelSpectrogram(n_fft=window_length, hop_length=hop_samples, f_min=20.0, f_max=sample_rate / 2.0, n_mels=n_mels,
                             sample_rate=sample_rate, power=1.0, normalized=True)
g = torch.Generator().manual_seed(16000)
t = torch.arange(6000) / sample_rate
tone = 0.3 * torch.sin(2 * np.pi * 440.0 * t) + 0.1 * torch.sin(2 * np.pi * 3000.0 * t)
wav = torch.stack([0.1 * torch.randn(6000, generator=g), tone + 0.01 * torch.randn(6000, generator=g)])
spec, mel = spectrogram(wav), mel_spec(wav)
feat = lambda x: torch.clamp((torch.log10(x) - 1 + 5) / 5, 0.0, 1.0)          # prepare_spectrogram.py:41-44
np.savez_compressed(os.path.join(OUT, "stft.npz"), wav=wav.numpy(), spec=spec.numpy(), mel=mel.numpy(),
                    spec_feat=feat(spec).numpy(), mel_feat=feat(mel).numpy(), mel_fb=mel_spec.mel_scale.fb.numpy())
print("spec", tuple(spec.shape), "mel", tuple(mel.shape), "max", float(spec.max()), float(mel.max()))
