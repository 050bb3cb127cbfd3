"""Golden vectors of the inverse STFT: outputs of torch.stft / torch.istft themselves (the arithmetic behind torchaudio's
Spectrogram that the reference's prepare_spectrogram.py:20-35 constructs; the reference never inverts a spectrogram, so torch.istft
is the definition).  Build container only:
    python tests/golden/make_golden_istft.py
Writes tests/golden/istft.npz: complex spectrograms (hamming and hann windows, hop 256 / 128, ragged lengths) and torch.istft outputs."""
import os

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
g = torch.Generator().manual_seed(2048)
out = {}
for tag, (wfn, hop, n) in {"hamming256": (torch.hamming_window, 256, 6000), "hann128": (torch.hann_window, 128, 3333),
                           "hann256": (torch.hann_window, 256, 2560)}.items():
    win = wfn(1024)
    wav = 0.2 * torch.randn(2, n, generator=g)
    spec = torch.stft(wav, 1024, hop, window=win, center=True, pad_mode="reflect", return_complex=True)
    # a spectrogram that is NOT the STFT of a signal (inconsistent frames): the inverse is then a genuine least-squares overlap-add
    spec2 = spec * (1.0 + 0.3 * torch.randn(spec.shape, generator=g)) + 0.05 * torch.view_as_complex(torch.randn(spec.shape + (2,), generator=g))
    length = hop * (spec.shape[-1] - 1)
    out[tag + ".spec"] = torch.view_as_real(spec2).numpy()
    out[tag + ".hop"] = np.asarray(hop)
    out[tag + ".window"] = win.numpy()
    out[tag + ".istft"] = torch.istft(spec2, 1024, hop, window=win, center=True, length=length).numpy()
    out[tag + ".wav"] = wav.numpy()
    out[tag + ".stft"] = torch.view_as_real(spec).numpy()
    print(tag, tuple(spec.shape), length)
np.savez_compressed(os.path.join(OUT, "istft.npz"), **out)
print(os.path.getsize(os.path.join(OUT, "istft.npz")))
