"""Batch-shape contract of the inference path (reference data_loader/data_loaders.py:101-164).

The reference ``InferDataset`` zero-pads every utterance to a multiple of T = num_samples and views it as
``[n_chunk, 1, T]``; ``infer_data_collate`` concatenates the chunks of several utterances along dim 0 and carries an
index tensor (utterance id per chunk); infer.py:81-120 regroups rows by that index.  The reference decodes audio with
torchaudio.load (needs torchcodec, absent in this image): here waveforms come from memory, ``.npy`` files or PCM / float
``.wav`` files read and written with scipy.io.wavfile (``load_wave`` / ``save_wave``).
"""
from __future__ import annotations

from math import ceil
from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def load_wave(path, sample_rate: int = None) -> torch.Tensor:
    """``.npy`` (float waveform) or ``.wav`` (int16 / int32 / uint8 PCM scaled to [-1, 1) as torchaudio.load does, or float) ->
    float32 [1, n].  A sample-rate mismatch raises (the reference resamples nothing either: data_loaders.py:108-111)."""
    path = str(path)
    if path.endswith(".npy"):
        return torch.as_tensor(np.load(path), dtype=torch.float32).reshape(1, -1)
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if sample_rate is not None and sr != sample_rate:
        raise ValueError("%s: sample rate %d != %d" % (path, sr, sample_rate))
    if data.ndim > 1:
        data = data[:, 0]
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(x)).reshape(1, -1)


def save_wave(path, wave: torch.Tensor, sample_rate: int = 16000, bits: int = 32) -> None:
    """float32 waveform [1, n] / [n] -> ``.wav`` or ``.npy``.  bits=32 (default) writes IEEE-float WAV, which is what the
    reference's torchaudio.save(path, float32 tensor, sr) produces (infer.py:115-120): no quantisation ahead of PESQ / SI-SNR
    evaluation; bits=16 writes rounded, clipped 16-bit PCM."""
    x = wave.detach().to("cpu", torch.float32).reshape(-1).numpy()
    path = str(path)
    if path.endswith(".npy"):
        np.save(path, x)
        return
    from scipy.io import wavfile
    if bits == 32:
        wavfile.write(path, sample_rate, np.ascontiguousarray(x, dtype=np.float32))
    elif bits == 16:
        wavfile.write(path, sample_rate, np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16))
    else:
        raise ValueError("bits must be 32 (float) or 16 (PCM)")


def chunk_waveform(wave: torch.Tensor, T: int) -> torch.Tensor:
    """[1, n] or [n] -> [ceil(n/T), 1, T], zero padded at the end (reference :112-118)."""
    wave = wave.reshape(1, -1)
    n = wave.shape[-1]
    n_chunk = max(1, ceil(n / T))
    return F.pad(wave, (0, n_chunk * T - n), "constant", 0).view(n_chunk, 1, T)


def chunk_counts(lengths: Sequence[int], T: int) -> List[int]:
    """Rows each utterance contributes to the [N,1,T] batch: max(1, ceil(n / T)) (reference :112-118)."""
    return [max(1, ceil(int(n) / T)) for n in lengths]


def rows_of_range(waves: Sequence[torch.Tensor], T: int, lo: int, hi: int, out: torch.Tensor = None) -> torch.Tensor:
    """Rows [lo, hi) of the batch InferDataset + infer_data_collate would build from `waves`, WITHOUT building the other rows:
    a rank of a sharded run pads / copies only the chunks it owns.  `out` may be a (pinned) [hi - lo, 1, T] staging buffer."""
    counts = chunk_counts([w.numel() for w in waves], T)
    if out is None:
        out = torch.zeros((hi - lo, 1, T), dtype=torch.float32)
    else:
        out.zero_()
    flat = out.view(hi - lo, T)
    row = 0
    for w, c in zip(waves, counts):
        a, b = max(row, lo), min(row + c, hi)
        if a < b:
            x = torch.as_tensor(w, dtype=torch.float32).reshape(-1)
            s0, s1 = (a - row) * T, min((b - row) * T, x.numel())
            if s1 > s0:
                dst = flat[a - lo:b - lo].reshape(-1)
                dst[: s1 - s0] = x[s0:s1]
        row += c
        if row >= hi:
            break
    return out


class DeviceBatch:
    """The dataset edge on the GPU (C ABI: sddm_chunk_rows / sddm_regroup_rows): the utterances of a run are uploaded ONCE, back to
    back; any row range of the [N, 1, T] batch InferDataset + infer_data_collate would build is produced on the device, and enhanced
    rows are scattered back into per-utterance waveforms trimmed to their input lengths (the regroup loop of reference infer.py:81-120)."""

    def __init__(self, waves: Sequence[torch.Tensor], T: int, device, lo: int = 0, hi: int = None):
        """Only the samples of the utterances that own rows [lo, hi) are uploaded (a rank of a sharded run passes its slice)."""
        self.T, self.device = int(T), torch.device(device)
        self.lengths = [int(torch.as_tensor(w).numel()) for w in waves]
        counts = chunk_counts(self.lengths, T)
        self.row_off = torch.zeros(len(waves) + 1, dtype=torch.int64)
        self.row_off[1:] = torch.cumsum(torch.tensor(counts), 0)
        self.n_rows = int(self.row_off[-1])
        hi = self.n_rows if hi is None else hi
        self.lo, self.hi = lo, hi
        owns = [(int(self.row_off[u]) < hi and int(self.row_off[u + 1]) > lo) for u in range(len(waves))]
        self.sample_off = torch.zeros(len(waves) + 1, dtype=torch.int64)
        self.sample_off[1:] = torch.cumsum(torch.tensor([n if o else 0 for n, o in zip(self.lengths, owns)]), 0)
        total = int(self.sample_off[-1])
        host = torch.empty(max(total, 1), dtype=torch.float32).pin_memory()
        for u, w in enumerate(waves):
            if owns[u]:
                host[int(self.sample_off[u]):int(self.sample_off[u + 1])] = torch.as_tensor(w, dtype=torch.float32).reshape(-1)
        self.owns = owns
        self.flat = host.to(self.device, non_blocking=True)
        self.sample_off_d = self.sample_off.to(self.device)
        self.row_off_d = self.row_off.to(self.device)

    def _call(self, fn, a, b, lo, hi):
        import ctypes as C
        from .. import _lib
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(fn(C.c_void_p(a.data_ptr()), C.c_void_p(self.sample_off_d.data_ptr()), C.c_void_p(self.row_off_d.data_ptr()),
                          len(self.lengths), self.T, int(lo), int(hi), C.c_void_p(b.data_ptr()), C.c_void_p(st)))

    def rows(self, lo: int, hi: int) -> torch.Tensor:
        """[hi - lo, 1, T] rows of the batch, built on the device."""
        from .. import _lib
        assert self.lo <= lo <= hi <= self.hi
        out = torch.empty((hi - lo, 1, self.T), device=self.device)
        self._call(_lib.lib().sddm_chunk_rows, self.flat, out, lo, hi)
        return out

    def regroup(self, rows: torch.Tensor, lo: int, hi: int, flat_out: torch.Tensor = None) -> torch.Tensor:
        """Scatter enhanced rows [lo, hi) into the flat per-utterance layout (trimmed); returns the flat buffer."""
        from .. import _lib
        if flat_out is None:
            flat_out = torch.zeros_like(self.flat)
        self._call(_lib.lib().sddm_regroup_rows, rows.contiguous(), flat_out, lo, hi)
        return flat_out

    def split(self, flat_out: torch.Tensor) -> List[torch.Tensor]:
        """Per-utterance [1, n_u] views of a flat buffer (None for utterances this batch does not own)."""
        return [flat_out[int(self.sample_off[u]):int(self.sample_off[u + 1])].reshape(1, -1) if o else None
                for u, o in enumerate(self.owns)]


def balanced_splits(n: int, limit: int) -> List[Tuple[int, int]]:
    """[0, n) cut into ceil(n / limit) near-equal pieces (a 51-row tail behind four 64-row sub-batches costs almost a full
    sub-batch of GPU time; five pieces of 61-62 rows do not)."""
    if n <= 0:
        return []
    k = max(1, ceil(n / max(1, limit)))
    base, extra = divmod(n, k)
    out, lo = [], 0
    for i in range(k):
        hi = lo + base + (1 if i < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


class InferDataset(torch.utils.data.Dataset):
    """(clean, noisy) utterance pairs -> (clean_chunks, noisy_chunks, index) exactly as the reference yields them.

    ``items`` is a sequence of (clean, noisy) waveforms (tensors / arrays / paths to .npy); ``clean`` may be None
    (then the noisy waveform stands in: enhancement needs no target)."""

    def __init__(self, items: Sequence, sample_rate: int = 16000, T: int = 16448, names: Sequence[str] = None):
        self.items, self.sample_rate, self.T = list(items), sample_rate, T
        self.names = list(names) if names is not None else ["utt%05d" % i for i in range(len(self.items))]

    @staticmethod
    def _load(x):
        if isinstance(x, (str, Path)):
            return load_wave(x)
        return torch.as_tensor(x, dtype=torch.float32).reshape(1, -1)

    def __len__(self):
        return len(self.items)

    def getName(self, idx):
        return self.names[idx]

    def __getitem__(self, index):
        clean, noisy = self.items[index]
        noisy = self._load(noisy)
        clean = noisy if clean is None else self._load(clean)
        assert clean.shape[-1] == noisy.shape[-1]
        noisy_c = chunk_waveform(noisy, self.T)
        clean_c = chunk_waveform(clean, self.T)
        return clean_c, noisy_c, index * torch.ones(noisy_c.shape[0], dtype=torch.long)


def infer_data_collate(batch) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Concatenate the chunk stacks of several utterances (reference :143-155)."""
    clean = torch.cat([b[0] for b in batch], dim=0)
    noisy = torch.cat([b[1] for b in batch], dim=0)
    index = torch.cat([b[2] for b in batch], dim=0)
    return clean, noisy, index


class InferDataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size, num_workers=0):
        super().__init__(dataset, batch_size=batch_size, shuffle=False, num_workers=num_workers, collate_fn=infer_data_collate)


def regroup(rows: torch.Tensor, index: torch.Tensor, lengths: Sequence[int] = None) -> List[torch.Tensor]:
    """Inverse of the collate: rows [N,1,T] + utterance id per row -> list of [1, n_i] waveforms
    (reference infer.py:81-120; optionally trimmed back to the un-padded lengths)."""
    out = []
    ids = index.tolist()
    start = 0
    for k in range(1, len(ids) + 1):
        if k == len(ids) or ids[k] != ids[start]:
            wav = rows[start:k].reshape(1, -1)
            if lengths is not None:
                wav = wav[:, :lengths[len(out)]]
            out.append(wav)
            start = k
    return out
