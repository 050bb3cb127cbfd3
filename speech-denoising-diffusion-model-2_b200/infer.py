"""Inference entry point — host mirror of reference infer.py:20-146.

``build_model(config)`` performs infer.py:36-51 (registry construction through ``ConfigParser.init_obj``, ``.to(device)``,
``.eval()``, optional checkpoint ``state_dict``); ``enhance_batch`` is the per-batch body infer.py:72-77 minus file IO;
``enhance_utterances`` adds the chunk / regroup steps around it (infer.py:81-120) and the multi-GPU row sharding.
``generate_from_spectrograms`` is the same for the spectrogram-conditioned models (config_diffwave.json / config_wavegrad.json).
Run as ``python -m sddm_b200.infer -c configs/config_unet.json [-r checkpoint.pth] --inputs noisy1.wav noisy2.npy ...``
(waveforms for SDDM; ``[bins, frames]`` spectrogram ``.npy`` files — or, for DiffWave, ``.wav`` files that go through the STFT
front-end — for SDDM_spectrogram).
"""
from __future__ import annotations

import argparse
from typing import List, Optional, Sequence

import torch

from .data_loader import data_loaders as module_data
from .model import diffusion as module_diffusion
from .model import model as module_arch
from .model import network as module_network
from .parse_config import ConfigParser
from .sharding import gather_rows, shard_bounds


def build_model(config, device: Optional[torch.device] = None, state_dict: Optional[dict] = None):
    device = torch.device("cuda") if device is None else torch.device(device)
    diffusion = config.init_obj("diffusion", module_diffusion, device=device)
    network = config.init_obj("network", module_network, num_samples=config["num_samples"])
    model = config.init_obj("arch", module_arch, diffusion, network)
    model = model.to(device)
    model.eval()
    if state_dict is None and getattr(config, "resume", None) is not None:
        # reference checkpoints pickle the ConfigParser object next to the weights (base_trainer.py:100-128)
        checkpoint = torch.load(config.resume, map_location=device, weights_only=False)
        state_dict = checkpoint["state_dict"]
    if state_dict is not None:
        state_dict = {k[len("module."):] if k.startswith("module.") else k: v for k, v in state_dict.items()}
        model.load_state_dict(state_dict)
    return model


@torch.no_grad()
def enhance_batch(model, condition: torch.Tensor, **kw) -> torch.Tensor:
    """infer.py:72-77: ``condition`` [B,1,L] (host or device) -> enhanced [B,1,L] on the model's device."""
    device = next(model.parameters()).device
    return model.infer(condition.to(device, non_blocking=True), **kw)


@torch.no_grad()
def _resolve_seed(seed: Optional[int], world: int) -> int:
    """seed=None: draw one from torch's CPU generator (so torch.manual_seed governs it, as it governs the reference's torch.randn
    draws) and share rank 0's value, since the Philox stream is keyed by (seed, GLOBAL row)."""
    if seed is not None:
        return int(seed)
    s = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)
    if world > 1:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            s = s.to(dev)
            dist.broadcast(s, src=0)
            s = s.cpu()
    return int(s.item())


def enhance_utterances(model, waves: Sequence[torch.Tensor], batch_chunks: int = 64, seed: Optional[int] = None,
                       rank: int = 0, world: int = 1, timings: Optional[dict] = None) -> List[torch.Tensor]:
    """Chunk every utterance to [n_i,1,L] (InferDataset semantics), enhance the rows this rank owns in sub-batches,
    gather, and regroup to one waveform per utterance trimmed to its original length.  The Philox stream is keyed
    by the GLOBAL row id, so the result does not depend on `world` or `batch_chunks`."""
    L = model.noise_estimate_model.cfg["num_samples"]
    device = next(model.parameters()).device
    seed = _resolve_seed(seed, world)
    import time
    lengths = [int(w.numel()) for w in waves]
    n = int(sum(module_data.chunk_counts(lengths, L)))
    lo, hi = shard_bounds(n, world, rank)

    def tick(name, t0):   # phase timeline for bench.py (adds device syncs; off by default)
        if timings is not None:
            torch.cuda.synchronize()
            timings[name] = timings.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)
        return time.perf_counter()

    t0 = time.perf_counter()
    # dataset edge on the device: one upload of the samples this rank owns, rows built / regrouped by sddm_chunk_rows / sddm_regroup_rows
    batch = module_data.DeviceBatch(waves, L, device, lo, hi)
    t0 = tick("upload_ms", t0)
    if world == 1:
        flat_out = torch.zeros_like(batch.flat)
        for a, b in module_data.balanced_splits(hi - lo, batch_chunks):      # near-equal sub-batches: no short tail
            out = model.infer(batch.rows(lo + a, lo + b), seed=seed, row0=lo + a)
            batch.regroup(out, lo + a, lo + b, flat_out)
        t0 = tick("chunk_enhance_regroup_ms", t0)
        return batch.split(flat_out)
    out_local = torch.empty((hi - lo, 1, L), device=device)
    for a, b in module_data.balanced_splits(hi - lo, batch_chunks):
        out_local[a:b] = model.infer(batch.rows(lo + a, lo + b), seed=seed, row0=lo + a)
    t0 = tick("chunk_enhance_ms", t0)
    full = gather_rows(out_local, n)                                         # every rank gets every enhanced row (reference: one process sees all)
    t0 = tick("gather_ms", t0)
    whole = module_data.DeviceBatch.__new__(module_data.DeviceBatch)         # index tables of the WHOLE set (no samples uploaded)
    whole.T, whole.device, whole.lengths = L, device, lengths
    counts = module_data.chunk_counts(lengths, L)
    whole.row_off = torch.zeros(len(waves) + 1, dtype=torch.int64)
    whole.row_off[1:] = torch.cumsum(torch.tensor(counts), 0)
    whole.sample_off = torch.zeros(len(waves) + 1, dtype=torch.int64)
    whole.sample_off[1:] = torch.cumsum(torch.tensor(lengths), 0)
    whole.owns = [True] * len(waves)
    whole.sample_off_d, whole.row_off_d = whole.sample_off.to(device), whole.row_off.to(device)
    flat_out = torch.zeros(int(whole.sample_off[-1]), device=device)
    whole._call(_lib_mod().sddm_regroup_rows, full, flat_out, 0, n)
    res = whole.split(flat_out)
    tick("regroup_ms", t0)
    return res


def _lib_mod():
    from . import _lib
    return _lib.lib()


@torch.no_grad()
def generate_from_spectrograms(model, specs: Sequence[torch.Tensor], batch: int = 8, seed: Optional[int] = None, rank: int = 0,
                               world: int = 1) -> List[torch.Tensor]:
    """SDDM_spectrogram.infer over a list of [bins, frames] spectrograms: utterances are sharded contiguously over the ranks
    (no communication inside the loop), equal-length neighbours are batched, Philox is keyed by the GLOBAL utterance id.
    Returns this rank's waveforms ([1, hop * frames] each) in utterance order, with None for utterances of other ranks."""
    device = next(model.parameters()).device
    seed = _resolve_seed(seed, world)
    lo, hi = shard_bounds(len(specs), world, rank)
    outs: List[Optional[torch.Tensor]] = [None] * len(specs)
    i = lo
    while i < hi:
        j = i + 1
        while j < hi and j - i < batch and specs[j].shape == specs[i].shape:
            j += 1
        x = torch.stack([torch.as_tensor(s, dtype=torch.float32) for s in specs[i:j]]).to(device, non_blocking=True)
        y = model.infer(x, seed=seed, row0=i)
        for k in range(i, j):
            outs[k] = y[k - i].reshape(1, -1)
        i = j
    return outs


def main(config, inputs: Sequence[str], out_dir: Optional[str] = None, seed: Optional[int] = None):
    import os
    import pathlib
    import numpy as np
    logger = config.get_logger("infer")
    model = build_model(config)
    logger.info(model)
    sr = config.config.get("sample_rate", 16000)
    bs = max(1, int(config.config.get("infer_data_loader", {}).get("args", {}).get("batch_size", 64)))
    if isinstance(model, module_arch.SDDM_spectrogram):
        from . import prepare_spectrogram as PS
        device = next(model.parameters()).device
        specs = []
        for f in inputs:
            if f.endswith(".npy"):
                specs.append(torch.from_numpy(np.load(f).astype("float32")))
            else:   # DiffWave: wav -> hamming STFT magnitude, log / clamp compressed (prepare_spectrogram.py:20-41)
                cfg = config.config.get("spectrogram", {"window_length": 1024, "hop_samples": 256})
                tr = PS.Spectrogram(n_fft=cfg["window_length"], hop_length=cfg["hop_samples"], window_fn=torch.hamming_window, log_clamp=True)
                specs.append(tr(module_data.load_wave(f, sr).to(device))[0].cpu())
        outs = generate_from_spectrograms(model, specs, batch=bs, seed=seed)
    else:
        waves = [module_data.load_wave(f, sr).reshape(-1) for f in inputs]
        outs = enhance_utterances(model, waves, batch_chunks=bs, seed=seed)
    out_path = (config.save_dir / "samples" / "output") if out_dir is None else pathlib.Path(out_dir)
    out_path.mkdir(parents=True, exist_ok=True)
    for f, o in zip(inputs, outs):
        stem = os.path.basename(f)
        stem = stem[:-4] if stem.endswith((".wav", ".npy")) else stem
        module_data.save_wave(out_path / (stem + (".npy" if f.endswith(".npy") and not isinstance(model, module_arch.SDDM_spectrogram) else ".wav")), o, sr)
    logger.info("processed %d utterances -> %s", len(outs), out_path)


if __name__ == "__main__":
    args = argparse.ArgumentParser(description="sddm_b200 inference")
    args.add_argument("-c", "--config", default=None, type=str, help="config file path (default: None)")
    args.add_argument("-r", "--resume", default=None, type=str, help="path to latest checkpoint (default: None)")
    args.add_argument("-d", "--device", default=None, type=str, help="indices of GPUs to enable (default: all)")
    args.add_argument("--inputs", "--npy", nargs="+", default=[], dest="inputs",
                      help="noisy utterances (.wav / .npy waveforms) or, for SDDM_spectrogram, spectrograms (.npy [bins, frames])")
    args.add_argument("--out", default=None, type=str)
    args.add_argument("--seed", default=None, type=int, help="Philox seed of the sampling noise (default: drawn from torch's generator)")
    parsed = args.parse_args()
    main(ConfigParser.from_args(parsed), parsed.inputs, parsed.out, parsed.seed)
