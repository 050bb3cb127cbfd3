"""Inference entry point — host mirror of reference infer.py:20-146.

``build_model(config)`` performs infer.py:36-51 (registry construction through ``ConfigParser.init_obj``, ``.to(device)``,
``.eval()``, optional checkpoint ``state_dict``); ``enhance_batch`` is the per-batch body infer.py:72-77 minus file IO;
``enhance_utterances`` adds the chunk / regroup steps around it (infer.py:81-120) and the multi-GPU row sharding.
Run as ``python -m sddm_b200.infer -c config_unet.json [-r checkpoint.pth] --npy noisy1.npy noisy2.npy ...``.
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
def enhance_utterances(model, waves: Sequence[torch.Tensor], batch_chunks: int = 64, seed: int = 0,
                       rank: int = 0, world: int = 1) -> List[torch.Tensor]:
    """Chunk every utterance to [n_i,1,L] (InferDataset semantics), enhance the rows this rank owns in sub-batches,
    gather, and regroup to one waveform per utterance trimmed to its original length.  The Philox stream is keyed
    by the GLOBAL row id, so the result does not depend on `world` or `batch_chunks`."""
    L = model.noise_estimate_model.cfg["num_samples"]
    device = next(model.parameters()).device
    ds = module_data.InferDataset([(None, w) for w in waves], T=L)
    _, rows, index = module_data.infer_data_collate([ds[i] for i in range(len(ds))])
    n = rows.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    out_local = torch.empty((hi - lo, 1, L), device=device)
    for s in range(lo, hi, batch_chunks):
        e = min(hi, s + batch_chunks)
        out_local[s - lo:e - lo] = model.infer(rows[s:e].to(device, non_blocking=True), seed=seed, row0=s)
    full = gather_rows(out_local, n) if world > 1 else out_local
    return module_data.regroup(full, index, [int(w.numel()) for w in waves])


def main(config, npy_files: Sequence[str], out_dir: Optional[str] = None):
    import numpy as np
    logger = config.get_logger("infer")
    model = build_model(config)
    logger.info(model)
    waves = [torch.from_numpy(np.load(f).astype("float32")).reshape(-1) for f in npy_files]
    bs = config.config.get("infer_data_loader", {}).get("args", {}).get("batch_size", 64)
    outs = enhance_utterances(model, waves, batch_chunks=max(1, int(bs)))
    out_path = (config.save_dir / "samples" / "output") if out_dir is None else __import__("pathlib").Path(out_dir)
    out_path.mkdir(parents=True, exist_ok=True)
    for f, o in zip(npy_files, outs):
        np.save(out_path / (__import__("os").path.basename(f)), o.cpu().numpy())
    logger.info("enhanced %d utterances -> %s", len(outs), out_path)


if __name__ == "__main__":
    args = argparse.ArgumentParser(description="sddm_b200 inference")
    args.add_argument("-c", "--config", default=None, type=str, help="config file path (default: None)")
    args.add_argument("-r", "--resume", default=None, type=str, help="path to latest checkpoint (default: None)")
    args.add_argument("-d", "--device", default=None, type=str, help="indices of GPUs to enable (default: all)")
    args.add_argument("--npy", nargs="+", default=[], help="noisy utterances as .npy float32 waveforms")
    args.add_argument("--out", default=None, type=str)
    parsed = args.parse_args()
    main(ConfigParser.from_args(parsed), parsed.npy, parsed.out)
