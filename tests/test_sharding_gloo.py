"""N>1 host path on CPU: two gloo ranks shard the chunk rows, 'enhance' their slice, and all-gather the result."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, ret):
    import sys
    sys.path.insert(0, ROOT)
    from sddm_b200.sharding import gather_rows, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = torch.arange(n_rows * 8, dtype=torch.float32).reshape(n_rows, 1, 8)     # every rank builds the same batch
    lo, hi = shard_bounds(n_rows, world, rank)
    local = rows[lo:hi] * 2 + 1                                                    # stand-in for the per-row enhancement
    full = gather_rows(local, n_rows)
    ok = torch.equal(full, rows * 2 + 1)
    t = torch.tensor([float(hi - lo)])
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                                       # slices cover all rows exactly once
    ret[rank] = bool(ok) and int(t.item()) == n_rows
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    for n_rows in (7, 8):
        port = _free_port()
        with mp.Manager() as mgr:
            ret = mgr.dict()
            mp.spawn(_worker, args=(2, port, n_rows, ret), nprocs=2, join=True)
            assert dict(ret) == {0: True, 1: True}


def _worker_utterances(rank, world, port, ret):
    """The N>1 host path of infer.enhance_utterances on CPU: each rank builds ONLY the rows it owns from the utterance list
    (rows_of_range), processes them in balanced sub-batches, all-gathers, and regroups per utterance."""
    import sys
    sys.path.insert(0, ROOT)
    from sddm_b200.data_loader import data_loaders as D
    from sddm_b200.sharding import gather_rows, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = 16
    g = torch.Generator().manual_seed(3)
    waves = [torch.randn(n, generator=g) for n in (40, 16, 3, 70, 17, 1)]
    lengths = [int(w.numel()) for w in waves]
    counts = D.chunk_counts(lengths, T)
    n = sum(counts)
    index = torch.repeat_interleave(torch.arange(len(waves)), torch.tensor(counts))
    lo, hi = shard_bounds(n, world, rank)
    local = torch.empty(hi - lo, 1, T)
    for a, b in D.balanced_splits(hi - lo, 3):
        local[a:b] = D.rows_of_range(waves, T, lo + a, lo + b) * 2 + 1            # stand-in for the per-row enhancement
    full = gather_rows(local, n)
    outs = D.regroup(full, index, lengths)
    ok = all(torch.equal(o.reshape(-1), w * 2 + 1) for o, w in zip(outs, waves)) and len(outs) == len(waves)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_utterance_path():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_utterances, args=(2, port, ret), nprocs=2, join=True)
        assert dict(ret) == {0: True, 1: True}
