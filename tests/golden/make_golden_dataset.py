"""Golden vectors for the dataset edge of the path, produced by the REAL reference classes (build container only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_dataset.py

Pinned (reference file:line):
  * chunking ........ InferDataset.__getitem__ ('.logwav.npy' branch: np.load -> pad to a multiple of T -> view [n_chunk, 1, T])   data_loader/data_loaders.py:101-140
  * collation ....... infer_data_collate                                                                                          data_loader/data_loaders.py:143-155
  * regrouping ...... the per-file loop of infer.py:81-120 (rows of one index -> reshape(1, -1)), restated here on the reference's own collated batch
The waveforms are written to a temporary data_root/{clean,noisy}/*.logwav.npy; the golden stores them, the inventory order the
reference saw (glob order), the collated rows / index, and the regrouped per-file signals.
"""
import os
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import data_loader.data_loaders as ref_data  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
T = 64
LENGTHS = (150, 64, 7, 200, 65, 128)


def main():
    g = torch.Generator().manual_seed(31)
    waves = {("utt%02d" % i): (0.1 * torch.randn(1, n, generator=g)).numpy().astype(np.float32) for i, n in enumerate(LENGTHS)}
    with tempfile.TemporaryDirectory() as root:
        for sub in ("clean", "noisy"):
            os.makedirs(os.path.join(root, sub))
            for name, w in waves.items():
                np.save(os.path.join(root, sub, name + ".logwav.npy"), w if sub == "noisy" else 0.5 * w)
        ds = ref_data.InferDataset(root, ".logwav.npy", sample_rate=16000, T=T)
        inventory = list(ds.inventory)
        clean, noisy, index = ref_data.infer_data_collate([ds[i] for i in range(len(ds))])
        names = [ds.getName(i) for i in range(len(ds))]
    # the regroup loop of infer.py:81-120 on the collated batch (files come out when the index changes; the reference drops the LAST
    # file of a batch because it only flushes on a change - restated faithfully below, and the complete regrouping next to it)
    files, cur, prev = [], [], -1
    for b in range(noisy.shape[0]):
        ind = int(index[b])
        if ind == prev:
            cur.append(b)
            continue
        if prev > -1:
            files.append((prev, noisy[cur].reshape(1, -1).numpy()))
        cur, prev = [b], ind
    last = (prev, noisy[cur].reshape(1, -1).numpy())
    out = {"T": np.asarray(T), "inventory": np.asarray(inventory), "names": np.asarray(names), "noisy": noisy.numpy(), "clean": clean.numpy(),
           "index": index.numpy(), "n_flushed_by_reference_loop": np.asarray(len(files))}
    for k, (ind, sig) in enumerate(files + [last]):
        out["file%d.index" % k] = np.asarray(ind)
        out["file%d.signal" % k] = sig
    for name, w in waves.items():
        out["wave." + name] = w
    path = os.path.join(OUT, "dataset.npz")
    np.savez_compressed(path, **out)
    print("inventory", inventory, "rows", tuple(noisy.shape), "index", index.tolist(), "->", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
