"""Host-side mirror of the reference's config / registry / model API (no GPU needed)."""
import argparse
import collections
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import L, ROOT, UNET_CFG


def sha(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


def test_config_parser_builds_the_reference_objects(tmp_path, monkeypatch):
    from sddm_b200.model import diffusion as module_diffusion, model as module_arch, network as module_network
    from sddm_b200.parse_config import ConfigParser, read_json
    cfg = read_json(os.path.join(ROOT, "configs", "config_unet.json"))
    cfg["trainer"]["save_dir"] = str(tmp_path)
    config = ConfigParser(cfg, run_id="t0")
    assert (tmp_path / "SDDM2_UNet" / "t0" / "config.json").exists()           # run dir + config copy
    assert config["num_samples"] == 16448 and config.save_dir == config.log_dir
    diffusion = config.init_obj("diffusion", module_diffusion, device="cpu")
    network = config.init_obj("network", module_network, num_samples=config["num_samples"])
    model = config.init_obj("arch", module_arch, diffusion, network)
    assert type(model).__name__ == "SDDM" and model.num_timesteps == 100 and model.p_transition == "condition_in"
    with pytest.raises(AssertionError, match="Overwriting kwargs"):
        config.init_obj("diffusion", module_diffusion, n_timestep=5)
    ftn = config.init_ftn("diffusion", module_diffusion, device="cpu")
    assert ftn().num_timesteps == 100
    assert config.get_logger("x", 1).level == 20
    with pytest.raises(AssertionError):
        config.get_logger("x", 5)
    assert "Trainable parameters: 5229793" in str(model)


def test_config_parser_from_args_and_overrides(tmp_path):
    from sddm_b200.parse_config import ConfigParser, read_json
    cfg = read_json(os.path.join(ROOT, "configs", "config_unet.json"))
    cfg["trainer"]["save_dir"] = str(tmp_path)
    path = tmp_path / "c.json"
    path.write_text(json.dumps(cfg))
    Opt = collections.namedtuple("CustomArgs", "flags type target")
    args = argparse.ArgumentParser()
    args.add_argument("-c", "--config", default=None, type=str)
    args.add_argument("-r", "--resume", default=None, type=str)
    args.add_argument("-d", "--device", default=None, type=str)
    options = [Opt(["--bs", "--batch_size"], int, "infer_data_loader;args;batch_size")]
    for o in options:
        args.add_argument(*o.flags, default=None, type=o.type)
    ns = args.parse_args(["-c", str(path), "--bs", "8"])
    config = ConfigParser.from_args(ns, options=[])
    assert config.resume is None
    config2 = ConfigParser(read_json(path), modification={"infer_data_loader;args;batch_size": 8}, run_id="m")
    assert config2["infer_data_loader"]["args"]["batch_size"] == 8
    with pytest.raises(AssertionError, match="Configuration file"):
        ConfigParser.from_args(args.parse_args([]))


def test_state_dict_layout_and_init_match_reference(meta):
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM
    from sddm_b200.model.network import UNetModified2
    torch.manual_seed(0)
    d = GaussianDiffusion("linear", 100, 1e-6, 1e-3, device="cpu")
    net = UNetModified2(**UNET_CFG)
    sd = SDDM(d, net, p_transition="condition_in").state_dict()
    ref = meta["weights_seed0"]
    assert list(sd) == list(ref)                                 # same keys in the same order (232 tensors)
    for k, v in sd.items():
        assert list(v.shape) == ref[k]["shape"] and sha(v) == ref[k]["sha256"], k
    assert sha(net.noise_level_mlp[0].embedding_vector) == meta["pe_vector_sha256"]
    # a reference checkpoint's state_dict loads (incl. DataParallel 'module.' prefixes handled by build_model)
    m2 = SDDM(GaussianDiffusion("linear", 100, 1e-6, 1e-3, device="cpu"), UNetModified2(**UNET_CFG))
    m2.load_state_dict(sd)


@pytest.mark.parametrize("tag,args", [("linear100", ("linear", 100, 1e-6, 1e-3)), ("quad50", ("quad", 50, 1e-4, 2e-2)),
                                      ("cosine20", ("cosine", 20, 1e-4, 2e-2))])
def test_diffusion_buffers_bit_exact(golden, tag, args):
    from sddm_b200.model.diffusion import GaussianDiffusion
    g = golden("schedules.npz")
    d = GaussianDiffusion(*args, device="cpu")
    for name in ("betas", "alphas", "alpha_bar", "sqrt_alpha_bar", "predicted_noise_coeff", "sigma", "supportive_gamma",
                 "supportive_sigma_hat", "m", "sqrt_delta", "c_xt", "c_yt", "c_epst", "sqrt_delta_estimated"):
        a, b = getattr(d, name), g[f"{tag}.{name}"]
        assert torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)), name
    k8 = d.step_scalars(args[1], "original")
    assert k8[1] == float(np.sqrt(np.float32(d.alphas[args[1]].item())))


def test_reference_error_behaviour():
    from sddm_b200.model.diffusion import GaussianDiffusion
    from sddm_b200.model.model import SDDM
    from sddm_b200.model.network import UNetModified2
    with pytest.raises(NotImplementedError):
        GaussianDiffusion("warmup10", 10, device="cpu")              # diffusion.py:83-84
    d = GaussianDiffusion("linear", 10, device="cpu")
    net = UNetModified2(**UNET_CFG)
    for kw in (dict(noise_condition="foo"), dict(p_transition="foo"), dict(q_transition="foo")):
        with pytest.raises(NotImplementedError):                     # model.py:17-26
            SDDM(d, net, **kw)
    with pytest.raises(AssertionError):
        UNetModified2(num_samples=16449)                             # UNetModified2.py:13
    with pytest.raises(RuntimeError, match="no CPU fallback"):           # the training-step forward runs on the GPU only
        SDDM(d, net).forward(torch.zeros(1, 1, L), torch.zeros(1, 1, L))


def test_chunking_collate_regroup_roundtrip():
    from sddm_b200.data_loader.data_loaders import InferDataset, chunk_waveform, infer_data_collate, regroup
    g = torch.Generator().manual_seed(0)
    waves = [torch.randn(n, generator=g) for n in (32000, 16448, 5, 16449)]
    ds = InferDataset([(None, w) for w in waves], T=L)
    assert [ds[i][1].shape[0] for i in range(4)] == [2, 1, 1, 2]       # ceil(n / 16448) chunks each
    assert chunk_waveform(waves[0], L)[1, 0, 32000 - L:].abs().sum() == 0    # zero padded tail
    clean, noisy, index = infer_data_collate([ds[i] for i in range(4)])
    assert noisy.shape == (6, 1, L) and index.tolist() == [0, 0, 1, 2, 3, 3]
    back = regroup(noisy, index, [w.numel() for w in waves])
    for w, b in zip(waves, back):
        assert torch.equal(w.reshape(1, -1), b)
    assert ds.getName(2) == "utt00002"


def test_shard_bounds_cover_rows_exactly():
    from sddm_b200.sharding import all_bounds, shard_bounds
    for n in (0, 1, 7, 64, 2301):
        for world in (1, 2, 3, 8):
            b = all_bounds(n, world)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


@pytest.mark.parametrize("name,arch_hop,net_cls", [("config_diffwave.json", 256, "DiffWave"), ("config_wavegrad.json", 300, "WaveGrad")])
def test_spectrogram_model_configs_build(tmp_path, monkeypatch, name, arch_hop, net_cls):
    """config_diffwave.json / config_wavegrad.json stay loadable through the reference's registry path (infer.py:36-40)."""
    from sddm_b200.model import diffusion as module_diffusion
    from sddm_b200.model import model as module_arch
    from sddm_b200.model import network as module_network
    from sddm_b200.parse_config import ConfigParser
    import json
    monkeypatch.chdir(tmp_path)
    with open(os.path.join(ROOT, "configs", name)) as f:
        cfg = json.load(f)
    cfg["network"]["args"] = dict(cfg["network"]["args"])
    if net_cls == "DiffWave":
        cfg["network"]["args"].update(residual_layers=2, dilation_cycle_length=2)      # keep the CPU test light
    config = ConfigParser(cfg)
    diffusion = config.init_obj("diffusion", module_diffusion, device="cpu")
    network = config.init_obj("network", module_network, num_samples=config["num_samples"])
    model = config.init_obj("arch", module_arch, diffusion, network)
    assert type(network).__name__ == net_cls and isinstance(model, module_arch.SDDM_spectrogram)
    assert model.hop_samples == arch_hop and model.num_timesteps == cfg["diffusion"]["args"]["n_timestep"]
    with pytest.raises(RuntimeError):
        model.infer(torch.zeros(1, 513 if net_cls == "DiffWave" else 128, 2))        # CPU tensors: no fallback


def test_wave_io_roundtrip(tmp_path):
    from sddm_b200.data_loader import data_loaders as D
    x = (0.3 * torch.randn(1, 5000, generator=torch.Generator().manual_seed(0))).clamp(-1, 1)
    D.save_wave(tmp_path / "a.wav", x, 16000)
    y = D.load_wave(tmp_path / "a.wav", 16000)
    assert y.shape == x.shape and float((x - y).abs().max()) <= 1.0 / 32768 + 1e-7     # 16-bit PCM quantisation
    D.save_wave(tmp_path / "a.npy", x)
    assert torch.equal(D.load_wave(tmp_path / "a.npy"), x)
    with pytest.raises(ValueError):
        D.load_wave(tmp_path / "a.wav", 8000)
    ds = D.InferDataset([(None, str(tmp_path / "a.wav"))], T=2048)
    clean, noisy, idx = ds[0]
    assert noisy.shape == (3, 1, 2048) and torch.equal(clean, noisy) and idx.tolist() == [0, 0, 0]


def test_dataset_edge_vs_reference_golden():
    """Chunking / collation / regrouping against outputs of the reference's own InferDataset + infer_data_collate
    and the regroup loop of infer.py:81-120 (tests/golden/make_golden_dataset.py), plus the sharded variants built on them."""
    import numpy as np
    from conftest import GOLDEN
    from sddm_b200.data_loader import data_loaders as D
    g = np.load(os.path.join(GOLDEN, "dataset.npz"))
    T = int(g["T"])
    order = [str(n).split(".")[0] for n in g["inventory"]]                     # the order the reference's glob produced
    waves = [torch.from_numpy(g["wave." + n]) for n in order]
    ds = D.InferDataset([(0.5 * w, w) for w in waves], T=T, names=order)
    clean, noisy, index = D.infer_data_collate([ds[i] for i in range(len(ds))])
    assert torch.equal(noisy, torch.from_numpy(g["noisy"])) and torch.equal(clean, torch.from_numpy(g["clean"]))
    assert index.tolist() == g["index"].tolist()
    assert [ds.getName(i) for i in range(len(ds))] == [str(n) for n in g["names"]]
    # regrouping: every file, untrimmed, equals the reference's reshape(1, -1) of its rows
    files = D.regroup(noisy, index)
    assert len(files) == len(waves)
    for k, f in enumerate(files):
        assert int(g["file%d.index" % k]) == k and torch.equal(f, torch.from_numpy(g["file%d.signal" % k]))
    assert int(g["n_flushed_by_reference_loop"]) == len(waves) - 1               # the reference loop never flushes the last file of a batch
    # trimmed regrouping returns the original signals
    lengths = [int(w.numel()) for w in waves]
    for f, w in zip(D.regroup(noisy, index, lengths), waves):
        assert torch.equal(f, w.reshape(1, -1))
    # per-rank row ranges (no full-dataset chunking) and balanced sub-batches
    n = noisy.shape[0]
    assert D.chunk_counts(lengths, T) == [int((index == i).sum()) for i in range(len(waves))]
    for lo, hi in ((0, n), (1, 5), (4, 5), (6, n), (0, 1)):
        assert torch.equal(D.rows_of_range([w.reshape(-1) for w in waves], T, lo, hi), noisy[lo:hi])
    assert D.balanced_splits(307, 64) == [(0, 62), (62, 124), (124, 185), (185, 246), (246, 307)]
    assert D.balanced_splits(64, 64) == [(0, 64)] and D.balanced_splits(0, 64) == [] and D.balanced_splits(65, 64) == [(0, 33), (33, 65)]
