"""Input-feature pipeline (csrc/fft.cu): per-second log-amplitude FFT + augmentation + standardisation.

CPU: the oracle (oracle/fft_oracle.py) against the golden outputs of the reference's own computeSliceMatrix /
_random_reflect / _random_scale / StandardScaler (tests/golden/fft_features.npz, made by make_fft_golden.py).
GPU: dcgru_fft_features through the C ABI (ops.fft_features) against the same golden vectors and against the oracle on
larger seeded inputs; size-independent properties at the BASELINE config-2 batch (B=512, T=60): scale equivariance
(signal * c -> features + log c), window independence, all-zero windows -> log(1e-8), and the correlation graph built
from the raw features equal to the one built from the oracle's features.
Tolerance: max|d| / max|ref| <= 1e-4 (BASELINE.json); measured ~1e-6.
"""
import os

import numpy as np
import pytest
import torch

from oracle import fft_oracle as FO
from tests.helpers import rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fft_features.npz")
TOL = 1e-4


def _golden():
    z = np.load(GOLD)
    return {k: z[k] for k in z.files}


def _clip(sig, c, clip_len=4):
    return sig[:, c * clip_len * 200:(c + 1) * clip_len * 200]


def test_oracle_matches_reference_pipeline():
    g = _golden()
    for c in range(3):
        pairs = [tuple(p) for p in g[f"pairs{c}"]]
        for kind, mean, std in (("vec", g["mean_vec"], g["std_vec"]), ("scalar", g["mean_scalar"], g["std_scalar"])):
            x, raw = FO.features(_clip(g["signal"], c), pairs, float(g[f"scale{c}"]), mean, std)
            assert np.abs(raw - g[f"raw{c}"]).max() < 1e-11
            assert np.abs(x - g[f"x_{kind}{c}"]).max() < 1e-11
    assert np.all(g["raw0"][1, 3] == np.log(1e-8))           # the all-zero window hit the amp == 0 rule


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _dest(pairs, n=19):
    d = np.arange(n, dtype=np.int32)
    for i, j in pairs:
        d[i], d[j] = j, i
    return d


@pytest.mark.gpu
def test_kernel_matches_reference_golden(dev):
    from eeg_gnn_ssl_b200 import ops
    g = _golden()
    sig = torch.tensor(np.stack([_clip(g["signal"], c) for c in range(3)]), dtype=torch.float32, device=dev)
    dest = torch.tensor(np.stack([_dest(g[f"pairs{c}"]) for c in range(3)]), device=dev)
    ls = torch.tensor([np.log(float(g[f"scale{c}"])) for c in range(3)], dtype=torch.float32, device=dev)
    for kind, mean, std in (("vec", g["mean_vec"], g["std_vec"]), ("scalar", g["mean_scalar"], g["std_scalar"])):
        x, raw = ops.fft_features(sig, mean=torch.tensor(mean), std=torch.tensor(std), dest_channel=dest, log_scale=ls,
                                  return_raw=True)
        for c in range(3):
            assert rel_err(raw[c].cpu().numpy(), g[f"raw{c}"]) < TOL
            assert rel_err(x[c].cpu().numpy(), g[f"x_{kind}{c}"]) < TOL
    assert torch.all(raw[0, 1, 3] == float(np.float32(np.log(np.float32(1e-8)))))
    # no augmentation, no scaler: x == raw
    x0 = ops.fft_features(sig)
    assert torch.equal(x0, raw)


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,N", [(5, 12, 19), (33, 3, 19), (2, 60, 7)])
def test_kernel_matches_oracle(dev, B, T, N):
    from eeg_gnn_ssl_b200 import ops
    rng = np.random.default_rng(B * 100 + T)
    sig = rng.standard_normal((B, N, T * 200)) * rng.uniform(0.1, 50.0, (B, N, 1)) + rng.uniform(-20, 20, (B, N, 1))
    sig32 = sig.astype(np.float32)
    mean = rng.uniform(2, 5, (1, N, 1))
    std = rng.uniform(1, 2, (1, N, 1))
    pairs = [(0, 1), (2, 5)]
    scales = rng.uniform(0.8, 1.2, B)
    flags = rng.integers(0, 2, B)
    dest = np.stack([_dest(pairs if flags[b] else [], N) for b in range(B)])
    # non-contiguous batch stride on purpose (a slice of a longer recording buffer)
    buf = torch.zeros((B, N, T * 200 + 40), dtype=torch.float32, device=dev)
    buf[:, :, :T * 200] = torch.tensor(sig32, device=dev)
    x, raw = ops.fft_features(buf[:, :, :T * 200], mean=torch.tensor(mean), std=torch.tensor(std),
                              dest_channel=torch.tensor(dest, device=dev),
                              log_scale=torch.tensor(np.log(scales), dtype=torch.float32, device=dev), return_raw=True)
    worst = 0.0
    for b in range(B):
        xr, rr = FO.features(sig32[b], pairs if flags[b] else None, scales[b], mean, std)
        worst = max(worst, rel_err(x[b].cpu().numpy(), xr), rel_err(raw[b].cpu().numpy(), rr))
    assert worst < TOL, worst


@pytest.mark.gpu
def test_full_size_properties(dev):
    from eeg_gnn_ssl_b200 import ops
    B, T, N = 512, 60, 19                                   # BASELINE config 2 batch: 0.47 GB of samples
    g = torch.Generator(device=dev).manual_seed(1)
    sig = torch.randn((B, N, T * 200), generator=g, device=dev) * 30.0
    sig[7, 4, 1000:1200] = 0.0
    raw = ops.fft_features(sig)
    assert raw.shape == (B, T, N, 100) and torch.isfinite(raw).all()
    assert torch.all(raw[7, 5, 4] == float(np.float32(np.log(np.float32(1e-8)))))
    # scale equivariance: |FFT(c x)| = c |FFT(x)| (c a power of two: exact in fp32)
    raw4 = ops.fft_features(sig * 4.0)
    m = raw > -18.0
    assert float((raw4 - raw - float(np.log(4.0)))[m].abs().max()) < 2e-6 * 10
    # windows are independent: a sub-batch gives bit-identical features
    sub = ops.fft_features(sig[100:103, :, 200 * 10:200 * 20])
    assert torch.equal(sub, raw[100:103, 10:20])
    # spot check against the oracle
    for b, t in ((0, 0), (511, 59), (255, 31)):
        ref = FO.log_amplitude_windows(sig[b, :, 200 * t:200 * (t + 1)].cpu().numpy())[0]
        assert rel_err(raw[b, t].cpu().numpy(), ref) < TOL
    # the correlation graph from the device features == from the oracle's features (top-k selection is tie-sensitive on
    # white noise, so give the channels shared structure first)
    common = torch.randn((4, 1, T * 200), generator=g, device=dev) * 30.0
    w = torch.linspace(0.2, 1.5, N, device=dev).reshape(1, N, 1)
    sig2 = sig[:4] * 0.5 + common * w
    raw2 = ops.fft_features(sig2)
    s_dev = ops.corr_supports(raw2, top_k=3)
    from oracle import graph_oracle as GO
    for b in range(4):
        feats = FO.log_amplitude_windows(sig2[b].cpu().numpy())
        adj = GO.correlation_adjacency(feats, top_k=3)
        s0, s1 = GO.dual_random_walk_supports(adj)
        assert rel_err(s_dev[0][b].cpu().numpy(), s0) < TOL and rel_err(s_dev[1][b].cpu().numpy(), s1) < TOL


@pytest.mark.gpu
def test_argument_errors(dev):
    from eeg_gnn_ssl_b200 import ops
    with pytest.raises(ValueError):
        ops.fft_features(torch.zeros((2, 19, 250), device=dev))
    with pytest.raises(ValueError):
        ops.fft_features(torch.zeros((2, 19, 400), device=dev), mean=torch.zeros(3), std=torch.ones(3))
    with pytest.raises(RuntimeError):
        ops.fft_features(torch.zeros((2, 19, 400)))          # CPU tensor: no fallback
