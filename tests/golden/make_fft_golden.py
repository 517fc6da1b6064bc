"""Golden fixture for the input-feature pipeline from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_fft_golden.py

Drives the reference's own ``computeSliceMatrix(is_fft=True)`` (data/dataloader_detection.py:30-85, which calls
``data_utils.computeFFT``), ``SeizureDataset._random_reflect`` / ``_random_scale`` (:233-256) and
``utils.StandardScaler.transform`` on synthetic resampled signals.  Only the I/O around them is stubbed: ``h5py.File``
returns the synthetic array and ``getSeizureTimes`` returns no seizures (there is no TUSZ data here).
Stores signals, the drawn augmentation (swap pairs, scale factor) and the float64 outputs in ``fft_features.npz``.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

sys.path.insert(0, REF)
for n in ["h5py", "pyedflib", "matplotlib", "matplotlib.cm", "tensorboardX", "dotted_dict"]:
    sys.modules.setdefault(n, types.ModuleType(n))
sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]

_SIGNALS = {}


class _FakeH5:
    def __init__(self, fn, mode="r"):
        self.d = {"resampled_signal": _SIGNALS[fn], "resample_freq": np.int64(200)}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __getitem__(self, k):
        v = self.d[k]

        class _Dset:
            def __getitem__(self, idx):
                return v
        return _Dset()


sys.modules["h5py"].File = _FakeH5

import data.dataloader_detection as dd      # noqa: E402
import utils as ru                          # noqa: E402

dd.getSeizureTimes = lambda fn: []


def synth(rng, n, length):
    """1/f-like noise + a few sinusoids + offset: spectra spanning ~4 decades, like resampled EEG in microvolts."""
    t = np.arange(length) / 200.0
    white = rng.standard_normal((n, length))
    f = np.fft.rfftfreq(length, 1 / 200.0)
    shape = 1.0 / np.maximum(f, 0.5) ** 1.0
    x = np.fft.irfft(np.fft.rfft(white, axis=-1) * shape, n=length, axis=-1)
    x = 40.0 * x / x.std()
    for k in range(n):
        x[k] += 15.0 * np.sin(2 * np.pi * (8.0 + 0.37 * k) * t + k) + 3.0 * np.sin(2 * np.pi * 60.0 * t)
    return x + rng.uniform(-5, 5, (n, 1))


def main():
    rng = np.random.default_rng(20240)
    n, clip_len, nclip = 19, 4, 3
    out = {}
    sig = synth(rng, n, clip_len * 200 * nclip)
    sig[3, 200:400] = 0.0                      # an all-zero window: the amp == 0 -> 1e-8 rule (zero-padded recordings)
    _SIGNALS["rec"] = sig
    out["signal"] = sig
    mean = rng.uniform(3.0, 5.0, (1, n, 1))
    std = rng.uniform(1.0, 2.0, (1, n, 1))
    out["mean_vec"], out["std_vec"] = mean, std
    out["mean_scalar"], out["std_scalar"] = np.float64(3.924), np.float64(1.560)

    class Dummy:
        use_fft = True

    np.random.seed(7)
    for c in range(nclip):
        clip, _ = dd.computeSliceMatrix(h5_fn="rec", edf_fn="rec.edf", clip_idx=c, time_step_size=1, clip_len=clip_len,
                                        is_fft=True)
        out[f"raw{c}"] = clip                                         # (T, N, 100) float64
        refl, pairs = dd.SeizureDataset._random_reflect(Dummy(), clip)
        scaled = dd.SeizureDataset._random_scale(Dummy(), refl.copy())
        out[f"pairs{c}"] = np.asarray(pairs if pairs is not None else [], dtype=np.int64).reshape(-1, 2)
        out[f"scale{c}"] = np.exp((scaled - refl).mean())
        out[f"x_vec{c}"] = ru.StandardScaler(mean, std).transform(scaled)
        out[f"x_scalar{c}"] = ru.StandardScaler(out["mean_scalar"], out["std_scalar"]).transform(scaled)
    np.savez_compressed(os.path.join(HERE, "fft_features.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
