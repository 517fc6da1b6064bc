"""CPU oracle for the input-feature pipeline -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy (float64) restatement of what the reference's DataLoader workers do to one clip:

* ``log_amplitude_windows``  <- /root/reference/data/dataloader_detection.py:58-72 (``computeSliceMatrix``: the clip
                                (channels, clip_len*200) is cut into 1-second windows, ``is_fft=True``) around
                                /root/reference/data/data_utils.py:13-34 (``computeFFT``: ``fft(signals, n=200)``, the
                                first ``floor(n/2)`` bins, ``abs``, exact zeros -> 1e-8, ``log``); result (T, N, 100)
* ``random_reflect``          <- /root/reference/data/dataloader_detection.py:233-245 (swap the channel pairs)
* ``random_scale``            <- /root/reference/data/dataloader_detection.py:247-256 (``+= log(scale_factor)`` with FFT)
* ``standardize``             <- /root/reference/utils.py:402-403
* ``features``                <- the order ``__getitem__`` applies them in, data/dataloader_detection.py:382-393

``scipy.fftpack.fft`` (pinned scipy 1.2.1) and ``numpy.fft.fft`` are both double-precision FFTs (agreement 1e-15); the
restatement uses numpy's.  Pinned by ``tests/golden/fft_features.npz``: outputs of the reference's own ``computeFFT`` /
``StandardScaler`` made by ``tests/golden/make_fft_golden.py`` in the build container.
"""
from __future__ import annotations

import numpy as np

WINDOW = 200          # FREQUENCY (constants.py) * time_step_size (args.py default 1)


def log_amplitude_windows(clip, window=WINDOW):
    """clip (N, T*window) -> (T, N, window//2) float64."""
    clip = np.asarray(clip, dtype=np.float64)
    n, s = clip.shape
    t = s // window
    w = clip[:, :t * window].reshape(n, t, window)
    f = np.fft.fft(w, n=window, axis=-1)[..., :window // 2]
    amp = np.abs(f)
    amp[amp == 0.0] = 1e-8
    return np.log(amp).transpose(1, 0, 2)


def random_reflect(feat, swap_pairs):
    """feat (T, N, F); swap_pairs list of (i, j) or None."""
    out = feat.copy()
    if swap_pairs:
        for i, j in swap_pairs:
            out[:, [i, j], :] = feat[:, [j, i], :]
    return out


def random_scale(feat, scale_factor):
    return feat + np.log(scale_factor)


def standardize(feat, mean, std):
    return (feat - mean) / std


def features(clip, swap_pairs=None, scale_factor=None, mean=None, std=None, window=WINDOW):
    """-> (x, raw): the model input and the un-augmented features the correlation graph is computed from."""
    raw = log_amplitude_windows(clip, window)
    x = random_reflect(raw, swap_pairs)
    if scale_factor is not None:
        x = random_scale(x, scale_factor)
    if mean is not None:
        x = standardize(x, mean, std)
    return x, raw
