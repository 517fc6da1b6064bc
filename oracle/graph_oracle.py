"""CPU oracle for the graph -> supports math -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement (dense 19x19 matrices instead of scipy.sparse) of

* ``scaled_laplacian``     <- /root/reference/utils.py:205-217,240-255
                              (``calculate_scaled_laplacian(adj, lambda_max=None)``:
                              symmetrise by max, L = I - D^-1/2 A D^-1/2, lambda_max by
                              ``eigsh(L, 1, 'LM')``, 2L/lambda_max - I)
* ``random_walk``          <- /root/reference/utils.py:220-230 (D^-1 A, inf -> 0)
* ``dual_random_walk_supports`` <- /root/reference/data/dataloader_detection.py:343-347
                              ([(D^-1 A)^T, (D_c^-1 A^T)^T])
* ``correlation_adjacency`` <- /root/reference/data/dataloader_detection.py:258-307 and
                              /root/reference/data/data_utils.py:174-222 (|normalised
                              zero-lag cross-correlation|, diag 1, directed top-k)
* ``diffusion_polynomials`` <- the per-sample matrices P_m with T_m = P_m Z that
                              /root/reference/model/cell.py:76-93 implies,
                              computed in float64.

Pinned by ``tests/golden/graph_*.npz`` (outputs of the reference helpers themselves, made
by ``tests/golden/make_golden.py`` in the build container).
"""
from __future__ import annotations

import numpy as np


def scaled_laplacian(adj, lambda_max=None):
    a = np.asarray(adj, dtype=np.float64)
    a = np.maximum(a, a.T)
    d = a.sum(1)
    with np.errstate(divide="ignore"):
        dis = np.power(d, -0.5)
    dis[np.isinf(dis)] = 0.0
    lap = np.eye(a.shape[0]) - (a * dis[None, :]).T * dis[None, :]
    if lambda_max is None:
        ev = np.linalg.eigvalsh((lap + lap.T) / 2)
        lambda_max = ev[np.argmax(np.abs(ev))]
    return (2.0 / lambda_max) * lap - np.eye(a.shape[0])


def random_walk(adj):
    a = np.asarray(adj, dtype=np.float64)
    d = a.sum(1)
    with np.errstate(divide="ignore"):
        dinv = np.power(d, -1.0)
    dinv[np.isinf(dinv)] = 0.0
    return dinv[:, None] * a


def dual_random_walk_supports(adj):
    """float32 supports exactly as the loader hands them to the model."""
    a = np.asarray(adj)
    return [random_walk(a).T.astype(np.float32), random_walk(a.T).T.astype(np.float32)]


def correlation_adjacency(clip, top_k=3):
    """clip: (T, N, F) raw (un-standardised) log-amplitude clip -> (N, N) float32 adjacency."""
    t, n, f = clip.shape
    x = np.transpose(clip, (1, 0, 2)).reshape(n, -1)
    adj = np.eye(n, dtype=np.float32)
    for i in range(n):
        for j in range(i + 1, n):
            xc = np.sum(x[i] * x[j])                      # 'valid' correlate of equal lengths
            cxx, cyy = np.sum(np.abs(x[i]) ** 2), np.sum(np.abs(x[j]) ** 2)
            if cxx != 0 and cyy != 0:
                xc = xc / (cxx * cyy) ** 0.5
            adj[i, j] = xc
            adj[j, i] = xc
    adj = np.abs(adj)
    no_self = adj.copy()
    np.fill_diagonal(no_self, 0)
    idx = (-no_self).argsort(axis=-1)[:, :top_k]
    mask = np.eye(n, dtype=bool)
    for i in range(n):
        mask[i, idx[i]] = True
    return (mask * adj).astype(np.float32)


def diffusion_polynomials(supports, max_diffusion_step):
    """supports: list of S arrays (N,N) -> (M-1, N, N) float64 with term_m = P_m @ Z, m=1..M-1.

    Runs the reference recurrence on the identity, so the carried-x0 quirk is
    reproduced by construction.
    """
    n = supports[0].shape[-1]
    eye = np.eye(n)
    out = []
    x0 = eye
    if max_diffusion_step == 0:
        return np.zeros((0, n, n))
    for s in supports:
        s = np.asarray(s, dtype=np.float64)
        x1 = s @ x0
        out.append(x1)
        for _ in range(2, max_diffusion_step + 1):
            x2 = 2 * (s @ x1) - x0
            out.append(x2)
            x1, x0 = x2, x1
    return np.stack(out, 0)
