"""CPU oracle for the DCGRU hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file.  The product path
(``eeg-gnn-ssl_b200/``) never does: it fails loudly when the CUDA library is
missing instead of falling back to anything in here.

What this is
------------
A functional restatement, in plain torch CPU ops (fp32 by default, fp64 when the
tensors passed in are fp64), of the reference's diffusion-convolutional GRU:

* ``diffusion_terms``   <- /root/reference/model/cell.py:76-93   (hop recurrence,
                           including the carried ``x0`` across supports)
* ``dconv``             <- /root/reference/model/cell.py:66-118  (column order c*M+m)
* ``cell_forward``      <- /root/reference/model/cell.py:182-210
* ``encoder_forward``   <- /root/reference/model/model.py:81-102
* ``decoder_forward``   <- /root/reference/model/model.py:149-204
* ``classification_head`` <- /root/reference/model/model.py:257-270,
                             /root/reference/utils.py:346-357
* ``masked_mae``        <- /root/reference/utils.py:431-442

Backward is torch autograd over these ops, which is exactly how the reference
obtains its gradients (it has no hand-written backward).

Pinning
-------
The reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the reference itself: ``tests/golden/make_golden.py``
imports the unmodified reference modules from /root/reference in the build
container, runs them on seeded inputs and stores inputs, weights, outputs and all
parameter gradients under ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py``
replays every fixture through this file.  Supports/graph math is in
``oracle/graph_oracle.py``.
"""
from __future__ import annotations

import torch


# ----------------------------------------------------------------------------------------
# diffusion graph convolution
# ----------------------------------------------------------------------------------------
def diffusion_terms(supports, z, max_diffusion_step):
    """All M = S*K+1 diffusion terms of ``z`` (B, N, C) as a list.

    Follows model/cell.py:76-93 literally: ``x0``/``x1`` are NOT reset between supports,
    so the second support's recurrence starts from what the first one left behind.
    """
    terms = [z]
    if max_diffusion_step == 0:
        return terms
    x0 = z
    for s in supports:
        x1 = torch.matmul(s, x0)
        terms.append(x1)
        for _ in range(2, max_diffusion_step + 1):
            x2 = 2 * torch.matmul(s, x1) - x0
            terms.append(x2)
            x1, x0 = x2, x1
    return terms


def dconv(supports, x, h, weight, biases, max_diffusion_step, num_nodes):
    """``[x | h]`` -> diffusion -> projection.  x: (B, N*Fin), h: (B, N*H) -> (B, N*out).

    The projection input has column index ``c * M + m`` (model/cell.py:98-114): stacking the
    terms on a trailing axis and flattening (C, M) reproduces that order.
    """
    b = x.shape[0]
    z = torch.cat([x.reshape(b, num_nodes, -1), h.reshape(b, num_nodes, -1)], dim=2)
    g = torch.stack(diffusion_terms(supports, z, max_diffusion_step), dim=3)  # (B,N,C,M)
    out = g.reshape(b * num_nodes, -1) @ weight + biases
    return out.reshape(b, -1)


def cell_forward(supports, x, h, p, max_diffusion_step, num_nodes, activation="tanh"):
    """One DCGRU step (model/cell.py:182-210).  ``p`` = dict(Wg, bg, Wc, bc).

    r = first H columns of each node's gate block, u = last H (cell.py:198-201).
    ``activation``: 'tanh' -> tanh, anything else -> relu (cell.py:146).
    """
    b = x.shape[0]
    hid = p["Wc"].shape[1]
    gate = torch.sigmoid(dconv(supports, x, h, p["Wg"], p["bg"], max_diffusion_step, num_nodes))
    gate = gate.reshape(b, num_nodes, 2 * hid)
    r = gate[..., :hid].reshape(b, -1)
    u = gate[..., hid:].reshape(b, -1)
    c = dconv(supports, x, r * h, p["Wc"], p["bc"], max_diffusion_step, num_nodes)
    c = torch.tanh(c) if activation == "tanh" else torch.relu(c)
    return u * h + (1 - u) * c


# ----------------------------------------------------------------------------------------
# sequence drivers
# ----------------------------------------------------------------------------------------
def encoder_forward(x_seq, h0, supports, layers, max_diffusion_step, num_nodes,
                    activation="tanh"):
    """Layer-outer / time-inner encoder (model/model.py:81-102).

    x_seq: (T, B, N, Fin) or (T, B, N*Fin); h0: (L, B, N*H); layers: list of param dicts.
    Returns (output_hidden (L,B,N*H), top-layer sequence (T,B,N*H)).
    """
    t_len, b = x_seq.shape[0], x_seq.shape[1]
    cur = x_seq.reshape(t_len, b, -1)
    last = []
    for l, p in enumerate(layers):
        h = h0[l]
        outs = []
        for t in range(t_len):
            h = cell_forward(supports, cur[t], h, p, max_diffusion_step, num_nodes, activation)
            outs.append(h)
        last.append(h)
        cur = torch.stack(outs, dim=0)
    return torch.stack(last, dim=0), cur


def decoder_forward(targets, h0, supports, layers, proj_w, proj_b, max_diffusion_step,
                    num_nodes, activation="tanh", teacher_force=None, dropout_masks=None):
    """Time-outer / layer-inner autoregressive decoder (model/model.py:149-204).

    targets: (To, B, N, Fo); h0: (L, B, N*H); ``layers[l]`` may be the same dict object for
    l >= 1 (weight tying, model/model.py:126,142-143).  ``teacher_force``: optional list of To
    bools -- the outcome of the reference's per-step ``random.random() < ratio`` draw
    (model/model.py:198-202); ``dropout_masks``: optional (To, B, N, H) multiplicative masks
    standing for ``nn.Dropout`` before the projection (model/model.py:192-193).
    Returns (To, B, N*Fo).
    """
    t_len, b = targets.shape[0], targets.shape[1]
    tgt = targets.reshape(t_len, b, -1)
    hid = proj_w.shape[1]
    cur = torch.zeros(b, tgt.shape[2], dtype=tgt.dtype)
    hs = [h0[l] for l in range(len(layers))]
    outs = []
    for t in range(t_len):
        for l, p in enumerate(layers):
            hs[l] = cell_forward(supports, cur, hs[l], p, max_diffusion_step, num_nodes,
                                 activation)
            cur = hs[l]
        top = cur.reshape(b, num_nodes, hid)
        if dropout_masks is not None:
            top = top * dropout_masks[t]
        y = (top @ proj_w.t() + proj_b).reshape(b, -1)
        outs.append(y)
        cur = tgt[t] if (teacher_force is not None and teacher_force[t]) else y
    return torch.stack(outs, dim=0)


# ----------------------------------------------------------------------------------------
# callers either side of the path -- needed to seed realistic gradients
# ----------------------------------------------------------------------------------------
def classification_head(top_seq, seq_lengths, fc_w, fc_b, num_nodes):
    """last-relevant gather -> ReLU -> FC -> max over nodes (model/model.py:257-270,
    utils.py:346-357).  Dropout p=0 (the reference default, args.py:148)."""
    t_len, b, _ = top_seq.shape
    idx = (seq_lengths.long() - 1).clamp(min=0)
    last = top_seq[idx, torch.arange(b)]                       # (B, N*H)
    last = last.reshape(b, num_nodes, -1)
    logits = torch.relu(last) @ fc_w.t() + fc_b                # (B, N, classes)
    return logits.max(dim=1).values


def masked_mae(pred, true, mask_val=0.0):
    """utils.py:431-442."""
    m = (true != mask_val).to(pred.dtype)
    m = m / m.mean()
    loss = (pred - true).abs() * m
    loss = torch.where(torch.isnan(loss), torch.zeros_like(loss), loss)
    return loss.mean()


def param_shapes(input_dim, hid, max_diffusion_step, num_supports):
    """Shapes of one cell's four tensors (model/cell.py:35-45)."""
    m = num_supports * max_diffusion_step + 1
    c = input_dim + hid
    return {"Wg": (c * m, 2 * hid), "bg": (2 * hid,), "Wc": (c * m, hid), "bc": (hid,)}
