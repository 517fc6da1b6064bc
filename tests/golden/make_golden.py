"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own modules (SURVEY Appendix B stub recipe: the six absent
third-party packages are replaced by empty modules; nothing of the reference is modified),
runs them on seeded CPU fp32 inputs and stores inputs, weights, outputs and every parameter
gradient.  The GPU box has no /root/reference, so these files are what travels.

Fixtures
--------
enc_*.npz   DCRNNModel_classification (encoder + head) forward + backward
ssl_*.npz   DCRNNModel_nextTimePred (encoder + decoder) forward + backward
graph_*.npz reference graph helpers (scaled Laplacian, random walk, correlation top-k)
"""
import json
import os
import pickle
import random
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.path.insert(0, REF)
    for n in ["h5py", "pyedflib", "matplotlib", "matplotlib.cm", "tensorboardX", "dotted_dict"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["dotted_dict"].DottedDict = dict
    import model.model as mm          # noqa: E402
    import utils as ru                # noqa: E402
    import data.data_utils as du      # noqa: E402
    return mm, ru, du


class Args:
    def __init__(self, **kw):
        self.num_nodes = 19
        self.num_rnn_layers = 2
        self.rnn_units = 64
        self.input_dim = 100
        self.output_dim = 100
        self.max_diffusion_step = 2
        self.dcgru_activation = "tanh"
        self.filter_type = "laplacian"
        self.dropout = 0.0
        self.cl_decay_steps = 3000
        self.use_curriculum_learning = False
        self.__dict__.update(kw)


def structured_clip(gen, t, n, f):
    """Raw log-amplitude-like clip whose channel correlations are well separated, so the
    top-k neighbour set is stable under fp32 rounding."""
    base = torch.randn(4, t, f, generator=gen)
    mix = torch.randn(n, 4, generator=gen)
    clip = torch.einsum("nk,ktf->tnf", mix, base) + 0.3 * torch.randn(t, n, f, generator=gen)
    return clip * 1.560 + 3.924


def np32(x):
    return x.detach().cpu().numpy().astype(np.float32)


def make_supports(ru, du, mm, filter_type, raw_clips, top_k=3):
    """raw_clips: (B,T,N,F) numpy.  Uses the reference helpers exactly as the loaders do
    (data/dataloader_detection.py:258-307,335-354) minus the Dataset object."""
    b = raw_clips.shape[0]
    if filter_type == "laplacian":
        adj = pickle.load(open(os.path.join(REF, "data/electrode_graph/adj_mx_3d.pkl"), "rb"))[-1]
        s = torch.FloatTensor(ru.calculate_scaled_laplacian(adj, lambda_max=None).toarray())
        return [s.unsqueeze(0).repeat(b, 1, 1)], np.stack([adj] * b)
    sup = [[], []]
    adjs = []
    for i in range(b):
        clip = raw_clips[i]                                  # (T,N,F)
        n = clip.shape[1]
        x = np.transpose(clip, (1, 0, 2)).reshape(n, -1)
        adj = np.eye(n, n, dtype=np.float32)
        for p in range(n):
            for q in range(p + 1, n):
                xc = du.comp_xcorr(x[p], x[q], mode="valid", normalize=True)
                adj[p, q] = xc
                adj[q, p] = xc
        adj = abs(adj)
        adj = du.keep_topk(adj, top_k=top_k, directed=True)
        adjs.append(adj)
        sup[0].append(torch.FloatTensor(ru.calculate_random_walk_matrix(adj).T.toarray()))
        sup[1].append(torch.FloatTensor(ru.calculate_random_walk_matrix(adj.T).T.toarray()))
    return [torch.stack(sup[0]), torch.stack(sup[1])], np.stack(adjs)


def save(name, meta, arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e3:.0f} kB")


def enc_case(mm, ru, du, name, seed, batch, t_len, classes, seq_lengths=None, **kw):
    args = Args(**kw)
    torch.manual_seed(seed)
    model = mm.DCRNNModel_classification(args, classes)
    # biases start at 0 in the reference; perturb them so their role is exercised
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n_, p_ in model.named_parameters():
            if n_.endswith("biases") or n_.endswith("bias"):
                p_.add_(0.1 * torch.randn(p_.shape, generator=gen))
    raw = torch.stack([structured_clip(gen, t_len, 19, args.input_dim) for _ in range(batch)])
    x = (raw - 3.924) / 1.560
    supports, adjs = make_supports(ru, du, mm, args.filter_type, raw.numpy())
    sl = torch.full((batch,), t_len, dtype=torch.long) if seq_lengths is None \
        else torch.tensor(seq_lengths, dtype=torch.long)
    if seq_lengths is not None:                               # classification loader zero-pads
        for i, l in enumerate(seq_lengths):
            x[i, l:] = 0
    model.train()
    logits = model(x, sl, supports)
    # also record what the encoder itself returns
    with torch.no_grad():
        h0 = model.encoder.init_hidden(batch)
        out_hidden, top_seq = model.encoder(torch.transpose(x, 0, 1), h0, supports)
    if classes == 1:
        y = (torch.rand(batch, generator=gen) > 0.5).float()
        loss = torch.nn.BCEWithLogitsLoss()(logits.view(-1), y)
    else:
        y = torch.randint(0, classes, (batch,), generator=gen)
        loss = torch.nn.CrossEntropyLoss()(logits, y)
    loss.backward()
    arrays = {"x": np32(x), "raw": np32(raw), "y": y.numpy(), "seq_lengths": sl.numpy(),
              "adj": adjs.astype(np.float32),
              "logits": np32(logits), "loss": np32(loss), "out_hidden": np32(out_hidden),
              "top_seq": np32(top_seq)}
    for i, s in enumerate(supports):
        arrays[f"support{i}"] = np32(s)
    for n_, p_ in model.named_parameters():
        arrays["param:" + n_] = np32(p_)
        arrays["grad:" + n_] = np32(p_.grad)
    meta = dict(kind="enc", batch=batch, T=t_len, classes=classes, **args.__dict__)
    save(name, meta, arrays)


def ssl_case(mm, ru, du, name, seed, batch, t_len, to_len, teacher=False, **kw):
    args = Args(**kw)
    torch.manual_seed(seed)
    model = mm.DCRNNModel_nextTimePred(args, device=None)
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n_, p_ in model.named_parameters():
            if n_.endswith("biases") or n_.endswith("bias"):
                p_.add_(0.1 * torch.randn(p_.shape, generator=gen))
    raw = torch.stack([structured_clip(gen, t_len + to_len, 19, args.input_dim)
                       for _ in range(batch)])
    xy = (raw - 3.924) / 1.560
    x, y = xy[:, :t_len].contiguous(), xy[:, t_len:].contiguous()
    supports, adjs = make_supports(ru, du, mm, args.filter_type, raw[:, :t_len].numpy())
    model.train()
    flags = None
    batches_seen = None
    if teacher:
        batches_seen = 9000
        ratio = ru.compute_sampling_threshold(args.cl_decay_steps, batches_seen)
        random.seed(seed)
        flags = [random.random() < ratio for _ in range(to_len)]
        random.seed(seed)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self            # model/model.py:336 hard-codes .cuda()
    try:
        pred = model(x, y, supports, batches_seen=batches_seen)
    finally:
        torch.Tensor.cuda = orig_cuda
    loss = ru.masked_mae_loss(pred, y)
    loss.backward()
    arrays = {"x": np32(x), "y": np32(y), "raw": np32(raw), "adj": adjs.astype(np.float32),
              "pred": np32(pred), "loss": np32(loss)}
    if flags is not None:
        arrays["teacher_flags"] = np.array(flags, dtype=np.bool_)
    for i, s in enumerate(supports):
        arrays[f"support{i}"] = np32(s)
    for n_, p_ in model.named_parameters():                   # tied cell appears once here
        arrays["param:" + n_] = np32(p_)
        arrays["grad:" + n_] = np32(p_.grad)
    arrays["state_keys"] = np.array(list(model.state_dict().keys()))
    meta = dict(kind="ssl", batch=batch, T=t_len, To=to_len, teacher=teacher, **args.__dict__)
    save(name, meta, arrays)


def graph_case(mm, ru, du, name, seed):
    adj = pickle.load(open(os.path.join(REF, "data/electrode_graph/adj_mx_3d.pkl"), "rb"))[-1]
    lap = ru.calculate_scaled_laplacian(adj, lambda_max=None).toarray()
    gen = torch.Generator().manual_seed(seed)
    raw = torch.stack([structured_clip(gen, 12, 19, 100) for _ in range(4)]).numpy()
    sup, adjs = make_supports(ru, du, mm, "dual_random_walk", raw)
    save(name, dict(kind="graph"),
         {"dist_adj": adj.astype(np.float32), "dist_scaled_laplacian": lap.astype(np.float64),
          "raw": raw.astype(np.float32), "corr_adj": adjs.astype(np.float32),
          "support0": np32(sup[0]), "support1": np32(sup[1])})


def main():
    mm, ru, du = import_reference()
    torch.set_num_threads(8)
    # BASELINE.json config 1: distance graph, detection, T=12, K=2, H=64, L=2, B=4
    enc_case(mm, ru, du, "enc_cfg1_distance", 123, batch=4, t_len=12, classes=1)
    # correlation graph (2 supports, carried-x0 quirk), relu cell, H=32
    enc_case(mm, ru, du, "enc_corr_relu", 7, batch=3, t_len=6, classes=1,
             filter_type="dual_random_walk", rnn_units=32, dcgru_activation="relu")
    # config-5 family: 4 classes, correlation graph, K=3, 3 layers, ragged seq lengths, H=32
    enc_case(mm, ru, du, "enc_cls_k3", 11, batch=3, t_len=5, classes=4, seq_lengths=[5, 3, 1],
             filter_type="dual_random_walk", rnn_units=32, max_diffusion_step=3, num_rnn_layers=3)
    # config-4 family: SSL encoder-decoder, distance graph, 3 layers (tied decoder cells)
    ssl_case(mm, ru, du, "ssl_distance", 21, batch=3, t_len=6, to_len=4,
             num_rnn_layers=3, rnn_units=32)
    ssl_case(mm, ru, du, "ssl_corr_teacher", 22, batch=2, t_len=4, to_len=5, teacher=True,
             num_rnn_layers=2, rnn_units=32, filter_type="dual_random_walk",
             use_curriculum_learning=True)
    graph_case(mm, ru, du, "graph_supports", 5)


if __name__ == "__main__":
    main()
