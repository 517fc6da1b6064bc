"""GPU parity: the CUDA path (through the drop-in modules -> ctypes -> C ABI) against
(a) the golden fixtures produced by the unmodified reference and (b) the oracle on seeded inputs.

Tolerance: BASELINE.json asks for 1e-4 relative (max|d| / max|ref|) in fp32; the generic fp32 path
is expected to sit near 1e-6, so the tests use 2e-5 to catch regressions early."""
import random
import types

import numpy as np
import pytest
import torch

from oracle import dcgru_oracle as O
from oracle import graph_oracle as G
from tests.conftest import load_golden
from tests.helpers import rel_err, supports_of

pytestmark = pytest.mark.gpu
TOL = 2e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _args(meta):
    return types.SimpleNamespace(**{k: meta[k] for k in (
        "num_nodes", "num_rnn_layers", "rnn_units", "input_dim", "output_dim", "max_diffusion_step",
        "dcgru_activation", "filter_type", "dropout", "cl_decay_steps", "use_curriculum_learning")})


def _load_params(model, arr):
    sd = {k[len("param:"):]: torch.tensor(v) for k, v in arr.items() if k.startswith("param:")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    assert all(".decoding_cells." in m for m in missing), missing     # only aliases of the tied cell


def _check_grads(model, arr, tol=TOL):
    worst = 0.0
    for name, p in model.named_parameters():
        ref = arr["grad:" + name]
        assert p.grad is not None, name
        e = rel_err(p.grad.cpu().numpy(), ref)
        worst = max(worst, e)
        assert e < tol, (name, e)
    return worst


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,S,bcast", [(1, 1, False), (2, 1, True), (3, 1, False), (1, 2, False),
                                       (2, 2, False), (3, 2, False)])
def test_graph_poly(dev, K, S, bcast):
    from eeg_gnn_ssl_b200 import ops
    rng = np.random.default_rng(K * 7 + S)
    B, N = 5, 19
    sup = [rng.standard_normal((1 if bcast else B, N, N)).astype(np.float32) * 0.3 for _ in range(S)]
    t = [torch.tensor(s[0] if bcast else s, device=dev) for s in sup]
    P = ops.graph_poly(t, B, N, K).cpu().numpy()
    assert P.shape == (B, S * K, N, N)
    for b in range(B):
        ref = G.diffusion_polynomials([s[0 if bcast else b] for s in sup], K)
        assert rel_err(P[b], ref) < 1e-6


def test_corr_supports_golden(dev):
    from eeg_gnn_ssl_b200 import ops
    _, a = load_golden("graph_supports")
    raw = torch.tensor(a["raw"], device=dev)
    (s0, s1), adj = ops.corr_supports(raw, top_k=3, return_adj=True)
    assert np.array_equal(adj.cpu().numpy() != 0, a["corr_adj"] != 0)
    assert np.abs(adj.cpu().numpy() - a["corr_adj"]).max() < 2e-6
    assert np.abs(s0.cpu().numpy() - a["support0"]).max() < 2e-6
    assert np.abs(s1.cpu().numpy() - a["support1"]).max() < 2e-6
    # standardised input + (std, mean) undoes the scaler
    x = (raw - 3.924) / 1.560
    s0b, s1b = ops.corr_supports(x, top_k=3, scale=1.560, shift=3.924)
    assert np.abs(s0b.cpu().numpy() - a["support0"]).max() < 1e-4


@pytest.mark.parametrize("name", ["enc_cfg1_distance", "enc_corr_relu", "enc_cls_k3"])
def test_encoder_model_vs_reference_golden(dev, name):
    from eeg_gnn_ssl_b200.model.model import DCRNNModel_classification
    meta, a = load_golden(name)
    model = DCRNNModel_classification(_args(meta), meta["classes"]).to(dev)
    _load_params(model, a)
    model.train()
    x = torch.tensor(a["x"], device=dev)
    sup = [s.to(dev) for s in supports_of(a)]
    sl = torch.tensor(a["seq_lengths"], device=dev)
    logits = model(x, sl, sup)
    h0 = model.encoder.init_hidden(meta["batch"]).to(dev)
    with torch.no_grad():
        out_hidden, top = model.encoder(x.transpose(0, 1), h0, sup)
    assert rel_err(top.cpu().numpy(), a["top_seq"]) < TOL
    assert rel_err(out_hidden.cpu().numpy(), a["out_hidden"]) < TOL
    assert rel_err(logits.detach().cpu().numpy(), a["logits"]) < TOL
    if meta["classes"] == 1:
        loss = torch.nn.functional.binary_cross_entropy_with_logits(
            logits.view(-1), torch.tensor(a["y"], dtype=torch.float32, device=dev))
    else:
        loss = torch.nn.functional.cross_entropy(logits, torch.tensor(a["y"], device=dev))
    assert abs(float(loss.detach()) - float(a["loss"])) < 1e-5
    loss.backward()
    _check_grads(model, a)


@pytest.mark.parametrize("name", ["ssl_distance", "ssl_corr_teacher"])
def test_ssl_model_vs_reference_golden(dev, name):
    from eeg_gnn_ssl_b200.model.model import DCRNNModel_nextTimePred
    meta, a = load_golden(name)
    model = DCRNNModel_nextTimePred(_args(meta)).to(dev)
    _load_params(model, a)
    model.train()
    x, y = torch.tensor(a["x"], device=dev), torch.tensor(a["y"], device=dev)
    sup = [s.to(dev) for s in supports_of(a)]
    seen = None
    if meta["teacher"]:
        seen = 9000
        random.seed(22)                        # the seed make_golden.py used for this case
    pred = model(x, y, sup, batches_seen=seen)
    assert rel_err(pred.detach().cpu().numpy(), a["pred"]) < TOL
    loss = O.masked_mae(pred, y)
    assert abs(float(loss.detach()) - float(a["loss"])) < 1e-5
    loss.backward()
    _check_grads(model, a, tol=5e-5)


# ---------------------------------------------------------------------------------------------------
def _rand_case(dev, B, T, H, K, S, L, act, fin=100, seed=0):
    """random encoder case -> (ours outputs+grads, fp64 oracle outputs+grads)"""
    from eeg_gnn_ssl_b200.model.model import DCRNNEncoder
    g = torch.Generator().manual_seed(seed)
    N = 19
    ft = "dual_random_walk" if S == 2 else "laplacian"
    enc = DCRNNEncoder(fin, K, H, N, L, dcgru_activation=act, filter_type=ft)
    with torch.no_grad():
        for p in enc.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    x = torch.randn(T, B, N, fin, generator=g)
    sup = [torch.randn(B, N, N, generator=g) * 0.25 for _ in range(S)]
    h0 = 0.5 * torch.randn(L, B, N * H, generator=g)
    wt = torch.randn(T, B, N * H, generator=g)
    wl = torch.randn(L, B, N * H, generator=g)
    # oracle in fp64 (truth) and in fp32 (what the reference's own arithmetic achieves)
    def run_oracle(dt):
        layers = []
        for c in enc.encoding_cells:
            layers.append({k: v.detach().to(dt).requires_grad_(True) for k, v in zip(
                ("Wg", "bg", "Wc", "bc"), c.flat_params())})
        h0d = h0.detach().clone().to(dt).requires_grad_(True)
        oh, top = O.encoder_forward(x.to(dt), h0d, [s.to(dt) for s in sup], layers, K, N, act)
        ((top * wt.to(dt)).sum() + (oh * wl.to(dt)).sum()).backward()
        out = {"top": top.detach(), "oh": oh.detach(), "dh0": h0d.grad}
        for l in range(L):
            for k in ("Wg", "bg", "Wc", "bc"):
                out[f"L{l}.{k}"] = layers[l][k].grad
        return out
    ref64, ref32 = run_oracle(torch.float64), run_oracle(torch.float32)
    # ours
    enc = enc.to(dev)
    h0g = h0.to(dev).requires_grad_(True)
    oh2, top2 = enc(x.to(dev), h0g, [s.to(dev) for s in sup])
    ((top2 * wt.to(dev)).sum() + (oh2 * wl.to(dev)).sum()).backward()
    ours = {"top": top2.detach(), "oh": oh2.detach(), "dh0": h0g.grad}
    for l, c in enumerate(enc.encoding_cells):
        for k, p in zip(("Wg", "bg", "Wc", "bc"), c.flat_params()):
            ours[f"L{l}.{k}"] = p.grad
    return ours, ref64, ref32


@pytest.mark.parametrize("B,T,H,K,S,L,act", [
    (5, 3, 64, 2, 1, 2, "tanh"),        # batch not a multiple of the CTA's sample group
    (9, 4, 64, 2, 2, 2, "tanh"),        # two supports, carried-x0 quirk
    (3, 2, 128, 3, 2, 2, "tanh"),       # config-5 family: H=128, K=3, M=7
    (6, 3, 32, 1, 1, 1, "relu"),
    (2, 2, 64, 3, 1, 3, "relu"),
    (150, 2, 64, 2, 1, 1, "tanh"),      # more CTAs than one per sample group size 1
])
def test_encoder_vs_oracle_fp64(dev, B, T, H, K, S, L, act):
    ours, ref64, ref32 = _rand_case(dev, B, T, H, K, S, L, act)
    for k in ours:
        e = rel_err(ours[k].cpu().numpy(), ref64[k].numpy())
        e32 = rel_err(ref32[k].numpy(), ref64[k].numpy())      # fp32 round-off of the same math on CPU
        assert e < max(TOL, 8 * e32), (k, e, e32)


def test_cell_single_step(dev):
    from eeg_gnn_ssl_b200.model.cell import DCGRUCell
    torch.manual_seed(3)
    cell = DCGRUCell(100, 64, 2, 19, filter_type="dual_random_walk", nonlinearity="tanh")
    x, h = torch.randn(4, 1900), torch.randn(4, 19 * 64) * 0.5
    sup = [torch.randn(4, 19, 19) * 0.2 for _ in range(2)]
    p = dict(zip(("Wg", "bg", "Wc", "bc"), [t.detach() for t in cell.flat_params()]))
    ref = O.cell_forward(sup, x, h, p, 2, 19, "tanh")
    cell = cell.to(dev)
    out, new = cell([s.to(dev) for s in sup], x.to(dev), h.to(dev))
    assert out is new
    assert rel_err(out.detach().cpu().numpy(), ref.numpy()) < TOL


def test_unsupported_configs_fail_loudly(dev):
    from eeg_gnn_ssl_b200.model.cell import DCGRUCell
    cell = DCGRUCell(100, 48, 2, 19).to(dev)            # hid_dim 48 has no kernel
    with pytest.raises(RuntimeError, match="hid_dim"):
        cell([torch.eye(19, device=dev)], torch.zeros(2, 1900, device=dev), torch.zeros(2, 19 * 48, device=dev))
    cell = DCGRUCell(100, 64, 2, 19)
    with pytest.raises(RuntimeError, match="CUDA"):      # CPU tensors: no fallback
        cell([torch.eye(19)], torch.zeros(2, 1900), torch.zeros(2, 19 * 64))


def test_full_size_properties(dev):
    """BASELINE config 2 size (B=512, T=60, H=64, K=2, L=2): oracle is too slow here, so check
    size-independent properties: sample independence (a sub-batch reproduces the same rows),
    run-to-run determinism, and finite outputs/gradients."""
    from eeg_gnn_ssl_b200.model.model import DCRNNEncoder
    torch.manual_seed(1)
    B, T, N, H = 512, 60, 19, 64
    enc = DCRNNEncoder(100, 2, H, N, 2, dcgru_activation="tanh").to(dev)
    _, a = load_golden("graph_supports")
    lap = torch.tensor(G.scaled_laplacian(a["dist_adj"]).astype(np.float32), device=dev)
    x = torch.randn(T, B, N, 100, device=dev)
    sup = [lap.unsqueeze(0).expand(B, N, N).contiguous()]
    h0 = torch.zeros(2, B, N * H, device=dev)
    oh, top = enc(x, h0, sup)
    top.square().mean().backward()
    g1 = [p.grad.clone() for p in enc.parameters()]
    assert torch.isfinite(top).all() and all(torch.isfinite(g).all() for g in g1)
    enc.zero_grad()
    oh2, top2 = enc(x, h0, sup)
    top2.square().mean().backward()
    assert torch.equal(top, top2)
    assert all(torch.equal(a_, b_.grad) for a_, b_ in zip(g1, enc.parameters()))
    with torch.no_grad():
        _, sub = enc(x[:, 100:107].contiguous(), h0[:, 100:107].contiguous(), [sup[0][100:107].contiguous()])
    assert rel_err(sub.cpu().numpy(), top[:, 100:107].detach().cpu().numpy()) < 1e-5
