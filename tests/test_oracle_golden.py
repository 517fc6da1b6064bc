"""Pin the oracle: replay every golden fixture (made from the unmodified reference by
tests/golden/make_golden.py) through oracle/ and require agreement at fp32 round-off."""
import numpy as np
import pytest
import torch

from oracle import dcgru_oracle as O
from oracle import graph_oracle as G
from tests.conftest import load_golden
from tests.helpers import cell_params, cell_grads, supports_of, rel_err

TOL = 5e-6      # same math, same fp32 ops; op ordering inside stack/cat and the BLAS thread partition of the host differ
GTOL = 5e-5     # gradients (sums over time and batch)


def _filter_S(meta):
    return 2 if meta["filter_type"] == "dual_random_walk" else 1


@pytest.mark.parametrize("name", ["enc_cfg1_distance", "enc_corr_relu", "enc_cls_k3"])
def test_encoder_and_head(name):
    meta, a = load_golden(name)
    L, K, N = meta["num_rnn_layers"], meta["max_diffusion_step"], meta["num_nodes"]
    layers = [cell_params(a, f"encoder.encoding_cells.{l}") for l in range(L)]
    fc_w = torch.tensor(a["param:fc.weight"], requires_grad=True)
    fc_b = torch.tensor(a["param:fc.bias"], requires_grad=True)
    x = torch.tensor(a["x"]).transpose(0, 1)
    sup = supports_of(a)
    h0 = torch.zeros(L, meta["batch"], N * meta["rnn_units"])
    out_hidden, top = O.encoder_forward(x, h0, sup, layers, K, N, meta["dcgru_activation"])
    logits = O.classification_head(top, torch.tensor(a["seq_lengths"]), fc_w, fc_b, N)
    assert rel_err(out_hidden.detach(), a["out_hidden"]) < TOL
    assert rel_err(top.detach(), a["top_seq"]) < TOL
    assert rel_err(logits.detach(), a["logits"]) < TOL
    if meta["classes"] == 1:
        loss = torch.nn.functional.binary_cross_entropy_with_logits(
            logits.view(-1), torch.tensor(a["y"], dtype=torch.float32))
    else:
        loss = torch.nn.functional.cross_entropy(logits, torch.tensor(a["y"]))
    assert abs(float(loss.detach()) - float(a["loss"])) < 2e-6
    loss.backward()
    for l in range(L):
        g = cell_grads(a, f"encoder.encoding_cells.{l}")
        for k in g:
            assert rel_err(layers[l][k].grad, g[k]) < GTOL, (l, k)
    assert rel_err(fc_w.grad, a["grad:fc.weight"]) < GTOL


@pytest.mark.parametrize("name", ["ssl_distance", "ssl_corr_teacher"])
def test_encoder_decoder(name):
    meta, a = load_golden(name)
    L, K, N = meta["num_rnn_layers"], meta["max_diffusion_step"], meta["num_nodes"]
    enc = [cell_params(a, f"encoder.encoding_cells.{l}") for l in range(L)]
    dec0 = cell_params(a, "decoder.decoding_cells.0")
    # named_parameters() lists the tied cell once, under index 1 (model/model.py:126,142-143)
    tied = cell_params(a, "decoder.decoding_cells.1") if L > 1 else None
    dec = [dec0] + [tied] * (L - 1)
    pw = torch.tensor(a["param:decoder.projection_layer.weight"], requires_grad=True)
    pb = torch.tensor(a["param:decoder.projection_layer.bias"], requires_grad=True)
    x = torch.tensor(a["x"]).transpose(0, 1)
    y = torch.tensor(a["y"])
    sup = supports_of(a)
    h0 = torch.zeros(L, meta["batch"], N * meta["rnn_units"])
    ctx, _ = O.encoder_forward(x, h0, sup, enc, K, N, meta["dcgru_activation"])
    flags = list(a["teacher_flags"]) if "teacher_flags" in a else None
    out = O.decoder_forward(y.transpose(0, 1), ctx, sup, dec, pw, pb, K, N,
                            meta["dcgru_activation"], teacher_force=flags)
    pred = out.reshape(meta["To"], meta["batch"], N, -1).transpose(0, 1)
    assert rel_err(pred.detach(), a["pred"]) < TOL
    loss = O.masked_mae(pred, y)
    assert abs(float(loss.detach()) - float(a["loss"])) < 2e-6
    loss.backward()
    for l in range(L):
        g = cell_grads(a, f"encoder.encoding_cells.{l}")
        for k in g:
            assert rel_err(enc[l][k].grad, g[k]) < 5e-5, ("enc", l, k)
    g0 = cell_grads(a, "decoder.decoding_cells.0")
    for k in g0:
        assert rel_err(dec0[k].grad, g0[k]) < 5e-5
    if tied is not None:
        g1 = cell_grads(a, "decoder.decoding_cells.1")
        for k in g1:
            assert rel_err(tied[k].grad, g1[k]) < 5e-5
    assert rel_err(pw.grad, a["grad:decoder.projection_layer.weight"]) < 5e-5


def test_state_dict_keys_of_tied_decoder():
    meta, a = load_golden("ssl_distance")
    keys = [str(k) for k in a["state_keys"]]
    # the tied cell is exported under every index >= 1
    assert "decoder.decoding_cells.1.dconv_gate.weight" in keys
    assert "decoder.decoding_cells.2.dconv_gate.weight" in keys


def test_graph_oracle():
    _, a = load_golden("graph_supports")
    lap = G.scaled_laplacian(a["dist_adj"])
    # the reference keeps the float32 dtype of adj_mx_3d.pkl through D^-1/2 A D^-1/2, the oracle
    # works in float64; the support is cast to float32 afterwards (dataloader_detection.py:353)
    assert np.abs(lap - a["dist_scaled_laplacian"]).max() < 1e-6
    for i in range(a["raw"].shape[0]):
        adj = G.correlation_adjacency(a["raw"][i].astype(np.float32))
        assert np.array_equal(adj != 0, a["corr_adj"][i] != 0)
        assert np.abs(adj - a["corr_adj"][i]).max() < 1e-6
        s = G.dual_random_walk_supports(adj)
        assert np.abs(s[0] - a["support0"][i]).max() < 1e-6
        assert np.abs(s[1] - a["support1"][i]).max() < 1e-6


@pytest.mark.parametrize("K,S", [(1, 1), (2, 1), (2, 2), (3, 2), (1, 2)])
def test_polynomials_match_recurrence(K, S):
    """T_m = P_m Z, including the carried-x0 quirk for S=2."""
    rng = np.random.default_rng(K * 10 + S)
    sup = [rng.standard_normal((19, 19)) * 0.3 for _ in range(S)]
    z = rng.standard_normal((1, 19, 5))
    terms = O.diffusion_terms([torch.tensor(s) for s in sup], torch.tensor(z), K)
    P = G.diffusion_polynomials(sup, K)
    assert P.shape[0] == S * K
    for m in range(1, S * K + 1):
        assert np.abs(P[m - 1] @ z[0] - terms[m][0].numpy()).max() < 1e-10
