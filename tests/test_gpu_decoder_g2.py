"""Decoder on the second-generation tensor-core kernels (capi.cu: g2_decoder_fwd / the use_g2 branch of dcgru_decoder_bwd):
per-(step, layer) T = 1 launches of xproj + rnn_fwd + projection GEMM, BPTT with rnn_bwd + dX GEMM, one dA operand image.

Checked against (a) the float64 oracle (oracle/dcgru_oracle.py::decoder_forward, model/model.py:149-204) with the same
teacher-forcing draws and dropout masks, and (b) the single-launch fp32 FMA decoder of this library (DCGRU_G2_DEC=0) on the
same inputs.  Shapes: the README SSL decoder (L = 3 with tied upper cells, H = 64, Fo = 100, one support, K = 2) and the
dual-random-walk variant (two supports, M = 5), batch sizes that leave a partial 4-sample tile.
Metric max|d| / max|ref| per tensor; bar 1e-4 (BASELINE.json), expected ~1e-6."""
import os

import numpy as np
import pytest
import torch

from oracle import dcgru_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
N = 19


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _run(dev, dec, tgt, ctx, sup, dout, ratio, seed, g2, gsave=True):
    import random
    os.environ["DCGRU_G2_DEC"] = "1" if g2 else "0"
    os.environ["DCGRU_DISABLE_GSAVE"] = "0" if gsave else "1"
    try:
        dec.zero_grad(set_to_none=True)
        random.seed(seed)
        torch.manual_seed(seed)
        ctx = ctx.clone().requires_grad_(True)
        out = dec(tgt, ctx, sup, teacher_forcing_ratio=ratio)
        (out * dout).sum().backward()
        res = {"out": out.detach().cpu().numpy(), "dctx": ctx.grad.cpu().numpy()}
        for k, p in dec.named_parameters():
            res[k] = p.grad.detach().cpu().numpy().copy()
        return res
    finally:
        os.environ.pop("DCGRU_G2_DEC", None)
        os.environ.pop("DCGRU_DISABLE_GSAVE", None)


@pytest.mark.parametrize("S,L,B,To,ratio,p_drop", [(1, 3, 10, 6, None, 0.0), (1, 3, 7, 5, 0.5, 0.3), (2, 2, 9, 4, 0.5, 0.0)])
def test_decoder_g2_matches_oracle_and_fma(dev, S, L, B, To, ratio, p_drop):
    import random
    from eeg_gnn_ssl_b200.model.model import DCGRUDecoder
    H, K, Fo = 64, 2, 100
    ft = "laplacian" if S == 1 else "dual_random_walk"
    torch.manual_seed(3 + S + L)
    dec = DCGRUDecoder(input_dim=Fo, max_diffusion_step=K, num_nodes=N, hid_dim=H, output_dim=Fo, num_rnn_layers=L,
                       dcgru_activation="tanh", filter_type=ft, dropout=p_drop).to(dev)
    dec.train()
    g = torch.Generator().manual_seed(17 + B)
    with torch.no_grad():
        for p in dec.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g).to(dev))
    tgt = torch.randn(To, B, N, Fo, generator=g).to(dev)
    ctx = (0.5 * torch.randn(L, B, N * H, generator=g)).to(dev)
    sup = [torch.softmax(torch.randn(B, N, N, generator=g), -1).to(dev) for _ in range(S)]
    dout = torch.randn(To, B, N * Fo, generator=g).to(dev)
    seed = 1234
    a = _run(dev, dec, tgt, ctx, sup, dout, ratio, seed, True)
    b = _run(dev, dec, tgt, ctx, sup, dout, ratio, seed, False)
    c = _run(dev, dec, tgt, ctx, sup, dout, ratio, seed, True, gsave=False)     # recompute weight gradient instead of the image GEMM
    for k in a:
        assert rel_err(a[k], b[k]) < 2e-5, (k, rel_err(a[k], b[k]))
        assert rel_err(c[k], b[k]) < 2e-5, (k, rel_err(c[k], b[k]))

    # float64 oracle with the same draws: python's random for teacher forcing, torch's CUDA generator for the masks
    random.seed(seed)
    torch.manual_seed(seed)
    tf = None
    if ratio is not None:
        tf = [random.random() < ratio for _ in range(To)]
    masks = None
    if p_drop > 0:
        ones = torch.ones((B, N, H), device=dev)
        masks = torch.stack([torch.nn.functional.dropout(ones, p_drop, True) for _ in range(To)], 0).double().cpu()
    cells = list(dec.decoding_cells)
    uniq = {}
    layers = []
    for c in cells:
        if id(c) not in uniq:
            uniq[id(c)] = {k: v.detach().cpu().double().clone().requires_grad_(True)
                           for k, v in zip(("Wg", "bg", "Wc", "bc"), c.flat_params())}
        layers.append(uniq[id(c)])
    pw = dec.projection_layer.weight.detach().cpu().double().requires_grad_(True)
    pb = dec.projection_layer.bias.detach().cpu().double().requires_grad_(True)
    c64 = ctx.detach().cpu().double().requires_grad_(True)
    ref = O.decoder_forward(tgt.cpu().double(), c64, [s.cpu().double() for s in sup], layers, pw, pb, K, N, "tanh",
                            teacher_force=tf, dropout_masks=masks)
    (ref * dout.cpu().double()).sum().backward()
    assert rel_err(a["out"], ref.detach().numpy()) < 1e-5
    assert rel_err(a["dctx"], c64.grad.numpy()) < 2e-5
    assert rel_err(a["projection_layer.weight"], pw.grad.numpy()) < 2e-5
    assert rel_err(a["projection_layer.bias"], pb.grad.numpy()) < 2e-5
    names = {"Wg": "dconv_gate.weight", "bg": "dconv_gate.biases", "Wc": "dconv_candidate.weight", "bc": "dconv_candidate.biases"}
    seen = set()
    for l, c in enumerate(cells):
        if id(c) in seen:
            continue
        seen.add(id(c))
        for k, nm in names.items():
            key = f"decoding_cells.{l}.{nm}"
            assert rel_err(a[key], uniq[id(c)][k].grad.numpy()) < 2e-5, key
