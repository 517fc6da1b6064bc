"""Fused classification head (csrc/head.cu) and the sparse upstream gradient of the top encoder layer.

* dcgru_cls_head_fwd / _bwd through the C ABI against the reference's own formulation (gather at seq_len-1 ->
  dropout mask -> ReLU -> Linear -> max over nodes, model/model.py:257-270, utils.py:346-357) evaluated with torch
  in float64 on the same inputs and the same dropout mask, ragged lengths, 1 and 4 classes;
* dcgru_encoder_layer_bwd_sel (slab + step index) against dcgru_encoder_layer_bwd fed with the dense gradient the
  gather's autograd produces -- tensor-core path (H = 64) and the fp32 path (H = 32, expands the slab internally).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
N = 19


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


@pytest.mark.parametrize("H,ncls,p_drop", [(64, 1, 0.0), (64, 4, 0.5), (128, 4, 0.3), (32, 1, 0.5)])
def test_head_matches_reference_formulation(dev, H, ncls, p_drop):
    from eeg_gnn_ssl_b200 import _lib
    L = _lib.lib()
    B, T = 37, 9
    g = torch.Generator().manual_seed(5 + H + ncls)
    h_seq = torch.randn(T, B, N * H, generator=g).to(dev)
    lens = torch.randint(1, T + 1, (B,), generator=g)
    sel = (lens - 1).to(torch.int32).to(dev)
    fw = (0.3 * torch.randn(ncls, H, generator=g)).to(dev)
    fb = (0.1 * torch.randn(ncls, generator=g)).to(dev)
    mask = None
    if p_drop > 0:
        mask = ((torch.rand(B, N, H, generator=g) > p_drop).float() / (1 - p_drop)).to(dev)
    dlog = torch.randn(B, ncls, generator=g).to(dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    logits = torch.empty(B, ncls, device=dev)
    arg = torch.empty(B, ncls, device=dev, dtype=torch.int32)
    _lib.check(L.dcgru_cls_head_fwd(B, T, N, H, ncls, _ptr(h_seq), _ptr(sel), _ptr(mask), _ptr(fw), _ptr(fb), _ptr(logits),
                                    _ptr(arg), st), "cls_head_fwd")
    d_hsel = torch.empty(B, N * H, device=dev)
    dfw, dfb = torch.empty_like(fw), torch.empty_like(fb)
    nb = L.dcgru_cls_head_bwd_workspace(B, H, ncls)
    ws = torch.empty(nb, device=dev, dtype=torch.uint8)
    _lib.check(L.dcgru_cls_head_bwd(B, T, N, H, ncls, _ptr(h_seq), _ptr(sel), _ptr(mask), _ptr(fw), _ptr(arg), _ptr(dlog),
                                    _ptr(d_hsel), _ptr(dfw), _ptr(dfb), _ptr(ws), nb, st), "cls_head_bwd")
    torch.cuda.synchronize()

    # the reference's formulation in float64 (autograd backward)
    hs = h_seq.double().requires_grad_(True)
    w64, b64 = fw.double().requires_grad_(True), fb.double().requires_grad_(True)
    out = hs.transpose(0, 1)                                                   # (B,T,NH)
    idx = (lens.to(dev) - 1).view(-1, 1).expand(B, out.size(2)).unsqueeze(1)
    last = out.gather(1, idx).squeeze(1).view(B, N, H)
    if mask is not None:
        last = last * mask.double()
    z = torch.relu(last) @ w64.t() + b64
    ref, ref_arg = z.max(dim=1)
    (ref * dlog.double()).sum().backward()

    assert rel_err(logits.cpu().numpy(), ref.detach().cpu().numpy()) < 2e-6
    assert torch.equal(arg.long().cpu(), ref_arg.cpu())
    dense = torch.zeros(T, B, N * H, device=dev)
    dense[sel.long(), torch.arange(B, device=dev)] = d_hsel
    assert rel_err(dense.cpu().numpy(), hs.grad.cpu().numpy()) < 2e-6
    assert rel_err(dfw.cpu().numpy(), w64.grad.cpu().numpy()) < 5e-6
    assert rel_err(dfb.cpu().numpy(), b64.grad.cpu().numpy()) < 5e-6


@pytest.mark.parametrize("H,K,S", [(64, 2, 1), (64, 2, 2), (32, 2, 1)])
def test_sparse_upstream_gradient_equals_dense(dev, H, K, S):
    """top layer + fused head vs the same layer followed by torch's gather/ReLU/Linear/max (dense d_hseq)"""
    from eeg_gnn_ssl_b200 import ops
    B, T, fin, ncls = 10, 7, 64, 4
    M = S * K + 1
    g = torch.Generator().manual_seed(11 + H + S)
    x = torch.randn(T, B, N * fin, generator=g).to(dev)
    h0 = (0.1 * torch.randn(B, N * H, generator=g)).to(dev)
    sup = [torch.softmax(torch.randn(B, N, N, generator=g), -1).to(dev) for _ in range(S)]
    p = ops.graph_poly(sup, B, N, K)
    cm = (fin + H) * M

    def leaf(*shape, s=0.1):
        return (s * torch.randn(*shape, generator=g)).to(dev).requires_grad_(True)
    wg, bg, wc, bc = leaf(cm, 2 * H), leaf(2 * H), leaf(cm, H), leaf(H)
    fw, fb = leaf(ncls, H, s=0.3), leaf(ncls)
    lens = torch.randint(1, T + 1, (B,), generator=g)
    sel = (lens - 1).to(torch.int32).to(dev)
    mask = ((torch.rand(B, N, H, generator=g) > 0.4).float() / 0.6).to(dev)
    dlog = torch.randn(B, ncls, generator=g).to(dev)
    desc = ops.make_desc(N, fin, H, K, S, "tanh")
    params = (wg, bg, wc, bc, fw, fb)

    logits = ops.encoder_top_head(x, h0, p, wg, bg, wc, bc, fw, fb, desc, sel, mask)
    (logits * dlog).sum().backward()
    ours = [q.grad.clone() for q in params]
    for q in params:
        q.grad = None

    h_seq, _ = ops.encoder_layer(x, h0, p, wg, bg, wc, bc, desc)
    idx = sel.long().view(1, B, 1).expand(1, B, N * H)
    last = h_seq.gather(0, idx).squeeze(0).view(B, N, H) * mask
    ref_logits = (torch.relu(last) @ fw.t() + fb).max(dim=1).values
    (ref_logits * dlog).sum().backward()
    assert rel_err(logits.detach().cpu().numpy(), ref_logits.detach().cpu().numpy()) < 2e-6
    for name, a, q in zip(("Wg", "bg", "Wc", "bc", "fc.w", "fc.b"), ours, params):
        assert rel_err(a.cpu().numpy(), q.grad.cpu().numpy()) < 2e-5, name
