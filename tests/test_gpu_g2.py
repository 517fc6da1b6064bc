"""Stage-level tests of the second-generation (2xFP16 tcgen05) kernels through the C ABI's debug entry points:
each kernel against a float64 torch contraction of the same inputs (the full-model parity is in test_gpu_parity.py /
test_gpu_fullchain.py; these pin down layouts, chunking and the operand images one stage at a time)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
N = 19


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _terms(P, z):
    """[z, P_1 z, ...]: P (B,M-1,N,N), z (T,B,N,C) -> (M,T,B,N,C) in float64"""
    out = [z.double()]
    for m in range(P.shape[1]):
        out.append(torch.einsum("bnj,tbjc->tbnc", P[:, m].double(), z.double()))
    return torch.stack(out, 0)


def _case(dev, B, T, fin, H, M, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, B, N, fin, generator=g)
    P = torch.randn(B, M - 1, N, N, generator=g) * 0.3
    Wg = torch.randn((fin + H) * M, 2 * H, generator=g) * 0.1
    Wc = torch.randn((fin + H) * M, H, generator=g) * 0.1
    bias = torch.randn(3 * H, generator=g) * 0.1
    return [t.to(dev) for t in (x, P, Wg, Wc, bias)]


@pytest.mark.parametrize("B,T,fin,M", [(6, 3, 100, 3), (4, 2, 64, 3), (9, 2, 100, 5), (3, 2, 64, 2), (5, 1, 100, 1)])
def test_xproj_and_operand_image(dev, B, T, fin, M):
    """hoisted x-part: out = sum_m (P_m x_t) [Wg_x | Wc_x]_m + bias, and the fp16 hi/lo image of the diffused operand"""
    from eeg_gnn_ssl_b200 import _lib
    L = _lib.lib()
    H = 64
    x, P, Wg, Wc, bias = _case(dev, B, T, fin, H, M, 7 * B + M)
    out = torch.full((T, B, N, 3 * H), float("nan"), device=dev)
    kxp = (M * fin + 63) // 64 * 64
    cols = kxp + 64                                     # wider than the x part: the dump must leave the rest alone
    ntile = (B + 3) // 4
    img = torch.full((ntile * T, 2, 96, cols), 7.0, device=dev, dtype=torch.float16)
    nbytes = L.dcgru_debug_bulk_dp_workspace(0, fin, H, M)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    _lib.check(L.dcgru_debug_bulk_dp(0, B, T, N, fin, H, M, _ptr(x), _ptr(P), _ptr(Wg), _ptr(Wc), _ptr(bias), _ptr(out),
                                     _ptr(img), cols, 0, _ptr(ws), nbytes,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "debug_bulk_dp")
    torch.cuda.synchronize()
    D = _terms(P, x)                                    # (M,T,B,N,fin)
    W = torch.cat([Wg[: fin * M], Wc[: fin * M]], 1).double().view(fin, M, 3 * H)      # row c*M+m
    ref = torch.einsum("mtbnc,cmo->tbno", D, W) + bias.double()
    assert _rel(out, ref) < 5e-6
    # image: rows (s*24 + n), columns kk = m*fin + c; hi + lo reproduces the fp32 diffusion to 2^-22
    got = img.float()[:, 0] + img.float()[:, 1]        # (ntile*T, 96, cols)
    assert torch.all(img[..., kxp:] == 7.0)            # untouched columns
    want = torch.zeros(ntile, T, 4, 24, kxp, dtype=torch.float64, device=dev)
    Dk = D.permute(1, 2, 3, 0, 4).reshape(T, B, N, M * fin)                           # kk = m*fin + c
    for b in range(B):
        want[b // 4, :, b % 4, :N, : M * fin] = Dk[:, b]
    want = want.view(ntile * T, 96, kxp)
    assert float((got[..., :kxp].double() - want).abs().max()) < 2e-6 * float(want.abs().max())


@pytest.mark.parametrize("B,T,fin,M", [(6, 3, 64, 3), (5, 2, 64, 5), (2, 1, 64, 2)])
def test_dx_bulk(dev, B, T, fin, M):
    """input gradient: dX_t = sum_m P_m^T (dA_t W_x,m^T)"""
    from eeg_gnn_ssl_b200 import _lib
    L = _lib.lib()
    H = 64
    _, P, Wg, Wc, _ = _case(dev, B, T, fin, H, M, 11 * B + M)
    g = torch.Generator().manual_seed(5)
    dA = torch.randn(T, B, N, 3 * H, generator=g).to(dev)
    out = torch.full((T, B, N, fin), float("nan"), device=dev)
    nbytes = L.dcgru_debug_bulk_dp_workspace(1, fin, H, M)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    _lib.check(L.dcgru_debug_bulk_dp(1, B, T, N, fin, H, M, _ptr(dA), _ptr(P), _ptr(Wg), _ptr(Wc), None, _ptr(out),
                                     None, 0, 0, _ptr(ws), nbytes,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "debug_bulk_dp")
    torch.cuda.synchronize()
    W = torch.cat([Wg[: fin * M], Wc[: fin * M]], 1).double().view(fin, M, 3 * H)
    dG = torch.einsum("tbno,cmo->mtbnc", dA.double(), W)                              # (M,T,B,N,fin)
    ref = dG[0].clone()
    for m in range(1, M):
        ref += torch.einsum("bjn,tbjc->tbnc", P[:, m - 1].double(), dG[m])            # P^T
    assert _rel(out, ref) < 5e-6


def test_side_stream_bias_gradient_identical_and_capturable(dev, monkeypatch):
    """The bias gradient runs on the library's side stream next to the dX / dW GEMMs (capi.cu::side_stream): same bits as
    on the caller's stream (DCGRU_SIDE_STREAM=0), also when the whole step is captured in a CUDA graph and replayed."""
    import types
    from eeg_gnn_ssl_b200.model.model import DCRNNModel_classification
    torch.manual_seed(11)
    args = types.SimpleNamespace(num_nodes=N, num_rnn_layers=2, rnn_units=64, input_dim=100, output_dim=100,
                                 max_diffusion_step=2, dcgru_activation="tanh", filter_type="laplacian", dropout=0.0,
                                 cl_decay_steps=3000, use_curriculum_learning=False)
    model = DCRNNModel_classification(args, 1).to(dev)
    B, T = 10, 6
    x = torch.randn(B, T, N, 100, device=dev)
    sl = torch.full((B,), T, device=dev)
    sup = [torch.softmax(torch.randn(B, N, N, device=dev), -1)]

    def grads():
        model.zero_grad(set_to_none=True)
        model(x, sl, sup).square().mean().backward()
        torch.cuda.synchronize()
        return [p.grad.clone() for p in model.parameters()]

    monkeypatch.setenv("DCGRU_FUSE_DB", "0")                  # separate colsum16 kernel ...
    monkeypatch.setenv("DCGRU_SIDE_STREAM", "0")              # ... on the caller's stream
    ref = grads()
    monkeypatch.setenv("DCGRU_SIDE_STREAM", "1")              # ... on the library's side stream
    for a, b in zip(grads(), ref):
        assert torch.equal(a, b)
    # default: db fused into the weight-gradient GEMM (column sums by one more warp of dw_mm16): another summation order for
    # the biases, every other gradient bit-identical
    monkeypatch.delenv("DCGRU_FUSE_DB")
    names = [n for n, _ in model.named_parameters()]
    for n, a, b in zip(names, grads(), ref):
        if n.endswith("biases") and "encoder" in n:
            assert float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) < 5e-6, n
        else:
            assert torch.equal(a, b), n
    monkeypatch.setenv("DCGRU_FUSE_DB", "0")
    # captured: the fork / join must be part of the graph (a side stream left outside would make the capture fail or the
    # replay read stale bias gradients)
    for p in model.parameters():
        p.grad = torch.zeros_like(p)
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        for _ in range(2):
            for p in model.parameters():
                p.grad.zero_()
            model(x, sl, sup).square().mean().backward()
    torch.cuda.current_stream().wait_stream(stream)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for p in model.parameters():
            p.grad.zero_()
        model(x, sl, sup).square().mean().backward()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    for p, b in zip(model.parameters(), ref):
        assert torch.equal(p.grad, b)
