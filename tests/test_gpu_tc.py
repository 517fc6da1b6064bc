"""Tensor-core (tcgen05) path: hardware self-test of the UMMA plumbing and parity of the TC kernels."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["g2", "g1"])
def generation(request, monkeypatch):
    """every comparison in this file runs against both tensor-core generations: the default 2xFP16 kernels
    (bulk_dp / rnn_fwd / rnn_bwd / dw_mm16) and the first-generation 3xTF32 kernels (DCGRU_G2=0)"""
    monkeypatch.setenv("DCGRU_G2", "1" if request.param == "g2" else "0")
    return request.param


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("N,K", [(64, 32), (128, 64), (192, 320), (256, 96), (64, 2048), (64, 16384)])
def test_tcgen05_3xtf32_gemm(dev, N, K):
    from eeg_gnn_ssl_b200 import _lib
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    Ad, Bd = A.to(dev), B.to(dev)
    Cd = torch.full((128, N), float("nan"), device=dev)
    L = _lib.lib()
    _lib.check(L.dcgru_tc_selftest(C.c_void_p(Ad.data_ptr()), C.c_void_p(Bd.data_ptr()), C.c_void_p(Cd.data_ptr()),
                                   N, K, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tc_selftest")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    got = Cd.cpu().double()
    err = float((got - ref).abs().max() / ref.abs().max())
    tf32 = float(((A.to(dev) @ B.to(dev).t()).cpu().double() - ref).abs().max() / ref.abs().max())
    print(f"N={N} K={K}: 3xTF32 rel err {err:.3e} (fp32 matmul on device: {tf32:.3e})")
    assert np.isfinite(got.numpy()).all()
    # the tensor core truncates when adding into the fp32 accumulator: the bias grows with the number of
    # accumulation steps (K/8*3), which is why dw_tc flushes TMEM every 32 chunks
    assert err < 2e-6 + 1.5e-8 * K, err


def _grads(dev, B, T, H, K, S, L, seed=0):
    from eeg_gnn_ssl_b200.model.model import DCRNNEncoder
    g = torch.Generator().manual_seed(seed)
    ft = "dual_random_walk" if S == 2 else "laplacian"
    torch.manual_seed(seed)
    enc = DCRNNEncoder(100, K, H, 19, L, dcgru_activation="tanh", filter_type=ft).to(dev)
    x = torch.randn(T, B, 19, 100, generator=g).to(dev)
    sup = [(torch.softmax(torch.randn(B, 19, 19, generator=g), -1)).to(dev) for _ in range(S)]
    h0 = (0.3 * torch.randn(L, B, 19 * H, generator=g)).to(dev)
    w = torch.randn(T, B, 19 * H, generator=g).to(dev)
    _, top = enc(x, h0, sup)
    (top * w).sum().backward()
    return {n: p.grad.detach().cpu().double() for n, p in enc.named_parameters()}


@pytest.mark.parametrize("B,T,H,K,S,L", [(37, 5, 64, 2, 1, 2), (11, 3, 64, 2, 2, 2), (9, 2, 128, 3, 2, 2)])
def test_tc_weight_gradients_match_simt(dev, monkeypatch, B, T, H, K, S, L):
    """dw_tc_kernel (tcgen05, 3xTF32) against the fp32 FMA kernel on the same saved activations"""
    monkeypatch.setenv("DCGRU_DISABLE_TC", "1")
    ref = _grads(dev, B, T, H, K, S, L)
    monkeypatch.setenv("DCGRU_DISABLE_TC", "0")
    got = _grads(dev, B, T, H, K, S, L)
    for n in ref:
        e = float((got[n] - ref[n]).abs().max() / ref[n].abs().max())
        assert e < 1e-5, (n, e)


@pytest.mark.parametrize("B,T,L", [(13, 5, 2), (6, 3, 1), (150, 4, 2)])
def test_tc_forward_matches_simt(dev, monkeypatch, B, T, L):
    """seq_fwd_tc_kernel (tcgen05) against the fp32 FMA kernel: hidden-state sequences and input grads"""
    from eeg_gnn_ssl_b200.model.model import DCRNNEncoder
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("DCGRU_DISABLE_TC", flag)
        g = torch.Generator().manual_seed(7)
        torch.manual_seed(7)
        enc = DCRNNEncoder(100, 2, 64, 19, L, dcgru_activation="tanh").to(dev)
        with torch.no_grad():
            for p in enc.parameters():
                if p.dim() == 1:
                    p.add_(0.1 * torch.randn(p.shape, generator=g).to(dev))
        x = torch.randn(T, B, 19, 100, generator=g).to(dev)
        # row-stochastic supports (what the loaders produce); dense Gaussian ones make 2S^2-I ill conditioned
        sup = [torch.softmax(2 * torch.randn(B, 19, 19, generator=g), -1).to(dev)]
        h0 = (0.3 * torch.randn(L, B, 19 * 64, generator=g)).to(dev).requires_grad_(True)
        w = torch.randn(T, B, 19 * 64, generator=g).to(dev)
        oh, top = enc(x, h0, sup)
        (top * w).sum().backward()
        outs[flag] = (top.detach().cpu().double(), oh.detach().cpu().double(), h0.grad.cpu().double(),
                      enc.encoding_cells[0].dconv_gate.weight.grad.cpu().double())
    for a, b in zip(outs["0"], outs["1"]):
        e = float((a - b).abs().max() / b.abs().max())
        assert e < 2e-5, e


@pytest.mark.parametrize("B,T,L", [(13, 5, 2), (6, 2, 1), (150, 3, 2)])
def test_tc_backward_matches_simt(dev, monkeypatch, B, T, L):
    """seq_bwd_tc_kernel + bulk dX against the fp32 FMA BPTT kernel (all gradients)"""
    from eeg_gnn_ssl_b200.model.model import DCRNNEncoder
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("DCGRU_DISABLE_TC", flag)
        g = torch.Generator().manual_seed(11)
        torch.manual_seed(11)
        enc = DCRNNEncoder(100, 2, 64, 19, L, dcgru_activation="tanh").to(dev)
        x = torch.randn(T, B, 19, 100, generator=g).to(dev)
        sup = [torch.softmax(2 * torch.randn(B, 19, 19, generator=g), -1).to(dev)]
        h0 = (0.3 * torch.randn(L, B, 19 * 64, generator=g)).to(dev).requires_grad_(True)
        w = torch.randn(T, B, 19 * 64, generator=g).to(dev)
        wl = torch.randn(L, B, 19 * 64, generator=g).to(dev)
        oh, top = enc(x, h0, sup)
        ((top * w).sum() + (oh * wl).sum()).backward()
        outs[flag] = [h0.grad.cpu().double()] + [p.grad.cpu().double() for p in enc.parameters()]
    for i, (a, b) in enumerate(zip(outs["0"], outs["1"])):
        e = float((a - b).abs().max() / b.abs().max())
        assert e < 3e-5, (i, e)


def test_tc_path_full_size_accuracy(dev, monkeypatch):
    """BASELINE config 2 size (B=512, T=60, 2 layers): the tensor-core path against the fp32 FMA path.
    This is where accumulation length matters (dW sums 583 680 rows; TMEM is flushed every 16 chunks)."""
    from eeg_gnn_ssl_b200.model.model import DCRNNModel_classification
    import types
    from oracle.graph_oracle import scaled_laplacian
    from tests.conftest import load_golden
    _, a = load_golden("graph_supports")
    lap = torch.tensor(scaled_laplacian(a["dist_adj"]).astype(np.float32), device=dev)
    args = types.SimpleNamespace(num_nodes=19, num_rnn_layers=2, rnn_units=64, input_dim=100, output_dim=100,
                                 max_diffusion_step=2, dcgru_activation="tanh", filter_type="laplacian", dropout=0.0,
                                 cl_decay_steps=3000, use_curriculum_learning=False)
    B, T = 512, 60
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("DCGRU_DISABLE_TC", flag)
        torch.manual_seed(123)
        model = DCRNNModel_classification(args, 1).to(dev)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(B, T, 19, 100, generator=g).to(dev)
        y = (torch.rand(B, generator=g) > 0.5).float().to(dev)
        sl = torch.full((B,), T, dtype=torch.long, device=dev)
        logits = model(x, sl, [lap.unsqueeze(0).expand(B, 19, 19).contiguous()])
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits.view(-1), y)
        loss.backward()
        res[flag] = (logits.detach().cpu().double(), {n: p.grad.cpu().double() for n, p in model.named_parameters()})
    e = float((res["0"][0] - res["1"][0]).abs().max() / res["1"][0].abs().max())
    print(f"logits TC vs fp32: {e:.2e}")
    assert e < 1e-4
    for n in res["1"][1]:
        a_, b_ = res["0"][1][n], res["1"][1][n]
        e = float((a_ - b_).abs().max() / b_.abs().max())
        print(f"{n}: {e:.2e}")
        assert e < 1e-4, (n, e)


def _probe(dev, a_img, b_img, a_off, b_off, a_mn, b_mn, a_type, b_type, n):
    from eeg_gnn_ssl_b200 import _lib
    L = _lib.lib()
    a = torch.tensor(a_img, dtype=torch.float32, device=dev)
    b = torch.tensor(b_img, dtype=torch.float32, device=dev)
    d = torch.full((128, n), float("nan"), device=dev)
    _lib.check(L.dcgru_tc_probe(C.c_void_p(a.data_ptr()), a.numel() * 4, C.c_void_p(b.data_ptr()), b.numel() * 4,
                                a_off[0], a_off[1], b_off[0], b_off[1], a_mn, b_mn, a_type, b_type,
                                C.c_void_p(d.data_ptr()), n, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "probe")
    torch.cuda.synchronize()
    return d.cpu().numpy()


def test_umma_operand_layouts_pinned_on_hardware(dev):
    """The shared-memory layouts the kernels rely on, read back from the tensor core itself: one MMA with an
    index-valued A image and a one-hot B shows which word the hardware reads for a logical (row, k)."""
    idx = np.arange(2048, dtype=np.float32)                  # exact in tf32
    onehot_b = np.zeros(16 * 8, dtype=np.float32)            # K-major no-swizzle B, N = 16: [k/4][n][4], one-hot n == k
    for n in range(8):
        onehot_b[(n // 4) * 64 + n * 4 + n % 4] = 1.0
    m = np.arange(128)[:, None]
    k = np.arange(8)[None, :]
    # (1) K-major, no swizzle (weights): [k/4][row][4], LBO = 2048, SBO = 128
    d = _probe(dev, idx, onehot_b, (2048, 128), (256, 128), 0, 0, 0, 0, 16)
    assert (d[:, :8] == (k // 4) * 512 + m * 4 + k % 4).all()
    # (2) K-major, 32-byte swizzle (activation tiles of the sequence kernels): rows of 8 k, halves swapped in rows 4-7
    d = _probe(dev, idx, onehot_b, (16, 256), (256, 128), 0, 0, 6, 0, 16)
    assert (d[:, :8] == (m // 8) * 64 + (m % 8) * 8 + ((k // 4) ^ ((m % 8) >> 2)) * 4 + k % 4).all()
    # (3) MN-major fp32 in the no-swizzle layout is NOT read by the tensor core (exact zeros) ...
    d = _probe(dev, idx, onehot_b, (128, 128), (256, 128), 1, 0, 0, 0, 16)
    assert (d[:, :8] == 0).all()
    # (4) ... MN-major needs layout type 1: 128-byte rows of 32 mn, 32-byte chunks XOR-ed with the row (dw_mm.cu)
    d = _probe(dev, idx, onehot_b, (1024, 512), (256, 128), 1, 0, 1, 0, 16)
    exp = (m // 32) * 256 + (k // 4) * 128 + (k % 4) * 32 + ((((m % 32) // 8) ^ (k % 4)) * 8) + m % 8
    assert (d[:, :8] == exp).all()
