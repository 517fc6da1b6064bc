"""Operand-image path of the weight gradient (seq_fwd_tc / seq_bwd_tc dumps -> dw_mm GEMM).

Stage by stage through the C ABI: (1) the G image the forward kernel leaves in HBM against the diffused
operands computed with torch, (2) the dA image of the backward kernel against its row-major dA,
(3) the GEMM over the images against a float64 contraction of the same images, the recompute path
(dw_tc) and the fp32 FMA path; then the module-level parity of both weight-gradient paths."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def first_generation_path(monkeypatch):
    """this file pins the FIRST-generation (3xTF32) operand images through the C ABI; the second-generation (2xFP16)
    stages have their own file, tests/test_gpu_g2.py"""
    monkeypatch.setenv("DCGRU_G2", "0")
N, H, K = 19, 64, 2
SB, IMG_ROWS = 4, 96            # samples per CTA; image rows per (cta, t, hi|lo) = 4 samples x 24 rows (tc_common.cuh)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _layer_case(dev, B, T, fin, seed):
    from eeg_gnn_ssl_b200 import ops
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, B, N * fin, generator=g).to(dev)
    h0 = (0.3 * torch.randn(B, N * H, generator=g)).to(dev)
    sup = [torch.softmax(2 * torch.randn(B, N, N, generator=g), -1).to(dev)]
    P = ops.graph_poly(sup, B, N, K)
    cm = (fin + H) * 3
    wg = (torch.randn(cm, 2 * H, generator=g) * (2.0 / (cm + 2 * H)) ** 0.5 * 1.414).to(dev)
    wc = (torch.randn(cm, H, generator=g) * (2.0 / (cm + H)) ** 0.5 * 1.414).to(dev)
    bg = (0.1 * torch.randn(2 * H, generator=g)).to(dev)
    bc = (0.1 * torch.randn(H, generator=g)).to(dev)
    d_hseq = torch.randn(T, B, N * H, generator=g).to(dev)
    return x, h0, P, (wg, bg, wc, bc), d_hseq


def _run_layer(dev, B, T, fin, seed, use_gsave, want_dx=False):
    """direct C-ABI forward + backward of one encoder layer; returns everything the checks need"""
    from eeg_gnn_ssl_b200 import _lib, ops
    L = _lib.lib()
    x, h0, P, w, d_hseq = _layer_case(dev, B, T, fin, seed)
    desc = ops.make_desc(N, fin, H, K, 1, "tanh")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    h_seq = torch.empty(T, B, N * H, device=dev)
    ruc = torch.empty(T, B, N, 3 * H, device=dev)
    gbytes = L.dcgru_encoder_layer_gsave_bytes(C.byref(desc), B, T) if use_gsave else 0
    gsave = torch.full((gbytes,), 0x7F, device=dev, dtype=torch.uint8) if gbytes else None
    fb = L.dcgru_encoder_layer_fwd_workspace(C.byref(desc), B, T)
    fws = torch.empty(max(fb, 16), device=dev, dtype=torch.uint8)
    wp = ops._params([w])
    _lib.check(L.dcgru_encoder_layer_fwd(C.byref(desc), B, T, _p(x), x.stride(0), x.stride(1), _p(h0), _p(P), wp,
                                         _p(h_seq), _p(ruc), _p(gsave), gbytes, _p(fws), fb, st), "fwd")
    bb = L.dcgru_encoder_layer_bwd_workspace(C.byref(desc), B, T)
    bws = torch.zeros(bb, device=dev, dtype=torch.uint8)
    grads = [torch.full_like(t, float("nan")) for t in w]
    gr = _lib.CellGrads(*(t.data_ptr() for t in grads))
    dh0 = torch.empty_like(h0)
    dx = torch.empty(T, B, N * fin, device=dev) if want_dx else None
    _lib.check(L.dcgru_encoder_layer_bwd(C.byref(desc), B, T, _p(x), x.stride(0), x.stride(1), _p(h0), _p(P), wp,
                                         _p(h_seq), _p(ruc), _p(d_hseq), _p(None), _p(dx), _p(dh0), C.byref(gr),
                                         _p(gsave), gbytes, _p(bws), bb, st), "bwd")
    torch.cuda.synchronize()
    off = (C.c_size_t * 3)()
    _lib.check(L.dcgru_debug_encoder_bwd_offsets(C.byref(desc), B, T, off), "offsets")
    return dict(x=x, h0=h0, P=P, w=w, h_seq=h_seq, ruc=ruc, gsave=gsave, bws=bws, off=list(off), grads=grads,
                dh0=dh0, dx=dx, gbytes=gbytes)


def _unimage(buf, nslab, width, used):
    """row-major image [slab][hi|lo][96 rows][width floats] (bytes) -> float64 hi + lo, first `used` columns"""
    t = buf.view(torch.float32).view(nslab, 2, IMG_ROWS, width).double()
    return (t[:, 0] + t[:, 1])[..., :used].contiguous()


def _expected_g(r, B, T, fin):
    """diffused operands [x | h_prev | r*h_prev] in kk = c*3 + m order, rows = sample*24 + node, per (cta, t)"""
    nxc = (fin + 7) // 8
    ncta = (B + SB - 1) // SB
    P = r["P"].double()                                                   # (B, 2, N, N)
    x = r["x"].double().view(T, B, N, fin)
    hs = torch.cat([r["h0"].double().view(1, B, N, H), r["h_seq"].double().view(T, B, N, H)[:-1]], 0)
    rg = r["ruc"].double()[..., :H]

    def terms(z, width):                                                  # (T,B,N,c) -> (T,B,N,width*3)
        t1 = torch.einsum("bnj,tbjc->tbnc", P[:, 0], z)
        t2 = torch.einsum("bnj,tbjc->tbnc", P[:, 1], z)
        g = torch.stack([z, t1, t2], -1).reshape(T, B, N, -1)
        out = torch.zeros(T, B, N, width * 3, dtype=torch.float64, device=z.device)
        out[..., : g.shape[-1]] = g
        return out

    full = torch.cat([terms(x, nxc * 8), terms(hs, H), terms(rg * hs, H)], -1)      # (T,B,N,KK)
    kk = full.shape[-1]
    img = torch.zeros(ncta, T, IMG_ROWS, kk, dtype=torch.float64, device=full.device)
    for b in range(B):
        c, s = divmod(b, SB)
        img[c, :, s * 24: s * 24 + N] = full[:, b]
    return img.reshape(ncta * T, IMG_ROWS, kk)


@pytest.mark.parametrize("B,T,fin", [(7, 3, 100), (13, 2, 64), (6, 1, 100)])
def test_operand_images_and_gemm(dev, B, T, fin):
    r = _run_layer(dev, B, T, fin, seed=B + fin, use_gsave=True)
    assert r["gbytes"] > 0, "operand-image path not selected on this device"
    ncta = (B + SB - 1) // SB
    nslab = ncta * T
    kgt = ((fin + 7) // 8 + 16) * 6
    # (1) G image
    kkp = (kgt * 4 + 31) // 32 * 32
    gimg = _unimage(r["gsave"], nslab, kkp, kgt * 4)
    gexp = _expected_g(r, B, T, fin)
    eg = float((gimg - gexp).abs().max() / gexp.abs().max())
    print(f"G image vs torch: {eg:.2e}")
    assert eg < 5e-6, eg
    # (2) dA image vs the row-major dA of the same backward call
    o_da, o_img, _ = r["off"]
    assert o_img > 0
    da = r["bws"][o_da: o_da + T * B * N * 3 * H * 4].view(torch.float32).view(T, B, N, 3 * H).double()
    dimg = _unimage(r["bws"][o_img: o_img + nslab * 2 * IMG_ROWS * 3 * H * 4], nslab, 3 * H, 3 * H).view(ncta, T, IMG_ROWS, 3 * H)
    dexp = torch.zeros_like(dimg)
    for b in range(B):
        c, s = divmod(b, SB)
        dexp[c, :, s * 24: s * 24 + N] = da[:, b]
    ed = float((dimg - dexp).abs().max() / dexp.abs().max())
    print(f"dA image vs row-major dA: {ed:.2e}")
    assert ed < 1e-6, ed
    # (3) GEMM over the images
    full = torch.einsum("srk,sro->ko", gimg, dimg.view(nslab, IMG_ROWS, 3 * H))      # (KK, 192)
    nx = ((fin + 7) // 8) * 8 * 3
    ref_g = torch.cat([full[: fin * 3, : 2 * H], full[nx: nx + 3 * H, : 2 * H]], 0)
    ref_c = torch.cat([full[: fin * 3, 2 * H:], full[nx + 3 * H: nx + 6 * H, 2 * H:]], 0)
    dwg, dbg, dwc, dbc = (t.double() for t in r["grads"])
    for nm, got, ref in (("dWg", dwg, ref_g), ("dWc", dwc, ref_c), ("dbg", dbg, da.sum((0, 1, 2))[: 2 * H]),
                         ("dbc", dbc, da.sum((0, 1, 2))[2 * H:])):
        e = float((got - ref).abs().max() / ref.abs().max())
        print(f"{nm} (dw_mm) vs float64 contraction of the images: {e:.2e}")
        assert torch.isfinite(got).all(), nm
        assert e < 5e-6, (nm, e)


@pytest.mark.parametrize("B,T,fin", [(7, 3, 100), (150, 4, 64)])
def test_gsave_vs_recompute_layer(dev, B, T, fin):
    """same layer call with and without the operand image: all gradients and the hidden sequence agree"""
    a = _run_layer(dev, B, T, fin, seed=3, use_gsave=True, want_dx=(fin == 64))
    b = _run_layer(dev, B, T, fin, seed=3, use_gsave=False, want_dx=(fin == 64))
    assert a["gbytes"] > 0 and b["gbytes"] == 0
    assert torch.equal(a["h_seq"], b["h_seq"])
    assert torch.equal(a["dh0"], b["dh0"])
    if fin == 64:
        assert torch.equal(a["dx"], b["dx"])
    for nm, x, y in zip(("dWg", "dbg", "dWc", "dbc"), a["grads"], b["grads"]):
        e = float((x.double() - y.double()).abs().max() / y.double().abs().max())
        print(f"{nm}: operand-image GEMM vs recompute (dw_tc): {e:.2e}")
        assert e < 1e-5, (nm, e)


def _module_grads(dev, B, T, L, seed):
    from eeg_gnn_ssl_b200.model.model import DCRNNEncoder
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    enc = DCRNNEncoder(100, K, H, N, L, dcgru_activation="tanh").to(dev)
    x = torch.randn(T, B, N, 100, generator=g).to(dev)
    sup = [torch.softmax(2 * torch.randn(B, N, N, generator=g), -1).to(dev)]
    h0 = (0.3 * torch.randn(L, B, N * H, generator=g)).to(dev).requires_grad_(True)
    w = torch.randn(T, B, N * H, generator=g).to(dev)
    wl = torch.randn(L, B, N * H, generator=g).to(dev)
    oh, top = enc(x, h0, sup)
    ((top * w).sum() + (oh * wl).sum()).backward()
    return [top.detach().cpu().double(), h0.grad.cpu().double()] + [p.grad.cpu().double() for p in enc.parameters()]


@pytest.mark.parametrize("B,T,L", [(13, 5, 2), (150, 3, 2), (512, 7, 2), (700, 2, 2)])   # 700 clips = 175 CTAs: more than one wave
def test_gsave_module_parity(dev, monkeypatch, B, T, L):
    """DCRNNEncoder through the drop-in modules: operand-image path (default) vs recompute path vs fp32 FMA path"""
    monkeypatch.setenv("DCGRU_DISABLE_TC", "1")
    ref = _module_grads(dev, B, T, L, 21)
    monkeypatch.setenv("DCGRU_DISABLE_TC", "0")
    monkeypatch.setenv("DCGRU_DISABLE_GSAVE", "1")
    rec = _module_grads(dev, B, T, L, 21)
    monkeypatch.setenv("DCGRU_DISABLE_GSAVE", "0")
    got = _module_grads(dev, B, T, L, 21)
    for i, (a, b, c) in enumerate(zip(got, rec, ref)):
        e1 = float((a - c).abs().max() / c.abs().max())
        e2 = float((b - c).abs().max() / c.abs().max())
        assert e1 < 3e-5, (i, e1, e2)
        assert e2 < 3e-5, (i, e1, e2)
