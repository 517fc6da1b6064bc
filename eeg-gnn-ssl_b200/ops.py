"""torch-facing operators over the C ABI: device-memory plumbing + autograd glue only.

Every function here takes CUDA fp32 tensors and launches kernels of libdcgru_b200.so on the current
stream; nothing is computed with torch ops (torch allocates buffers and carries the autograd graph).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import CellDesc, CellParams, CellGrads, check


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise RuntimeError("dcgru_b200 operates on CUDA float32 tensors only (no CPU fallback): got "
                               f"{t.device}/{t.dtype}")


def make_desc(num_nodes, input_dim, hid_dim, max_diffusion_step, num_supports, activation):
    return CellDesc(num_nodes, input_dim, hid_dim, max_diffusion_step, num_supports,
                    _lib.ACT_TANH if activation == "tanh" else _lib.ACT_RELU)


def _params(ws):
    arr = (CellParams * len(ws))()
    for i, (wg, bg, wc, bc) in enumerate(ws):
        arr[i] = CellParams(wg.data_ptr(), bg.data_ptr(), wc.data_ptr(), bc.data_ptr())
    return arr


# ------------------------------------------------------------------------------------------------
# graph
# ------------------------------------------------------------------------------------------------
def graph_poly(supports, batch, num_nodes, max_diffusion_step):
    """supports: list of S tensors (B,N,N) or (N,N) -> P (B, S*K, N, N) with term_m = P_m Z
    (model/cell.py:76-93 as linear maps; carried x0 across supports included)."""
    s_count = len(supports)
    m1 = s_count * max_diffusion_step
    dev = supports[0].device
    p = torch.empty((batch, m1, num_nodes, num_nodes), device=dev, dtype=torch.float32)
    if m1 == 0:
        return p
    keep, ptrs, strides = [], (C.c_void_p * s_count)(), (C.c_int64 * s_count)()
    for i, s in enumerate(supports):
        _need_cuda(s)
        s = s.contiguous()
        keep.append(s)
        if s.dim() == 2:
            strides[i] = 0
        elif s.shape[0] == batch:
            strides[i] = num_nodes * num_nodes
        elif s.shape[0] == 1:
            strides[i] = 0
        else:
            raise ValueError(f"support batch {s.shape[0]} != {batch}")
        if tuple(s.shape[-2:]) != (num_nodes, num_nodes):
            raise ValueError(f"support shape {tuple(s.shape)} does not end in ({num_nodes},{num_nodes})")
        ptrs[i] = s.data_ptr()
    check(_lib.lib().dcgru_graph_poly(batch, num_nodes, max_diffusion_step, s_count, ptrs, strides,
                                      _ptr(p), _stream()), "graph_poly")
    return p


def corr_supports(clip, top_k=3, scale=1.0, shift=0.0, return_adj=False):
    """clip (B,T,N,F) -> [support0, support1] ((B,N,N) each) of the per-clip correlation graph
    (data/dataloader_detection.py:258-307,343-347).  ``clip*scale+shift`` is the raw clip."""
    _need_cuda(clip)
    if clip.dim() != 4:
        raise ValueError("clip must be (B,T,N,F)")
    if clip.stride(3) != 1 or clip.stride(2) != clip.shape[3]:
        clip = clip.contiguous()
    b, t, n, f = clip.shape
    s0 = torch.empty((b, n, n), device=clip.device, dtype=torch.float32)
    s1 = torch.empty_like(s0)
    adj = torch.empty_like(s0) if return_adj else None
    check(_lib.lib().dcgru_corr_supports(b, t, n, f, _ptr(clip), clip.stride(0), clip.stride(1),
                                         float(scale), float(shift), int(top_k), _ptr(adj), _ptr(s0),
                                         _ptr(s1), _stream()), "corr_supports")
    return ([s0, s1], adj) if return_adj else [s0, s1]


def fft_features(signal, mean=None, std=None, dest_channel=None, log_scale=None, return_raw=False, window=200):
    """signal (B, N, T*200) raw resampled EEG on the GPU -> x (B, T, N, 100): per-second log-amplitude spectrum
    (data/data_utils.py:13-34 via data/dataloader_detection.py:58-72), optional augmentation
    (``dest_channel`` (B,N) int32 = _random_reflect's pair swaps, ``log_scale`` (B) = log of _random_scale's factor,
    data/dataloader_detection.py:233-256) and StandardScaler.transform (utils.py:402-403; ``mean``/``std`` scalars or
    per-channel tensors).  ``return_raw`` also returns the un-augmented, un-scaled features, which is what the
    correlation graph is built from (``corr_supports``)."""
    _need_cuda(signal)
    if signal.dim() != 3 or signal.shape[2] % window:
        raise ValueError("signal must be (B, N, T*window)")
    if signal.dtype != torch.float32:
        raise ValueError("signal must be float32")
    if signal.stride(2) != 1 or signal.stride(0) % 4 or signal.stride(1) % 4 or signal.data_ptr() % 16:
        signal = signal.contiguous()
    b, n, s = signal.shape
    t = s // window
    dev = signal.device
    stat_len = 0
    if (mean is None) != (std is None):
        raise ValueError("mean and std go together")
    if mean is not None:
        mean = torch.as_tensor(mean, dtype=torch.float32, device=dev).reshape(-1).contiguous()
        std = torch.as_tensor(std, dtype=torch.float32, device=dev).reshape(-1).contiguous()
        if mean.numel() != std.numel() or mean.numel() not in (1, n):
            raise ValueError("mean/std must be scalars or one value per channel")
        stat_len = mean.numel()
    if dest_channel is not None:
        dest_channel = dest_channel.to(device=dev, dtype=torch.int32).contiguous()
        if tuple(dest_channel.shape) != (b, n):
            raise ValueError("dest_channel must be (B, N)")
    if log_scale is not None:
        log_scale = log_scale.to(device=dev, dtype=torch.float32).contiguous()
        if log_scale.numel() != b:
            raise ValueError("log_scale must be (B,)")
    x = torch.empty((b, t, n, window // 2), device=dev, dtype=torch.float32)
    raw = torch.empty_like(x) if return_raw else None
    check(_lib.lib().dcgru_fft_features(b, n, t, window, _ptr(signal), signal.stride(0), signal.stride(1),
                                        _ptr(dest_channel), _ptr(log_scale), _ptr(mean), _ptr(std), stat_len,
                                        _ptr(raw), _ptr(x), _stream()), "fft_features")
    return (x, raw) if return_raw else x


# ------------------------------------------------------------------------------------------------
# encoder layer
# ------------------------------------------------------------------------------------------------
def _seq_view(x):
    """(T,B,N*F) view whose last dim is contiguous; returns (tensor, stride_t, stride_b)."""
    if x.stride(2) != 1 or x.stride(0) % 4 or x.stride(1) % 4 or x.data_ptr() % 16:
        x = x.contiguous()
    return x, x.stride(0), x.stride(1)


class _EncoderLayerFn(torch.autograd.Function):
    """h_seq, h_last = layer(x_seq, h0): one persistent kernel over all T steps."""

    @staticmethod
    def forward(ctx, x, h0, p, wg, bg, wc, bc, desc):
        _need_cuda(x, h0, wg, bg, wc, bc)
        t_len, b = x.shape[0], x.shape[1]
        x, st, sb = _seq_view(x)
        h0 = h0.contiguous()
        nh = desc.num_nodes * desc.hid_dim
        h_seq = torch.empty((t_len, b, nh), device=x.device, dtype=torch.float32)
        need = any(ctx.needs_input_grad)
        ruc = torch.empty((t_len, b, desc.num_nodes, 3 * desc.hid_dim), device=x.device,
                          dtype=torch.float32) if need else None
        wg, bg, wc, bc = wg.contiguous(), bg.contiguous(), wc.contiguous(), bc.contiguous()
        ws = _params([(wg, bg, wc, bc)])
        L = _lib.lib()
        nbytes = L.dcgru_encoder_layer_fwd_workspace(C.byref(desc), b, t_len)
        ws_buf = torch.empty(max(nbytes, 16), device=x.device, dtype=torch.uint8)
        # operand image for the weight-gradient GEMM (tensor-core configurations only; 0 bytes otherwise)
        gbytes = L.dcgru_encoder_layer_gsave_bytes(C.byref(desc), b, t_len) if need else 0
        gsave = torch.empty(gbytes, device=x.device, dtype=torch.uint8) if gbytes else None
        check(L.dcgru_encoder_layer_fwd(C.byref(desc), b, t_len, _ptr(x), st, sb, _ptr(h0), _ptr(p), ws,
                                        _ptr(h_seq), _ptr(ruc), _ptr(gsave), gbytes, _ptr(ws_buf), nbytes,
                                        _stream()),
              "encoder_layer_fwd")
        ctx.desc = desc
        ctx.strides = (st, sb)
        ctx.gsave = gsave
        ctx.save_for_backward(x, h0, p, wg, bg, wc, bc, h_seq, ruc)
        ctx.set_materialize_grads(False)
        h_last = h_seq[t_len - 1].clone()
        return h_seq, h_last

    @staticmethod
    def backward(ctx, d_hseq, d_hlast):
        x, h0, p, wg, bg, wc, bc, h_seq, ruc = ctx.saved_tensors
        desc = ctx.desc
        t_len, b = x.shape[0], x.shape[1]
        st, sb = ctx.strides
        dev = x.device
        if d_hseq is not None:
            d_hseq = d_hseq.contiguous()
        if d_hlast is not None:
            d_hlast = d_hlast.contiguous()
        dx = torch.empty((t_len, b, x.shape[2]), device=dev, dtype=torch.float32) \
            if ctx.needs_input_grad[0] else None
        dh0 = torch.empty_like(h0)
        dwg, dbg, dwc, dbc = (torch.empty_like(t) for t in (wg, bg, wc, bc))
        L = _lib.lib()
        nbytes = L.dcgru_encoder_layer_bwd_workspace(C.byref(desc), b, t_len)
        ws_buf = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        ws = _params([(wg, bg, wc, bc)])
        g = CellGrads(dwg.data_ptr(), dbg.data_ptr(), dwc.data_ptr(), dbc.data_ptr())
        check(L.dcgru_encoder_layer_bwd(C.byref(desc), b, t_len, _ptr(x), st, sb, _ptr(h0), _ptr(p), ws,
                                        _ptr(h_seq), _ptr(ruc), _ptr(d_hseq), _ptr(d_hlast), _ptr(dx),
                                        _ptr(dh0), C.byref(g), _ptr(ctx.gsave),
                                        ctx.gsave.numel() if ctx.gsave is not None else 0,
                                        _ptr(ws_buf), nbytes, _stream()),
              "encoder_layer_bwd")
        ctx.gsave = None
        return dx, dh0, None, dwg, dbg, dwc, dbc, None


def encoder_layer(x, h0, p, wg, bg, wc, bc, desc):
    """x (T,B,N*Fin), h0 (B,N*H) -> (h_seq (T,B,N*H), h_last (B,N*H))."""
    return _EncoderLayerFn.apply(x, h0, p, wg, bg, wc, bc, desc)


class _EncoderTopHeadFn(torch.autograd.Function):
    """logits = head(top_layer(x_seq, h0)): the encoder's last layer and the classification head
    (gather at seq_len-1 -> dropout -> ReLU -> Linear -> max over nodes, model/model.py:257-270) as one autograd
    node, so the head's gradient reaches the BPTT kernel as one (B,N*H) slab + the step it belongs to instead
    of a dense (T,B,N*H) tensor."""

    @staticmethod
    def forward(ctx, x, h0, p, wg, bg, wc, bc, fc_w, fc_b, desc, sel_t, drop_mask):
        _need_cuda(x, h0, wg, bg, wc, bc, fc_w, fc_b, drop_mask)
        t_len, b = x.shape[0], x.shape[1]
        x, st, sb = _seq_view(x)
        h0 = h0.contiguous()
        n, hid = desc.num_nodes, desc.hid_dim
        ncls = fc_w.shape[0]
        dev = x.device
        h_seq = torch.empty((t_len, b, n * hid), device=dev, dtype=torch.float32)
        need = any(ctx.needs_input_grad)
        ruc = torch.empty((t_len, b, n, 3 * hid), device=dev, dtype=torch.float32) if need else None
        wg, bg, wc, bc = wg.contiguous(), bg.contiguous(), wc.contiguous(), bc.contiguous()
        fc_w, fc_b = fc_w.contiguous(), fc_b.contiguous()
        if sel_t is not None:
            sel_t = sel_t.to(device=dev, dtype=torch.int32).contiguous()
        if drop_mask is not None:
            drop_mask = drop_mask.contiguous()
        ws = _params([(wg, bg, wc, bc)])
        L = _lib.lib()
        nbytes = L.dcgru_encoder_layer_fwd_workspace(C.byref(desc), b, t_len)
        ws_buf = torch.empty(max(nbytes, 16), device=dev, dtype=torch.uint8)
        gbytes = L.dcgru_encoder_layer_gsave_bytes(C.byref(desc), b, t_len) if need else 0
        gsave = torch.empty(gbytes, device=dev, dtype=torch.uint8) if gbytes else None
        check(L.dcgru_encoder_layer_fwd(C.byref(desc), b, t_len, _ptr(x), st, sb, _ptr(h0), _ptr(p), ws,
                                        _ptr(h_seq), _ptr(ruc), _ptr(gsave), gbytes, _ptr(ws_buf), nbytes,
                                        _stream()), "encoder_layer_fwd")
        logits = torch.empty((b, ncls), device=dev, dtype=torch.float32)
        arg = torch.empty((b, ncls), device=dev, dtype=torch.int32)
        check(L.dcgru_cls_head_fwd(b, t_len, n, hid, ncls, _ptr(h_seq), _ptr(sel_t), _ptr(drop_mask), _ptr(fc_w),
                                   _ptr(fc_b), _ptr(logits), _ptr(arg), _stream()), "cls_head_fwd")
        ctx.desc, ctx.strides, ctx.gsave = desc, (st, sb), gsave
        ctx.save_for_backward(x, h0, p, wg, bg, wc, bc, fc_w, h_seq, ruc, sel_t, drop_mask, arg)
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        x, h0, p, wg, bg, wc, bc, fc_w, h_seq, ruc, sel_t, drop_mask, arg = ctx.saved_tensors
        desc = ctx.desc
        t_len, b = x.shape[0], x.shape[1]
        st, sb = ctx.strides
        dev = x.device
        n, hid, ncls = desc.num_nodes, desc.hid_dim, fc_w.shape[0]
        d_logits = d_logits.contiguous()
        L = _lib.lib()
        d_hsel = torch.empty((b, n * hid), device=dev, dtype=torch.float32)
        dfw, dfb = torch.empty_like(fc_w), torch.empty((ncls,), device=dev, dtype=torch.float32)
        hbytes = L.dcgru_cls_head_bwd_workspace(b, hid, ncls)
        hws = torch.empty(max(hbytes, 16), device=dev, dtype=torch.uint8)
        check(L.dcgru_cls_head_bwd(b, t_len, n, hid, ncls, _ptr(h_seq), _ptr(sel_t), _ptr(drop_mask), _ptr(fc_w),
                                   _ptr(arg), _ptr(d_logits), _ptr(d_hsel), _ptr(dfw), _ptr(dfb), _ptr(hws), hbytes,
                                   _stream()), "cls_head_bwd")
        dx = torch.empty((t_len, b, x.shape[2]), device=dev, dtype=torch.float32) \
            if ctx.needs_input_grad[0] else None
        dh0 = torch.empty_like(h0)
        dwg, dbg, dwc, dbc = (torch.empty_like(t) for t in (wg, bg, wc, bc))
        nbytes = L.dcgru_encoder_layer_bwd_sel_workspace(C.byref(desc), b, t_len)
        ws_buf = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        ws = _params([(wg, bg, wc, bc)])
        g = CellGrads(dwg.data_ptr(), dbg.data_ptr(), dwc.data_ptr(), dbc.data_ptr())
        check(L.dcgru_encoder_layer_bwd_sel(C.byref(desc), b, t_len, _ptr(x), st, sb, _ptr(h0), _ptr(p), ws,
                                            _ptr(h_seq), _ptr(ruc), _ptr(d_hsel), _ptr(sel_t), _ptr(None), _ptr(dx),
                                            _ptr(dh0), C.byref(g), _ptr(ctx.gsave),
                                            ctx.gsave.numel() if ctx.gsave is not None else 0,
                                            _ptr(ws_buf), nbytes, _stream()), "encoder_layer_bwd_sel")
        ctx.gsave = None
        return dx, dh0, None, dwg, dbg, dwc, dbc, dfw, dfb, None, None, None


def encoder_top_head(x, h0, p, wg, bg, wc, bc, fc_w, fc_b, desc, sel_t=None, drop_mask=None):
    """x (T,B,N*Fin), h0 (B,N*H), fc_w (C,H), fc_b (C), sel_t (B) int32 = seq_len-1 (None: T-1),
    drop_mask (B,N,H) or None -> pooled logits (B,C)."""
    return _EncoderTopHeadFn.apply(x, h0, p, wg, bg, wc, bc, fc_w, fc_b, desc, sel_t, drop_mask)


# ------------------------------------------------------------------------------------------------
# decoder
# ------------------------------------------------------------------------------------------------
class _DecoderFn(torch.autograd.Function):
    """outputs = decoder(targets, h0): all To steps x L cells + projection in one persistent kernel.

    ``cells``: list of L indices into the flat unique-parameter list (tied cells repeat an index)."""

    @staticmethod
    def forward(ctx, targets, h0, p, proj_w, proj_b, drop_mask, desc, num_layers, to_len, teacher_mask,
                cells, *flat):
        _need_cuda(h0, proj_w, proj_b, *flat)
        b = h0.shape[1]
        n, hid, fo = desc.num_nodes, desc.hid_dim, desc.input_dim
        dev = h0.device
        h0 = h0.contiguous()
        flat = tuple(t.contiguous() for t in flat)
        if targets is not None:
            targets = targets.contiguous()
        if drop_mask is not None:
            drop_mask = drop_mask.contiguous()
        proj_w, proj_b = proj_w.contiguous(), proj_b.contiguous()
        out = torch.empty((to_len, b, n * fo), device=dev, dtype=torch.float32)
        h_all = torch.empty((to_len, num_layers, b, n * hid), device=dev, dtype=torch.float32)
        ruc = torch.empty((to_len, num_layers, b, n, 3 * hid), device=dev, dtype=torch.float32)
        L = _lib.lib()
        ws = _params([tuple(flat[4 * c: 4 * c + 4]) for c in cells])
        nbytes = L.dcgru_decoder_fwd_workspace(C.byref(desc), num_layers, b, to_len)
        ws_buf = torch.empty(max(nbytes, 16), device=dev, dtype=torch.uint8)
        # operand image for the weight-gradient GEMMs (tensor-core configurations only; 0 bytes otherwise)
        need = any(ctx.needs_input_grad)
        gbytes = L.dcgru_decoder_gsave_bytes(C.byref(desc), num_layers, b, to_len) if need else 0
        gsave = torch.empty(gbytes, device=dev, dtype=torch.uint8) if gbytes else None
        check(L.dcgru_decoder_fwd_saved(C.byref(desc), num_layers, b, to_len, _ptr(targets), teacher_mask, _ptr(h0),
                                        _ptr(p), ws, _ptr(proj_w), _ptr(proj_b), _ptr(drop_mask), _ptr(out),
                                        _ptr(h_all), _ptr(ruc), _ptr(gsave), gbytes, _ptr(ws_buf), nbytes, _stream()),
              "decoder_fwd")
        ctx.gsave = gsave
        ctx.desc, ctx.meta = desc, (num_layers, to_len, teacher_mask, tuple(cells), len(flat))
        ctx.save_for_backward(targets, h0, p, proj_w, drop_mask, out, h_all, ruc, *flat)
        return out

    @staticmethod
    def backward(ctx, d_out):
        num_layers, to_len, teacher_mask, cells, nflat = ctx.meta
        targets, h0, p, proj_w, drop_mask, out, h_all, ruc = ctx.saved_tensors[:8]
        flat = ctx.saved_tensors[8:]
        desc = ctx.desc
        b, dev = h0.shape[1], h0.device
        d_out = d_out.contiguous()
        dh0 = torch.empty_like(h0)
        grads = [torch.empty_like(t) for t in flat]
        dpw = torch.empty_like(proj_w)
        dpb = torch.empty((desc.input_dim,), device=dev, dtype=torch.float32)
        L = _lib.lib()
        ws = _params([tuple(flat[4 * c: 4 * c + 4]) for c in cells])
        g = (CellGrads * num_layers)()
        for l, c in enumerate(cells):
            g[l] = CellGrads(*(grads[4 * c + k].data_ptr() for k in range(4)))
        nbytes = L.dcgru_decoder_bwd_workspace(C.byref(desc), num_layers, b, to_len)
        ws_buf = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        gsave = ctx.gsave
        check(L.dcgru_decoder_bwd_saved(C.byref(desc), num_layers, b, to_len, _ptr(targets), teacher_mask, _ptr(h0),
                                        _ptr(p), ws, _ptr(proj_w), _ptr(drop_mask), _ptr(out), _ptr(h_all),
                                        _ptr(ruc), _ptr(d_out), _ptr(dh0), g, _ptr(dpw), _ptr(dpb), _ptr(gsave),
                                        gsave.numel() if gsave is not None else 0, _ptr(ws_buf),
                                        nbytes, _stream()), "decoder_bwd")
        ctx.gsave = None
        return (None, dh0, None, dpw, dpb, None, None, None, None, None, None, *grads)


def decoder(targets, h0, p, proj_w, proj_b, drop_mask, desc, num_layers, to_len, teacher_mask, cells, flat):
    return _DecoderFn.apply(targets, h0, p, proj_w, proj_b, drop_mask, desc, num_layers, to_len,
                            teacher_mask, cells, *flat)
