"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares,
argument validation fails loudly, and the drop-in modules keep the reference's structure."""
import ctypes as C
import os
import re
import types

import numpy as np
import pytest
import torch

from tests.conftest import ROOT, load_golden


def _ensure_built():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "eeg-gnn-ssl_b200", "lib", "libdcgru_b200.so")):
        g.build()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    from eeg_gnn_ssl_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dcgru_b200.h")).read()
    declared = set(re.findall(r"\b(dcgru_[a-z_0-9]+)\s*\(", hdr))
    assert {"dcgru_encoder_layer_fwd", "dcgru_encoder_layer_bwd", "dcgru_decoder_fwd", "dcgru_decoder_bwd",
            "dcgru_graph_poly", "dcgru_corr_supports"} <= declared
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert set(_lib.SYMBOLS) <= declared
    assert L.dcgru_version() >= 100


def test_argument_validation_reports_errors():
    _ensure_built()
    from eeg_gnn_ssl_b200 import _lib
    L = _lib.lib()
    bad = _lib.CellDesc(19, 100, 48, 2, 1, 0)
    assert L.dcgru_encoder_layer_bwd_workspace(C.byref(bad), 4, 4) == 0
    assert b"hid_dim" in L.dcgru_last_error()
    bad = _lib.CellDesc(33, 100, 64, 2, 1, 0)
    assert L.dcgru_decoder_bwd_workspace(C.byref(bad), 3, 4, 4) == 0
    assert b"num_nodes" in L.dcgru_last_error()
    ok = _lib.CellDesc(19, 100, 64, 2, 1, 0)
    assert L.dcgru_encoder_layer_bwd_workspace(C.byref(ok), 512, 60) > 400e6   # dA alone is 448 MB
    assert L.dcgru_decoder_bwd_workspace(C.byref(ok), 3, 4, 65) == 0            # To > 64
    rc = L.dcgru_encoder_layer_fwd(C.byref(ok), 4, 4, None, 0, 0, None, None, None, None, None, None, 0, None, 0,
                                   None)
    assert rc != 0 and b"null" in L.dcgru_last_error()
    # no device here: the tensor-core operand image is not offered (0 bytes), nothing is emulated
    assert L.dcgru_encoder_layer_gsave_bytes(C.byref(ok), 512, 60) == 0
    # input-feature kernel (N2): window length, statistics length, alignment are checked before anything is launched
    one = C.c_void_p(16)                                    # non-null, 16-byte aligned dummy pointers: never dereferenced
    assert L.dcgru_fft_features(2, 19, 4, 256, one, 19 * 1024, 1024, None, None, None, None, 0, None, one, None) != 0
    assert b"window" in L.dcgru_last_error()
    assert L.dcgru_fft_features(2, 19, 4, 200, one, 19 * 800, 800, None, None, one, one, 3, None, one, None) != 0
    assert b"stat_len" in L.dcgru_last_error()
    assert L.dcgru_fft_features(2, 19, 4, 200, one, 19 * 802, 802, None, None, None, None, 0, None, one, None) != 0
    assert b"aligned" in L.dcgru_last_error()
    assert L.dcgru_fft_features(2, 19, 4, 200, one, 19 * 800, 800, None, None, None, None, 0, None, None, None) != 0
    assert b"no output" in L.dcgru_last_error()


def test_cpu_tensors_are_rejected_not_emulated():
    from eeg_gnn_ssl_b200.model.cell import DCGRUCell
    from eeg_gnn_ssl_b200 import ops
    cell = DCGRUCell(100, 64, 2, 19)
    with pytest.raises(RuntimeError, match="CUDA"):
        cell([torch.eye(19)], torch.zeros(2, 1900), torch.zeros(2, 19 * 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.fft_features(torch.zeros(2, 19, 400))                      # input-feature kernel: no CPU path either
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.corr_supports(torch.zeros(2, 4, 19, 100))


def _args(**kw):
    base = dict(num_nodes=19, num_rnn_layers=3, rnn_units=64, input_dim=100, output_dim=100,
                max_diffusion_step=2, dcgru_activation="tanh", filter_type="laplacian", dropout=0.0,
                cl_decay_steps=3000, use_curriculum_learning=False)
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_state_dict_layout_matches_reference():
    from eeg_gnn_ssl_b200.model.model import DCRNNModel_nextTimePred
    _, a = load_golden("ssl_distance")
    m = DCRNNModel_nextTimePred(_args(rnn_units=32))
    assert list(m.state_dict().keys()) == [str(k) for k in a["state_keys"]]
    sd = m.state_dict()
    # tied decoder cells: same storage under both indices; named_parameters lists it once
    assert sd["decoder.decoding_cells.1.dconv_gate.weight"].data_ptr() == \
        sd["decoder.decoding_cells.2.dconv_gate.weight"].data_ptr()
    names = [n for n, _ in m.named_parameters()]
    assert not any(".decoding_cells.2." in n for n in names)
    for n, p in m.named_parameters():
        assert tuple(p.shape) == a["param:" + n].shape, n
    # utils.build_finetune_model re-binds these attributes (utils.py:172-174)
    enc0 = m.encoder.encoding_cells[0]
    enc0.dconv_gate, enc0.dconv_candidate = enc0.dconv_gate, enc0.dconv_candidate


@pytest.mark.parametrize("name,seed", [("ssl_distance", 21), ("enc_cfg1_distance", 123)])
def test_seed_for_seed_init_parity(name, seed):
    """same seed -> bit-identical weights as the reference constructor"""
    from eeg_gnn_ssl_b200.model.model import DCRNNModel_nextTimePred, DCRNNModel_classification
    meta, a = load_golden(name)
    keys = ("num_nodes", "num_rnn_layers", "rnn_units", "input_dim", "output_dim", "max_diffusion_step",
            "dcgru_activation", "filter_type", "dropout", "cl_decay_steps", "use_curriculum_learning")
    args = types.SimpleNamespace(**{k: meta[k] for k in keys})
    torch.manual_seed(seed)
    m = DCRNNModel_nextTimePred(args) if meta["kind"] == "ssl" else DCRNNModel_classification(args, meta["classes"])
    for n, p in m.named_parameters():
        if n.endswith("weight"):            # biases were perturbed after construction in make_golden.py
            assert np.array_equal(p.detach().numpy(), a["param:" + n]), n


@pytest.mark.skipif(not os.path.exists("/root/reference/pretrained"), reason="reference checkpoints not present")
@pytest.mark.parametrize("ckpt,ft", [("pretrained_distance_graph_12s.pth.tar", "laplacian"),
                                     ("pretrained_correlation_graph_60s.pth.tar", "dual_random_walk")])
def test_pretrained_checkpoints_load_strict(ckpt, ft):
    from eeg_gnn_ssl_b200.model.model import DCRNNModel_nextTimePred
    sd = torch.load(os.path.join("/root/reference/pretrained", ckpt), map_location="cpu",
                    weights_only=True)["model_state"]
    m = DCRNNModel_nextTimePred(_args(filter_type=ft))
    m.load_state_dict(sd, strict=True)


def test_hyphenated_directory_works_as_drop_in_root():
    """`eeg-gnn-ssl_b200/` first on sys.path makes the reference's own import lines resolve here"""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from model.model import DCRNNModel_classification, "
            "DCRNNModel_nextTimePred; from model.cell import DCGRUCell; print(DCGRUCell.__module__)"
            % os.path.join(ROOT, "eeg-gnn-ssl_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "model.cell"


@pytest.mark.parametrize("fin,batch,seq,sms", [(100, 512, 60, 148), (64, 512, 60, 148), (100, 7, 3, 148), (64, 1, 1, 132),
                                               (100, 4096, 12, 148)])
def test_weight_gradient_gemm_plan_invariants(fin, batch, seq, sms):
    """dw_mm.cu's host-side planner: every K group of the operand image is covered by exactly one tile, every set
    fits TMEM (512 columns) and two TMA boxes, the CTAs of all sets fill the device in one wave."""
    _ensure_built()
    from eeg_gnn_ssl_b200 import _lib
    L = _lib.lib()
    buf = (C.c_int32 * 256)()
    assert L.dcgru_debug_dwmm_plan(fin, batch, seq, sms, buf, 256) == 0, L.dcgru_last_error()
    v = list(buf)
    nset, nkb = v[0], v[1]
    nslab = (batch + 3) // 4 * seq
    assert nkb == nslab * 12                                 # 96 image rows per slab = 12 K blocks of 8 rows
    kgt = ((fin + 7) // 8 + 16) * 6                          # K groups per step: x chunks + 8 gate + 8 candidate chunks
    kgx = (fin + 7) // 8 * 6
    pos, covered, ctas = 2, [], 0
    for _ in range(nset):
        ntile, ncol, ncta, nbox = v[pos:pos + 4]
        pos += 4
        assert 1 <= ntile <= 4 and ncol <= 512 and 1 <= nbox <= 2 and ncta >= 1
        tcols = 0
        for j in range(ntile):
            kg0, nkg, og0, nc, tcol, soff = v[pos:pos + 6]
            pos += 6
            assert kg0 % 32 == 0 and 1 <= nkg <= 32 and nc in (64, 128, 192) and tcol == tcols
            assert soff % 4096 == 0 and soff // 16384 < nbox and (soff // 4096) % 4 in (0, 1)
            # the dA columns of the tile cover what the parts it overlaps need: x -> r|u|c, gate h -> r|u, candidate h -> c
            lo, hi = og0, og0 + nc // 4
            if kg0 < kgx:
                assert (lo, hi) == (0, 48)
            if kg0 < kgx + 48 and kg0 + nkg > kgx:
                assert lo == 0 and hi >= 32
            if kg0 + nkg > kgx + 48:
                assert lo <= 32 and hi == 48
            tcols += nc
            covered += list(range(kg0, kg0 + nkg))
        assert tcols == ncol
        ctas += ncta
    assert sorted(covered) == list(range(kgt))
    assert ctas == min(sms, max(nset, ctas)) and ctas <= sms
    if nkb >= sms:
        assert ctas == sms                                    # one full wave


def test_fused_optimizer_entry_point_validates_arguments():
    _ensure_built()
    from eeg_gnn_ssl_b200 import _lib
    L = _lib.lib()
    assert L.dcgru_clip_adam_workspace(168641) >= 8 * 42          # one partial per 4096 elements
    assert L.dcgru_clip_adam_workspace(10 ** 9) <= 8192            # capped at 512 partials
    rc = L.dcgru_clip_adam_step(None, None, None, None, 16, None, None, 0.9, 0.999, 1e-8, 0.0, 5.0, 1.0, None, None, 0, None)
    assert rc != 0 and b"null" in L.dcgru_last_error()
    buf = (C.c_float * 16)()
    step = (C.c_int32 * 1)()
    p = C.cast(buf, C.c_void_p)
    rc = L.dcgru_clip_adam_step(p, p, p, p, 16, p, C.cast(step, C.c_void_p), 1.5, 0.999, 1e-8, 0.0, 5.0, 1.0, None, p, 4096, None)
    assert rc != 0 and b"hyper" in L.dcgru_last_error()             # rejected before any launch


def test_fused_optimizer_refuses_cpu_parameters():
    from eeg_gnn_ssl_b200.optim import FusedClipAdam
    with pytest.raises(RuntimeError, match="CUDA"):
        FusedClipAdam([torch.nn.Parameter(torch.zeros(4))], lr=1e-3)
