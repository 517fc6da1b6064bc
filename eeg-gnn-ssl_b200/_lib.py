"""ctypes binding of libdcgru_b200.so (include/dcgru_b200.h).  Fails loudly when the library is
missing or a call returns an error -- there is no fallback path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DCGRU_B200_LIB selects another build of the same library in lib/ (the jittered stress build, tests/test_gpu_stress.py)
LIB_PATH = os.path.join(_HERE, "lib", os.path.basename(os.environ.get("DCGRU_B200_LIB", "libdcgru_b200.so")))

SYMBOLS = [
    "dcgru_version", "dcgru_last_error", "dcgru_graph_poly", "dcgru_corr_supports", "dcgru_fft_features",
    "dcgru_encoder_layer_fwd", "dcgru_encoder_layer_gsave_bytes", "dcgru_encoder_layer_fwd_workspace", "dcgru_encoder_layer_bwd_workspace", "dcgru_encoder_layer_bwd",
    "dcgru_decoder_fwd_workspace", "dcgru_decoder_fwd", "dcgru_decoder_bwd_workspace", "dcgru_decoder_bwd",
    "dcgru_timing_enable", "dcgru_timing_collect", "dcgru_tc_selftest", "dcgru_debug_encoder_bwd_offsets",
    "dcgru_debug_dwmm_stamps", "dcgru_debug_dwmm_plan", "dcgru_tc_probe",
    "dcgru_clip_adam_workspace", "dcgru_clip_adam_step",
    "dcgru_debug_bulk_dp_workspace", "dcgru_debug_bulk_dp", "dcgru_debug_rnn_fwd_stamps", "dcgru_debug_rnn_bwd_stamps",
    "dcgru_cls_head_fwd", "dcgru_cls_head_bwd_workspace", "dcgru_cls_head_bwd",
    "dcgru_encoder_layer_bwd_sel_workspace", "dcgru_encoder_layer_bwd_sel",
    "dcgru_decoder_gsave_bytes", "dcgru_decoder_fwd_saved", "dcgru_decoder_bwd_saved",
]

MAX_LAYERS = 4
ACT_TANH, ACT_RELU = 0, 1


class CellDesc(C.Structure):
    _fields_ = [("num_nodes", C.c_int32), ("input_dim", C.c_int32), ("hid_dim", C.c_int32),
                ("max_diffusion_step", C.c_int32), ("num_supports", C.c_int32), ("activation", C.c_int32)]


class CellParams(C.Structure):
    _fields_ = [("Wg", C.c_void_p), ("bg", C.c_void_p), ("Wc", C.c_void_p), ("bc", C.c_void_p)]


class CellGrads(C.Structure):
    _fields_ = [("dWg", C.c_void_p), ("dbg", C.c_void_p), ("dWc", C.c_void_p), ("dbc", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"dcgru_b200: {LIB_PATH} not found -- build it with `python eeg-gnn-ssl_b200/build.py` "
            "(or __graft_entry__.build()); there is no fallback implementation")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_size_t
    pd, pp, pg = C.POINTER(CellDesc), C.POINTER(CellParams), C.POINTER(CellGrads)
    L.dcgru_version.restype = C.c_int
    L.dcgru_last_error.restype = C.c_char_p
    L.dcgru_graph_poly.argtypes = [i32, i32, i32, i32, C.POINTER(vp), C.POINTER(i64), vp, vp]
    L.dcgru_corr_supports.argtypes = [i32, i32, i32, i32, vp, i64, i64, f32, f32, i32, vp, vp, vp, vp]
    L.dcgru_fft_features.argtypes = [i32, i32, i32, i32, vp, i64, i64, vp, vp, vp, vp, i32, vp, vp, vp]
    L.dcgru_encoder_layer_fwd.argtypes = [pd, i32, i32, vp, i64, i64, vp, vp, pp, vp, vp, vp, sz, vp, sz, vp]
    L.dcgru_encoder_layer_gsave_bytes.argtypes = [pd, i32, i32]
    L.dcgru_encoder_layer_gsave_bytes.restype = sz
    L.dcgru_encoder_layer_fwd_workspace.argtypes = [pd, i32, i32]
    L.dcgru_encoder_layer_fwd_workspace.restype = sz
    L.dcgru_encoder_layer_bwd_workspace.argtypes = [pd, i32, i32]
    L.dcgru_encoder_layer_bwd_workspace.restype = sz
    L.dcgru_encoder_layer_bwd.argtypes = [pd, i32, i32, vp, i64, i64, vp, vp, pp, vp, vp, vp, vp, vp, vp, pg,
                                          vp, sz, vp, sz, vp]
    L.dcgru_decoder_fwd_workspace.argtypes = [pd, i32, i32, i32]
    L.dcgru_decoder_fwd_workspace.restype = sz
    L.dcgru_decoder_fwd.argtypes = [pd, i32, i32, i32, vp, u64, vp, vp, pp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.dcgru_decoder_bwd_workspace.argtypes = [pd, i32, i32, i32]
    L.dcgru_decoder_bwd_workspace.restype = sz
    L.dcgru_decoder_bwd.argtypes = [pd, i32, i32, i32, vp, u64, vp, vp, pp, vp, vp, vp, vp, vp, vp, vp, pg,
                                    vp, vp, vp, sz, vp]
    L.dcgru_decoder_gsave_bytes.argtypes = [pd, i32, i32, i32]
    L.dcgru_decoder_gsave_bytes.restype = sz
    L.dcgru_decoder_fwd_saved.argtypes = [pd, i32, i32, i32, vp, u64, vp, vp, pp, vp, vp, vp, vp, vp, vp, vp, sz, vp, sz, vp]
    L.dcgru_decoder_bwd_saved.argtypes = [pd, i32, i32, i32, vp, u64, vp, vp, pp, vp, vp, vp, vp, vp, vp, vp, pg,
                                          vp, vp, vp, sz, vp, sz, vp]
    L.dcgru_tc_selftest.argtypes = [vp, vp, vp, i32, i32, vp]
    L.dcgru_debug_encoder_bwd_offsets.argtypes = [pd, i32, i32, C.POINTER(sz)]
    L.dcgru_tc_probe.argtypes = [vp, i32, vp, i32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, i32, i32, i32, i32, vp, i32, vp]
    L.dcgru_debug_dwmm_stamps.argtypes = [C.POINTER(C.c_longlong), i32]
    L.dcgru_debug_rnn_fwd_stamps.argtypes = [C.POINTER(C.c_longlong), i32]
    L.dcgru_debug_rnn_bwd_stamps.argtypes = [C.POINTER(C.c_longlong), i32]
    L.dcgru_clip_adam_workspace.argtypes = [sz]
    L.dcgru_clip_adam_workspace.restype = sz
    L.dcgru_clip_adam_step.argtypes = [vp, vp, vp, vp, sz, vp, vp, f32, f32, f32, f32, f32, f32, vp, vp, sz, vp]
    L.dcgru_debug_dwmm_plan.argtypes = [i32, i32, i32, i32, C.POINTER(i32), i32]
    L.dcgru_debug_bulk_dp_workspace.argtypes = [i32, i32, i32, i32]
    L.dcgru_debug_bulk_dp_workspace.restype = sz
    L.dcgru_debug_bulk_dp.argtypes = [i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, sz, vp]
    L.dcgru_cls_head_fwd.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dcgru_cls_head_bwd_workspace.argtypes = [i32, i32, i32]
    L.dcgru_cls_head_bwd_workspace.restype = sz
    L.dcgru_cls_head_bwd.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.dcgru_encoder_layer_bwd_sel_workspace.argtypes = [pd, i32, i32]
    L.dcgru_encoder_layer_bwd_sel_workspace.restype = sz
    L.dcgru_encoder_layer_bwd_sel.argtypes = [pd, i32, i32, vp, i64, i64, vp, vp, pp, vp, vp, vp, vp, vp, vp, vp, pg,
                                              vp, sz, vp, sz, vp]
    L.dcgru_timing_enable.argtypes = [C.c_int]
    L.dcgru_timing_collect.argtypes = [C.c_char_p, sz]
    for name in SYMBOLS:
        if name not in ("dcgru_last_error", "dcgru_encoder_layer_bwd_workspace", "dcgru_encoder_layer_fwd_workspace",
                        "dcgru_encoder_layer_gsave_bytes", "dcgru_clip_adam_workspace",
                        "dcgru_decoder_fwd_workspace", "dcgru_decoder_bwd_workspace",
                        "dcgru_debug_bulk_dp_workspace", "dcgru_cls_head_bwd_workspace",
                        "dcgru_encoder_layer_bwd_sel_workspace", "dcgru_decoder_gsave_bytes"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().dcgru_last_error().decode(errors="replace")
        raise RuntimeError(f"dcgru_b200.{what} failed: {msg}")
