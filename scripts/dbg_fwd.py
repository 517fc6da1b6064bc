"""Timeline of seq_fwd_tc_kernel (CTA 0): clock64 stamps of the MMA issuer, a producer thread and the dump warp (DCGRU_DBG=4).
usage: python scripts_dbg_fwd.py [gsave: 0|1]"""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DCGRU_DBG"] = "4"
from eeg_gnn_ssl_b200 import _lib, ops
from eeg_gnn_ssl_b200.model.cell import DCGRUCell
use_g = len(sys.argv) > 1 and sys.argv[1] == "1"
dev = torch.device("cuda:0"); B, T, N, H = 512, 6, 19, 64
torch.manual_seed(0)
cell = DCGRUCell(100, H, 2, N).to(dev)
x = torch.randn(T, B, N * 100, device=dev); h0 = torch.zeros(B, N * H, device=dev)
sup = [torch.softmax(torch.randn(B, N, N, device=dev), -1)]
P = ops.graph_poly(sup, B, N, 2)
desc = cell.desc(); L = _lib.lib()
nb = L.dcgru_encoder_layer_fwd_workspace(C.byref(desc), B, T)
ws = torch.zeros(nb, device=dev, dtype=torch.uint8)
gb = L.dcgru_encoder_layer_gsave_bytes(C.byref(desc), B, T) if use_g else 0
gs = torch.empty(gb, device=dev, dtype=torch.uint8) if gb else None
hseq = torch.empty(T, B, N * H, device=dev); ruc = torch.empty(T, B, N, 3 * H, device=dev)
prm = ops._params([tuple(p.detach() for p in cell.flat_params())])
for _ in range(2):
    _lib.check(L.dcgru_encoder_layer_fwd(C.byref(desc), B, T, ops._ptr(x), x.stride(0), x.stride(1), ops._ptr(h0), ops._ptr(P), prm,
                                     ops._ptr(hseq), ops._ptr(ruc), ops._ptr(gs), gb, ops._ptr(ws), nb, ops._stream()), "fwd")
torch.cuda.synchronize()
wimg = 13 * 36864 + 8 * (24576 + 12288)
off = (wimg + 255) // 256 * 256
d = ws[off:off + 128 * 64].view(torch.int64).cpu().numpy().reshape(128, 8)
e = ws[off + 8192:off + 8192 + 128 * 32].view(torch.int64).cpu().numpy().reshape(128, 4)
t0 = d[0, 0]
print("gsave:", use_g, gb)
print("g  iss:start bfull afull issued | prod:start gotdone produced arrived | dump: afull issued read-done  (cycles rel.)")
for g in range(0, 64):
    print(g, *(int(v - t0) for v in d[g]), "|", *(int(v - t0) for v in e[g, :3]))
f = ws[off + 12288:off + 12288 + 8 * 128].view(torch.int64).cpu().numpy().reshape(8, 16)
print("epilogue stamps of thread 0, per step: [0] before wait_all_mma(gate) [1] after [2] tmem_ld done [3] sigmoid+ZH done [4] ruc stored [5] barrier passed | [6] Hc produced [7] MMAs done [8] tmem loaded [9] math+ZH done [10] stash stored [11] globals stored")
for t in range(T):
    print(t, *(int(v - t0) for v in f[t, :12]))
