import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ["DCGRU_DBG"] = "4"
from eeg_gnn_ssl_b200 import _lib, ops
from eeg_gnn_ssl_b200.model.cell import DCGRUCell
dev = torch.device("cuda:0"); B, T, N, H = 512, 6, 19, 64
torch.manual_seed(0)
cell = DCGRUCell(100, H, 2, N).to(dev)
x = torch.randn(T, B, N * 100, device=dev); h0 = torch.zeros(B, N * H, device=dev)
sup = [torch.softmax(torch.randn(B, N, N, device=dev), -1)]
P = ops.graph_poly(sup, B, N, 2)
desc = cell.desc(); L = _lib.lib()
nb = L.dcgru_encoder_layer_fwd_workspace(C.byref(desc), B, T)
ws = torch.zeros(nb, device=dev, dtype=torch.uint8)
hseq = torch.empty(T, B, N * H, device=dev); ruc = torch.empty(T, B, N, 3 * H, device=dev)
prm = ops._params([tuple(p.detach() for p in cell.flat_params())])
for _ in range(2):
    _lib.check(L.dcgru_encoder_layer_fwd(C.byref(desc), B, T, ops._ptr(x), x.stride(0), x.stride(1), ops._ptr(h0), ops._ptr(P), prm,
                                     ops._ptr(hseq), ops._ptr(ruc), None, 0, ops._ptr(ws), nb, ops._stream()), "fwd")
torch.cuda.synchronize()
wimg = 13 * 36864 + 8 * (24576 + 12288)
off = (wimg + 255) // 256 * 256
d = ws[off:off + 128 * 64].view(torch.int64).cpu().numpy().reshape(128, 8)
t0 = d[0, 0]
print("g  iss:start bfull afull issued | prod:start gotdone produced arrived   (cycles rel.)")
for g in range(0, 96):
    print(g, *(int(v - t0) for v in d[g]))
