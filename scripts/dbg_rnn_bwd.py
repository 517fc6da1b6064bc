"""Timeline of rnn_bwd_kernel (CTA 0, DCGRU_DBG=32): clock64 stamps of worker thread 0 per step.
usage: python scripts/dbg_rnn_bwd.py [M: 3|5]"""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DCGRU_DBG"] = "32"
os.environ["DCGRU_G2"] = "1"
from eeg_gnn_ssl_b200 import _lib, ops
from eeg_gnn_ssl_b200.model.cell import DCGRUCell
M = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0"); B, T, N, H, fin = 512, 12, 19, 64, 64
torch.manual_seed(0)
cell = DCGRUCell(fin, H, 2, N, filter_type="laplacian" if M == 3 else "dual_random_walk").to(dev)
x = torch.randn(T, B, N * fin, device=dev, requires_grad=True); h0 = torch.zeros(B, N * H, device=dev)
sup = [torch.softmax(torch.randn(B, N, N, device=dev), -1) for _ in range(1 if M == 3 else 2)]
P = ops.graph_poly(sup, B, N, 2)
for _ in range(2):
    hs, hl = ops.encoder_layer(x, h0, P, *cell.flat_params(), cell.desc())
    hs.square().mean().backward()
torch.cuda.synchronize()
buf = (C.c_longlong * 1024)()
_lib.check(_lib.lib().dcgru_debug_rnn_bwd_stamps(buf, 1024), "stamps")
d = np.array(buf[:], dtype=np.int64).reshape(64, 16)
names = ["wait state", "wait B2(prev)", "E1+publish", "barrier", "diffuse c terms", "diffuse u terms", "wait B1", "E2+publish", "barrier+diffuse r terms"]
for k in (3, 4, 5):
    print("step", k, {n: int(d[k, i + 1] - d[k, i]) for i, n in enumerate(names)}, "total", int(d[k + 1, 0] - d[k, 0]))
