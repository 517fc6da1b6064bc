"""Timeline of dw_mm_kernel's loader / MMA-issuer threads (CTA 0) at BASELINE config 2 size: DCGRU_DBG=8."""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DCGRU_DBG"] = "8"
from eeg_gnn_ssl_b200 import _lib
from eeg_gnn_ssl_b200.model.model import DCRNNEncoder
B, T, N, H, L = 512, 60, 19, 64, 1
dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = DCRNNEncoder(100, 2, H, N, L, dcgru_activation="tanh").to(dev)
sup = [torch.softmax(torch.randn(B, N, N, device=dev), -1)]
x = torch.randn(T, B, N, 100, device=dev)
h0 = torch.zeros(L, B, N * H, device=dev)
w = torch.randn(T, B, N * H, device=dev)
for _ in range(2):
    enc.zero_grad()
    _, top = enc(x, h0, sup)
    (top * w).sum().backward()
    torch.cuda.synchronize()
buf = (C.c_longlong * 2048)()
_lib.check(_lib.lib().dcgru_debug_dwmm_stamps(buf, 2048), "stamps")
d = np.array(buf[:], dtype=np.int64).reshape(256, 8)
t0 = d[0, 0]
print("i  ld:top waited issued | iss:top full accfree issued  (cycles rel.)")
for i in list(range(0, 24)) + list(range(76, 100)) + list(range(156, 170)):
    print(i, *(int(v - t0) for v in d[i, :7]))
