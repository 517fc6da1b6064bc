"""Quick device-time probe of the encoder fwd/bwd kernels at BASELINE config 2 (no pinned memory, no
nvidia-smi sampling): prints per-kernel ms from the library's event hooks."""
import ctypes, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eeg_gnn_ssl_b200 import _lib
from eeg_gnn_ssl_b200.model.model import DCRNNEncoder
from oracle.graph_oracle import scaled_laplacian
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T, N, H, L = 60, 19, 64, 2
dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = DCRNNEncoder(100, 2, H, N, L, dcgru_activation="tanh").to(dev)
z = np.load("tests/golden/graph_supports.npz")
lap = torch.tensor(scaled_laplacian(z["dist_adj"]).astype(np.float32), device=dev)
sup = [lap.unsqueeze(0).repeat(B, 1, 1)]
x = torch.randn(T, B, N, 100, device=dev)
h0 = torch.zeros(L, B, N * H, device=dev)
w = torch.randn(T, B, N * H, device=dev)
lib = _lib.lib()
for it in range(3):
    if it == 2:
        lib.dcgru_timing_enable(1)
    enc.zero_grad()
    _, top = enc(x, h0, sup)
    (top * w).sum().backward()
    torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 16)
_lib.check(lib.dcgru_timing_collect(buf, len(buf)), "collect")
print(buf.value.decode())
