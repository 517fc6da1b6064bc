"""Timeline of rnn_fwd_kernel (CTA 0, DCGRU_DBG=16): clock64 stamps of worker thread 0 and the MMA issuer per step.
usage: python scripts/dbg_rnn_fwd.py [M: 3|5] [fin]"""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DCGRU_DBG"] = os.environ.get("DCGRU_DBG", "16")
os.environ["DCGRU_G2"] = "1"
from eeg_gnn_ssl_b200 import _lib, ops
from eeg_gnn_ssl_b200.model.cell import DCGRUCell
M = int(sys.argv[1]) if len(sys.argv) > 1 else 3
fin = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0"); B, T, N, H = 512, 12, 19, 64
torch.manual_seed(0)
ft = "laplacian" if M == 3 else "dual_random_walk"
cell = DCGRUCell(fin, H, 2, N, filter_type=ft).to(dev)
x = torch.randn(T, B, N * fin, device=dev); h0 = torch.zeros(B, N * H, device=dev)
sup = [torch.softmax(torch.randn(B, N, N, device=dev), -1) for _ in range(1 if M == 3 else 2)]
P = ops.graph_poly(sup, B, N, 2)
with torch.no_grad():
    for _ in range(2):
        ops.encoder_layer(x, h0, P, *cell.flat_params(), cell.desc())
torch.cuda.synchronize()
buf = (C.c_longlong * 1536)()
_lib.check(_lib.lib().dcgru_debug_rnn_fwd_stamps(buf, 1536), "stamps")
d = np.array(buf[:1024], dtype=np.int64).reshape(64, 16)
e2 = np.array(buf[1024:], dtype=np.int64).reshape(64, 8)
t0 = d[0, 0]
print("step | worker: start gate-diffused gate-MMA-done epi1-done barrier cand-diffused cand-MMA-done state-published stored barrier | issuer: start xp-ready gate-issued cand-issued")
for t in range(T):
    print(t, *(int(v - t0) for v in d[t, :10]), "|", *(int(v - t0) for v in d[t, 10:14]))
print("per-step deltas (worker), step 5:", [int(d[5, i + 1] - d[5, i]) for i in range(9)], "total", int(d[6, 0] - d[5, 0]))
print("first gate term of step 5 (worker thread 0): acquire, load_pfrag, diffuse_mma16 (2 groups), fence+arrive ->",
      [int(e2[5, i + 1] - e2[5, i]) for i in range(4)])
