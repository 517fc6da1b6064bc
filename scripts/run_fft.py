"""fft_features on the BASELINE config-2 batch, a few launches (target of the ncu capture in profiles/)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eeg_gnn_ssl_b200 import ops
dev = torch.device("cuda:0")
b, n, t = 512, 19, 60
sig = torch.randn((b, n, t * 200), device=dev) * 30.0
mean, std = torch.tensor([3.924], device=dev), torch.tensor([1.560], device=dev)
ls = torch.zeros(b, device=dev)
dest = torch.arange(n, dtype=torch.int32, device=dev).repeat(b, 1)
for _ in range(8):
    x, raw = ops.fft_features(sig, mean, std, dest, ls, return_raw=True)
torch.cuda.synchronize()
print(float(x.abs().mean()), float(raw.abs().mean()))
