"""Host->device copy bandwidth of the GPU box (pinned memory), alone and under a running training step.

The e2e number of bench.py copies 234 MB per step (config 2); this tells how much of the e2e step is the PCIe link.
  python scripts/h2d_probe.py
"""
import subprocess
import sys
import time

import torch


def bw(nbytes, reps=5, stream=None):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s = stream or torch.cuda.current_stream()
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(s)
            dev.copy_(host, non_blocking=True)
            e1.record(s)
        torch.cuda.synchronize()
        best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def main():
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,"
                            "pcie.link.width.max", "--format=csv"], capture_output=True, text=True, timeout=30).stdout
        print(q.strip().replace("\n", " | "))
    except Exception as exc:  # noqa: BLE001
        print("nvidia-smi:", exc)
    for mb in (16, 64, 234, 1024):
        print(f"H2D pinned {mb:5d} MB: {bw(mb << 20):6.1f} GB/s")
    # split copy on two streams (two copy engines)
    n = 234 << 20
    h = [torch.empty(n // 2, dtype=torch.uint8).pin_memory() for _ in range(2)]
    d = [torch.empty(n // 2, dtype=torch.uint8, device="cuda") for _ in range(2)]
    ss = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        for i in range(2):
            with torch.cuda.stream(ss[i]):
                d[i].copy_(h[i], non_blocking=True)
    torch.cuda.synchronize()
    print(f"H2D 2 streams x 117 MB: {5 * n / (time.perf_counter() - t0) / 1e9:6.1f} GB/s")


if __name__ == "__main__":
    sys.exit(main())
