"""Synchronisation stress run: small shapes through every tensor-core kernel of the path, many
iterations, results compared bit for bit with the first iteration.  Meant to be run against the jittered build:

    DCGRU_B200_LIB=libdcgru_b200_jitter.so python scripts/stress.py [iterations] [device]

(random sleeps of up to 3 us in front of one mbarrier operation in four, csrc/tc_common.cuh).  A lost arrival or an
overtaken phase traps the context (bounded spins), a slot reused too early changes the result."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eeg_gnn_ssl_b200.model.model import DCRNNModel_classification, DCRNNModel_nextTimePred  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", int(sys.argv[2]) if len(sys.argv) > 2 else 0)
torch.cuda.set_device(dev)
n, f, h, layers = 19, 100, 64, 2


def make(kind, ft, k):
    torch.manual_seed(3)
    args = types.SimpleNamespace(num_nodes=n, num_rnn_layers=layers, rnn_units=h, input_dim=f, output_dim=f,
                                 max_diffusion_step=k, dcgru_activation="tanh", filter_type=ft, dropout=0.0,
                                 cl_decay_steps=3000, use_curriculum_learning=False)
    return (DCRNNModel_classification(args, 1) if kind == "cls" else DCRNNModel_nextTimePred(args)).to(dev)


def run(model, kind, x, y, sl, sup):
    model.zero_grad(set_to_none=True)
    out = model(x, sl, sup) if kind == "cls" else model(x, y, sup)
    out.square().mean().backward()
    return [out.detach().clone()] + [p.grad.detach().clone() for p in model.parameters() if p.grad is not None]


cases = []
for kind, ft, k, b, t in (("cls", "laplacian", 2, 9, 7), ("cls", "dual_random_walk", 2, 6, 5), ("ssl", "laplacian", 2, 5, 6)):
    g = torch.Generator().manual_seed(b)
    x = torch.randn(b, t, n, f, generator=g).to(dev)
    y = torch.randn(b, 4, n, f, generator=g).to(dev)
    sl = torch.randint(1, t + 1, (b,), generator=g).to(dev)
    ns = 1 if ft == "laplacian" else 2
    sup = [torch.softmax(torch.randn(b, n, n, generator=g), -1).to(dev) for _ in range(ns)]
    cases.append((kind, make(kind, ft, k), x, y, sl, sup))

ref = [run(m, kind, x, y, sl, sup) for kind, m, x, y, sl, sup in cases]
torch.cuda.synchronize()
bad = 0
for it in range(iters):
    for ci, (kind, m, x, y, sl, sup) in enumerate(cases):
        got = run(m, kind, x, y, sl, sup)
        for a, b_ in zip(got, ref[ci]):
            if not torch.equal(a, b_):
                bad += 1
torch.cuda.synchronize()
print(f"stress: {iters} iterations x {len(cases)} cases on {dev}, lib={os.environ.get('DCGRU_B200_LIB', 'libdcgru_b200.so')}, "
      f"mismatching tensors: {bad}")
sys.exit(1 if bad else 0)
