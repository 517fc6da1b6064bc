#!/usr/bin/env python
"""Benchmark of the DCGRU training step (BASELINE.json metric: EEG clips/s, fwd+bwd, T=60, N=19).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2]

N=1 workload = BASELINE.json configs[1]: distance-graph DCRNN detection, T=60, K=2, rnn_units=64,
2 layers, batch 512 per GPU, synthetic standardised FFT-like input (weak scaling: every rank owns
512 clips; N>1 adds one flat NCCL gradient all-reduce per step).

A step = zero grad, forward, BCE-with-logits loss, backward, (all-reduce,) global-norm clip, Adam
step -- the body of the reference's training loop (train.py:253-275).

  value : clips/s with the batch resident in HBM (CUDA events, max over ranks)
  e2e   : the same step through the public module API with the batch in pinned HOST memory:
          H2D copies of x / labels / seq_lengths / supports and the D2H read of the loss are
          inside the timed region (train.py:246-250,269)
  roofline : the dominant kernel's algorithmic FLOP/s (events recorded by the library around its
          own launches on the launching stream) against the measured bf16 tensor peak
  cpu_baseline : oracle/ (torch-CPU restatement of the reference, same ATen ops) on a bounded sample

--impl reference times that CPU port alone (the reference is pure Python/PyTorch and is not on the
GPU box; the port executes the same ATen kernels, see oracle/dcgru_oracle.py).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # key: BASELINE.json configs index
    1: dict(name="cfg1 distance-graph detection T=12 K=2 H=64 L=2 B=4", B=4, T=12, K=2, H=64, L=2,
            filter_type="laplacian", classes=1),
    2: dict(name="cfg2 distance-graph detection T=60 K=2 H=64 L=2 B=512/GPU", B=512, T=60, K=2, H=64, L=2,
            filter_type="laplacian", classes=1),
    3: dict(name="cfg3 correlation-graph detection T=60 K=2 H=64 L=2 B=512/GPU", B=512, T=60, K=2, H=64, L=2,
            filter_type="dual_random_walk", classes=1),
    5: dict(name="cfg5 correlation-graph 4-class T=12 K=3 H=128 L=3 B=1024/GPU", B=1024, T=12, K=3, H=128,
            L=3, filter_type="dual_random_walk", classes=4),
}
N_NODES, F_IN = 19, 100


def fcell(c, h, s, k):
    """as-written forward FLOPs of one cell step for one sample"""
    m = s * k + 1
    return 2 * (s * k * 2 * N_NODES * N_NODES * c) + 2 * N_NODES * (c * m) * 3 * h


def flops_per_clip_fwd(cfg):
    s = 2 if cfg["filter_type"] == "dual_random_walk" else 1
    per_layer = [cfg["T"] * fcell((F_IN if l == 0 else cfg["H"]) + cfg["H"], cfg["H"], s, cfg["K"])
                 for l in range(cfg["L"])]
    return per_layer


def distance_supports(b):
    """scaled Laplacian of the fixed 19-electrode distance graph (utils.py:240-255)"""
    from oracle.graph_oracle import scaled_laplacian
    z = np.load(os.path.join(ROOT, "tests", "golden", "graph_supports.npz"))
    lap = torch.tensor(scaled_laplacian(z["dist_adj"]).astype(np.float32))
    return [lap.unsqueeze(0).repeat(b, 1, 1)]


def make_batch(cfg, seed):
    g = torch.Generator().manual_seed(seed)
    b, t = cfg["B"], cfg["T"]
    x = torch.randn(b, t, N_NODES, F_IN, generator=g)
    if cfg["classes"] == 1:
        y = (torch.rand(b, generator=g) > 0.5).float()
    else:
        y = torch.randint(0, cfg["classes"], (b,), generator=g)
    sl = torch.full((b,), t, dtype=torch.long)
    if cfg["filter_type"] == "laplacian":
        sup = distance_supports(b)
    else:
        sup = None        # built on the device from the raw clip (x*std+mean) by the graph kernel
    return x, y, sl, sup


def model_args(cfg):
    import types
    return types.SimpleNamespace(num_nodes=N_NODES, num_rnn_layers=cfg["L"], rnn_units=cfg["H"], input_dim=F_IN,
                                 output_dim=F_IN, max_diffusion_step=cfg["K"], dcgru_activation="tanh",
                                 filter_type=cfg["filter_type"], dropout=0.0, cl_decay_steps=3000,
                                 use_curriculum_learning=False)


def loss_of(cfg, logits, y):
    if cfg["classes"] == 1:
        return torch.nn.functional.binary_cross_entropy_with_logits(logits.view(-1), y)
    return torch.nn.functional.cross_entropy(logits, y)


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, gpu_index):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self.gpu = gpu_index
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
def cpu_reference_leg(cfg, sample_b, steps, warmup):
    """oracle port (torch CPU, all host threads) fwd+loss+bwd on a bounded sample of the workload"""
    from oracle import dcgru_oracle as O
    c = dict(cfg, B=sample_b)
    x, y, sl, sup = make_batch(c, 123)
    s = 2 if cfg["filter_type"] == "dual_random_walk" else 1
    if sup is None:
        g = torch.Generator().manual_seed(5)
        sup = [torch.softmax(torch.randn(sample_b, N_NODES, N_NODES, generator=g), -1) for _ in range(2)]
    torch.manual_seed(123)
    layers = []
    for l in range(cfg["L"]):
        shp = O.param_shapes(F_IN if l == 0 else cfg["H"], cfg["H"], cfg["K"], s)
        p = {k: torch.empty(v) for k, v in shp.items()}
        torch.nn.init.xavier_normal_(p["Wg"], gain=1.414)
        torch.nn.init.xavier_normal_(p["Wc"], gain=1.414)
        p["bg"].zero_(); p["bc"].zero_()
        layers.append({k: v.requires_grad_(True) for k, v in p.items()})
    fc_w = (torch.randn(cfg["classes"], cfg["H"]) * 0.1).requires_grad_(True)
    fc_b = torch.zeros(cfg["classes"], requires_grad=True)
    h0 = torch.zeros(cfg["L"], sample_b, N_NODES * cfg["H"])
    xs = x.transpose(0, 1)
    def one_step(xseq):
        t0 = time.perf_counter()
        _, top = O.encoder_forward(xseq, h0, sup, layers, cfg["K"], N_NODES, "tanh")
        lens = torch.full((sample_b,), xseq.shape[0], dtype=torch.long)
        logits = O.classification_head(top, lens, fc_w, fc_b, N_NODES)
        loss = loss_of(cfg, logits, y)
        loss.backward()
        for p in layers:
            for v in p.values():
                v.grad = None
        return time.perf_counter() - t0

    # "all the host threads it can use": ATen's small matmuls slow down badly when oversubscribed
    # (128 threads were 19x slower than 8 on the first B200 host), so take the best thread count
    ncpu = os.cpu_count() or 1
    best_n, best_t = ncpu, None
    for n in sorted({min(ncpu, v) for v in (4, 8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(n)
        one_step(xs[:2])
        dt = one_step(xs[:4])
        if best_t is None or dt < best_t:
            best_n, best_t = n, dt
    torch.set_num_threads(best_n)
    times = [one_step(xs) for _ in range(warmup + steps)][warmup:]
    return sample_b / float(np.mean(times)), float(np.mean(times))


def run_reference_arm(args, cfg, rank):
    if rank != 0:
        return
    sample_b = min(cfg["B"], 64)
    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    cps, sec = cpu_reference_leg(cfg, sample_b, steps, warm)
    line = {"impl": "reference", "metric": "EEG clips/sec (fwd+bwd, T=60, N=19)", "value": cps, "unit": "clips/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": cfg["name"], "sample_batch": sample_b},
            "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"B={sample_b} of {cfg['B']}, same T/K/H/L, {steps} timed steps"},
            "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, cfg, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    from eeg_gnn_ssl_b200 import _lib, ops
    from eeg_gnn_ssl_b200.dist import FlatGradSync, broadcast_parameters
    from eeg_gnn_ssl_b200.model.model import DCRNNModel_classification

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the CUDA path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(123)
    model = DCRNNModel_classification(model_args(cfg), cfg["classes"]).to(dev)
    broadcast_parameters(model)
    model.train()
    # optimiser tail of the step (train.py:273-275): fused clip + Adam over the flat buffers (optim.cu; parity with
    # torch in tests/test_gpu_optim.py), or torch's own clip_grad_norm_ + Adam with DCGRU_FUSED_OPT=0
    fused_opt = os.environ.get("DCGRU_FUSED_OPT", "1") == "1"
    sync = FlatGradSync(model.parameters(), world_size=world, align=4 if fused_opt else 1)
    if fused_opt:
        from eeg_gnn_ssl_b200.optim import FusedClipAdam
        opt = FusedClipAdam(model.parameters(), lr=3e-4, weight_decay=5e-4, max_grad_norm=5.0, grad_sync=sync)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=3e-4, weight_decay=5e-4, capturable=True)

    x, y, sl, sup = make_batch(cfg, 123 + rank)
    corr = sup is None
    # ---- device-resident batch ("value") and pinned host batch ("e2e") -----------------------------
    host = {"x": x.pin_memory(), "y": y.pin_memory(), "sl": sl.pin_memory()}
    if not corr:
        host["sup"] = sup[0].pin_memory()
    d_x, d_y, d_sl = x.to(dev), y.to(dev), sl.to(dev)
    d_sup = [sup[0].to(dev)] if not corr else None

    def supports_for(xd):
        if not corr:
            return d_sup
        # per-clip correlation graph built on the device from the raw clip = x*std + mean
        return ops.corr_supports(xd, top_k=3, scale=1.560, shift=3.924)

    def step(xd, yd, sld, supd):
        sync.zero()
        logits = model(xd, sld, supd)
        loss = loss_of(cfg, logits, yd)
        loss.backward()
        sync.sync()
        if not fused_opt:
            torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def resident_step():
        step(d_x, d_y, d_sl, supports_for(d_x))

    h2d = sum(t.numel() * t.element_size() for t in host.values())

    def e2e_step():
        xd = host["x"].to(dev, non_blocking=True)
        yd = host["y"].to(dev, non_blocking=True)
        sld = host["sl"].to(dev, non_blocking=True)
        supd = [host["sup"].to(dev, non_blocking=True)] if not corr else supports_for(xd)
        loss = step(xd, yd, sld, supd)
        return loss.item()                      # D2H read, as train.py:269 does every step

    for _ in range(args.warmup):
        resident_step()
    # ---- whole training step captured in a CUDA graph (removes ~200 launch gaps per step; same work) ----------
    graph, static_loss, graph_note = None, None, "eager"
    # (N > 1: the flat-gradient NCCL all-reduce is captured inside the graph too; DCGRU_BENCH_GRAPH=0 -> eager)
    if os.environ.get("DCGRU_BENCH_GRAPH", "1") == "1":
        try:
            barrier()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = step(d_x, d_y, d_sl, supports_for(d_x))
            graph.replay()
            torch.cuda.synchronize()
            graph_note = "cuda_graph(whole step)"
        except Exception as exc:                                  # fall back to eager launches
            graph, static_loss = None, None
            graph_note = f"eager (graph capture failed: {type(exc).__name__}: {str(exc)[:160]})"
            torch.cuda.synchronize()

    def resident_fast():
        if graph is not None:
            graph.replay()
        else:
            resident_step()

    # ---- e2e: double-buffered input pipeline ------------------------------------------------------------------
    # Every step's batch is copied from pinned host memory (K copies for K steps, all inside the timed region) and
    # every step's loss is read back; the copy of batch k+1 runs on a copy stream while step k computes, which is
    # what a training loop with a prefetching loader does (the reference's DataLoader + .to(device), train.py:246).
    e2e_note = "serial copy then step"
    bufs, graphs2, losses2 = None, None, None
    if graph is not None:
        try:
            copy_stream = torch.cuda.Stream()
            bufs = [dict(x=d_x, y=d_y, sl=d_sl, sup=d_sup),
                    dict(x=torch.empty_like(d_x), y=torch.empty_like(d_y), sl=torch.empty_like(d_sl),
                         sup=[torch.empty_like(d_sup[0])] if not corr else None)]
            bufs[1]["x"].copy_(d_x); bufs[1]["y"].copy_(d_y); bufs[1]["sl"].copy_(d_sl)
            if not corr:
                bufs[1]["sup"][0].copy_(d_sup[0])
            torch.cuda.synchronize()
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                l1 = step(bufs[1]["x"], bufs[1]["y"], bufs[1]["sl"],
                          bufs[1]["sup"] if not corr else supports_for(bufs[1]["x"]))
            graphs2, losses2 = [graph, g1], [static_loss, l1]
            copied = [torch.cuda.Event(), torch.cuda.Event()]
            consumed = [torch.cuda.Event(), torch.cuda.Event()]
            e2e_note = "double-buffered: H2D copy of batch k+1 on a copy stream overlaps step k"
        except Exception as exc:
            bufs = None
            e2e_note = f"serial copy then step (double buffering failed: {type(exc).__name__})"
            torch.cuda.synchronize()

    def enqueue_copy(i):
        b = bufs[i]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i])           # the step that last read this buffer has finished
            b["x"].copy_(host["x"], non_blocking=True)
            b["y"].copy_(host["y"], non_blocking=True)
            b["sl"].copy_(host["sl"], non_blocking=True)
            if not corr:
                b["sup"][0].copy_(host["sup"], non_blocking=True)
            copied[i].record(copy_stream)

    def e2e_run(steps):
        """K steps: returns nothing; every step = (its own H2D copy) + graph + loss.item()"""
        if bufs is None:
            for _ in range(steps):
                if graph is None:
                    e2e_step()
                else:
                    d_x.copy_(host["x"], non_blocking=True)
                    d_y.copy_(host["y"], non_blocking=True)
                    d_sl.copy_(host["sl"], non_blocking=True)
                    if not corr:
                        d_sup[0].copy_(host["sup"], non_blocking=True)
                    graph.replay()
                    static_loss.item()
            return
        main = torch.cuda.current_stream()
        for i in range(2):
            consumed[i].record(main)
        enqueue_copy(0)
        for k in range(steps):
            i = k & 1
            main.wait_event(copied[i])
            graphs2[i].replay()
            consumed[i].record(main)
            if k + 1 < steps:
                enqueue_copy(i ^ 1)
            losses2[i].item()                           # D2H read of this step's loss, as train.py:269 does every step

    L = _lib.lib()
    with ClockSampler(local) as clk:
        for _ in range(2):
            resident_fast()
        total_ms = timed(resident_fast, args.steps)
        e2e_run(2)
        e2e_ms = timed(lambda: e2e_run(args.steps), 1)
        # per-kernel device times: a few eager steps with the library's event hooks on (same kernels as the graph)
        L.dcgru_timing_enable(1)
        ksteps = min(args.steps, 3)
        timed(resident_step, ksteps)
        buf = ctypes.create_string_buffer(1 << 16)
        _lib.check(L.dcgru_timing_collect(buf, len(buf)), "timing_collect")
        L.dcgru_timing_enable(0)
    ms_step = total_ms / args.steps
    clips = cfg["B"] * world
    value = clips / (ms_step * 1e-3)
    e2e_value = clips / (e2e_ms / args.steps * 1e-3)

    # ---- per-kernel device time -> roofline of the dominant kernel ----------------------------------
    kern = {}
    for ln in buf.value.decode().strip().splitlines():
        nm, cnt, ms = ln.split()
        kern[nm] = (int(cnt), float(ms))
    launches = sum(c for c, _ in kern.values()) // ksteps
    # algorithmic FLOPs per launch of each kernel family (as-written count, SURVEY 8(d) / DESIGN.md section 5):
    #   forward layer l            : T*B*F_cell(C_l)
    #   BPTT layer l (dH/dA side)  : the recurrent (h) columns of F_cell;  dX kernel: the input (x) columns
    #   weight gradient layer l    : T*B*F_cell(C_l)
    s_ = 2 if cfg["filter_type"] == "dual_random_walk" else 1
    m_ = s_ * cfg["K"] + 1
    tb = cfg["T"] * cfg["B"]
    def f_cols(cols):          # diffusion of `cols` columns twice-as-written + projection of cols*M rows onto 3H outputs
        return 2 * (s_ * cfg["K"] * 2 * N_NODES * N_NODES * cols) + 2 * N_NODES * (cols * m_) * 3 * cfg["H"]
    fam = {}
    for l in range(cfg["L"]):
        fin_l = F_IN if l == 0 else cfg["H"]
        full, hpart, xpart = tb * f_cols(fin_l + cfg["H"]), tb * f_cols(cfg["H"]), tb * f_cols(fin_l)
        for nm in ("seq_fwd", "seq_fwd_tc", "dw", "dw_tc", "dw_mm"):
            fam.setdefault(nm, []).append(full)
        fam.setdefault("seq_bwd", []).append(full if l > 0 else hpart)
        fam.setdefault("seq_bwd_tc", []).append(hpart)
        if l > 0:
            fam.setdefault("dx_tc", []).append(xpart)
            fam.setdefault("dx", []).append(xpart)
    big = [k for k in fam if k in kern]
    dom = max(big, key=lambda k: kern[k][1])
    dcount, dms = kern[dom]
    flops_per_launch = sum(fam[dom]) / len(fam[dom])
    achieved = flops_per_launch / (dms / dcount * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    traffic = None
    try:                                   # dram bytes per launch of the dominant kernel, from the committed ncu capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json")))
        traffic = tr.get(dom)
    except Exception:
        pass
    total_alg = sum(sum(fam[k]) for k in big)
    total_ms = sum(kern[k][1] for k in big) / ksteps
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback",
                "note": "kernels compute in fp32-equivalent 3xTF32 on tcgen05 (3 TF32 MMAs per product: the attainable "
                        "peak of this arithmetic is ~1/6 of the bf16 peak); FLOPs are the as-written count of SURVEY 8(d)",
                "all_kernels_tflops": total_alg / (total_ms * 1e-3) / 1e12,
                "kernel_ms_per_step": {k: v[1] / ksteps for k, v in kern.items()}}

    line = {"metric": "EEG clips/sec (fwd+bwd, T=60, N=19)", "value": value, "unit": "clips/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (3xTF32 on tcgen05, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": cfg["name"], "global_batch": clips, "seq_len": cfg["T"],
                       "parallelism": f"dp{world}", "l2": "inputs larger than L2 (x = 233 MB/rank, saved "
                       "activations ~1.2 GB/rank are rewritten every step)",
                       "step": "zero_grad+fwd+loss+bwd+allreduce+clip+adam", "grad_allreduce_bytes": sync.nbytes,
                       "launch": graph_note, "e2e_pipeline": e2e_note,
                       "optimizer": "fused clip+Adam (optim.cu)" if fused_opt else "torch clip_grad_norm_ + Adam"},
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "roofline": roofline, "clocks": clk.summary()}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cps, sec = cpu_reference_leg(cfg, min(cfg["B"], 64), 3, 1)
        line["cpu_baseline"] = {"value": cps, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"B={min(cfg['B'], 64)} of {cfg['B']}, same T/K/H/L, 3 timed steps "
                                          f"({sec:.2f} s/step)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that captured NCCL work have to die before the communicator does, and tearing either down can
        # block; the numbers are out, so leave without running destructors (every rank, after a last rendezvous)
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
