#!/usr/bin/env python
"""Benchmark of the DCGRU training step (BASELINE.json metric: EEG clips/s, fwd+bwd, T=60, N=19).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2] [--no-extra]

N=1 workload = BASELINE.json configs[1]: distance-graph DCRNN detection, T=60, K=2, rnn_units=64,
2 layers, batch 512 per GPU, synthetic standardised FFT-like input (weak scaling: every rank owns
512 clips; N>1 adds one flat NCCL gradient all-reduce per step).

A step = zero grad, forward, loss, backward, (all-reduce,) global-norm clip, Adam step -- the body of
the reference's training loop (train.py:253-275 / train_ssl.py:163-177).

  value        : clips/s with the batch resident in HBM, whole step replayed as one CUDA graph
  value_eager  : the same step launched eagerly (what an unmodified train.py does), loss read every step
  e2e          : the step through the public module API with the batch in pinned HOST memory:
                 H2D copies of x / labels / seq_lengths / supports and the D2H read of the loss are
                 inside the timed region (train.py:246-250,269)
  roofline     : the dominant kernel's algorithmic FLOP/s (events recorded by the library around its
                 own launches on the launching stream) against the measured bf16 tensor peak
  configs      : short runs of the other BASELINE.json configs (3: correlation graph, 4: SSL encoder-
                 decoder, 5: 4-class K=3 H=128 L=3) at their per-GPU batch, same step definition
  cpu_baseline : oracle/ (torch-CPU restatement of the reference, same ATen ops) on a bounded sample

--impl reference times that CPU port alone (the reference is pure Python/PyTorch and is not on the
GPU box; the port executes the same ATen kernels, see oracle/dcgru_oracle.py).
"""
import argparse
import ctypes
import gc
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
    # key: BASELINE.json configs index (1-based)
    1: dict(name="cfg1 distance-graph detection T=12 K=2 H=64 L=2 B=4", B=4, T=12, K=2, H=64, L=2,
            filter_type="laplacian", classes=1, task="cls"),
    2: dict(name="cfg2 distance-graph detection T=60 K=2 H=64 L=2 B=512/GPU", B=512, T=60, K=2, H=64, L=2,
            filter_type="laplacian", classes=1, task="cls"),
    3: dict(name="cfg3 correlation-graph detection T=60 K=2 H=64 L=2 B=512/GPU (graph built on device)", B=512,
            T=60, K=2, H=64, L=2, filter_type="dual_random_walk", classes=1, task="cls"),
    4: dict(name="cfg4 SSL encoder(T=60)-decoder(To=12) distance graph K=2 H=64 L=3 B=512/GPU", B=512, T=60, To=12,
            K=2, H=64, L=3, filter_type="laplacian", classes=0, task="ssl"),
    5: dict(name="cfg5 correlation-graph 4-class T=12 K=3 H=128 L=3 B=1024/GPU", B=1024, T=12, K=3, H=128,
            L=3, filter_type="dual_random_walk", classes=4, task="cls"),
}
N_NODES, F_IN = 19, 100
METRIC = "EEG clips/sec (fwd+bwd, T=60, N=19)"
DTYPE = "f32 (2xFP16 split operands on tcgen05 kind::f16, 3 MMAs per product, fp32 accumulate: 22-bit significands)"


def fcell(c, h, s, k):
    """as-written forward FLOPs of one cell step for one sample"""
    m = s * k + 1
    return 2 * (s * k * 2 * N_NODES * N_NODES * c) + 2 * N_NODES * (c * m) * 3 * h


def nsup(cfg):
    return 2 if cfg["filter_type"] == "dual_random_walk" else 1


def flops_per_clip(cfg):
    """fwd+bwd = 3 x forward, as-written count (cfg2 1.302, cfg3 2.221, cfg4 2.256, cfg5 2.952 GF)"""
    s, h, k = nsup(cfg), cfg["H"], cfg["K"]
    f = sum(cfg["T"] * fcell((F_IN if l == 0 else h) + h, h, s, k) for l in range(cfg["L"]))
    if cfg["task"] == "ssl":
        f += cfg["To"] * (fcell(F_IN + h, h, s, k) + (cfg["L"] - 1) * fcell(2 * h, h, s, k) + 2 * N_NODES * h * F_IN)
    return 3 * f


def distance_supports(b):
    """scaled Laplacian of the fixed 19-electrode distance graph (utils.py:240-255)"""
    from oracle.graph_oracle import scaled_laplacian
    z = np.load(os.path.join(ROOT, "tests", "golden", "graph_supports.npz"))
    lap = torch.tensor(scaled_laplacian(z["dist_adj"]).astype(np.float32))
    return [lap.unsqueeze(0).repeat(b, 1, 1)]


def make_batch(cfg, seed):
    """-> dict of CPU tensors: x, y, (sl), (sup)"""
    g = torch.Generator().manual_seed(seed)
    b, t = cfg["B"], cfg["T"]
    out = {"x": torch.randn(b, t, N_NODES, F_IN, generator=g)}
    if cfg["task"] == "ssl":
        out["y"] = torch.randn(b, cfg["To"], N_NODES, F_IN, generator=g)
    elif cfg["classes"] == 1:
        out["y"] = (torch.rand(b, generator=g) > 0.5).float()
    else:
        out["y"] = torch.randint(0, cfg["classes"], (b,), generator=g)
    if cfg["task"] == "cls":
        out["sl"] = torch.full((b,), t, dtype=torch.long)
    if cfg["filter_type"] == "laplacian":
        out["sup"] = distance_supports(b)[0]
    # else: built on the device from the raw clip (x*std+mean) by the graph kernel
    return out


def model_args(cfg):
    import types
    return types.SimpleNamespace(num_nodes=N_NODES, num_rnn_layers=cfg["L"], rnn_units=cfg["H"], input_dim=F_IN,
                                 output_dim=F_IN, max_diffusion_step=cfg["K"], dcgru_activation="tanh",
                                 filter_type=cfg["filter_type"], dropout=0.0, cl_decay_steps=3000,
                                 use_curriculum_learning=False)


def masked_mae(pred, true):
    """utils.py:431-442 (the caller's loss of train_ssl.py:165; plain torch glue on the device)"""
    m = (true != 0.0).to(pred.dtype)
    m = m / m.mean()
    loss = (pred - true).abs() * m
    return torch.where(torch.isnan(loss), torch.zeros_like(loss), loss).mean()


def loss_of(cfg, out, y):
    if cfg["task"] == "ssl":
        return masked_mae(out, y)
    if cfg["classes"] == 1:
        return torch.nn.functional.binary_cross_entropy_with_logits(out.view(-1), y)
    return torch.nn.functional.cross_entropy(out, y)


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
def cpu_reference_leg(cfg, sample_b, steps, warmup, full_batch_once=False):
    """oracle port (torch CPU, all host threads) fwd+loss+bwd on a bounded sample of the workload.
    -> (clips/s, s/step, optional s/step of ONE full-batch step)"""
    from oracle import dcgru_oracle as O
    s = nsup(cfg)
    torch.manual_seed(123)

    def new_cell(fin):
        shp = O.param_shapes(fin, cfg["H"], cfg["K"], s)
        p = {k: torch.empty(v) for k, v in shp.items()}
        torch.nn.init.xavier_normal_(p["Wg"], gain=1.414)
        torch.nn.init.xavier_normal_(p["Wc"], gain=1.414)
        p["bg"].zero_(); p["bc"].zero_()
        return {k: v.requires_grad_(True) for k, v in p.items()}

    layers = [new_cell(F_IN if l == 0 else cfg["H"]) for l in range(cfg["L"])]
    params = [v for p in layers for v in p.values()]
    if cfg["task"] == "ssl":
        d0, d1 = new_cell(F_IN), new_cell(cfg["H"])
        dec = [d0] + [d1] * (cfg["L"] - 1)
        pw = (torch.randn(F_IN, cfg["H"]) * 0.1).requires_grad_(True)
        pb = torch.zeros(F_IN, requires_grad=True)
        params += list(d0.values()) + list(d1.values()) + [pw, pb]
    else:
        fc_w = (torch.randn(cfg["classes"], cfg["H"]) * 0.1).requires_grad_(True)
        fc_b = torch.zeros(cfg["classes"], requires_grad=True)
        params += [fc_w, fc_b]

    def batch_of(b):
        bt = make_batch(dict(cfg, B=b), 123)
        if "sup" in bt:
            sup = [bt["sup"]]
        else:
            g = torch.Generator().manual_seed(5)
            sup = [torch.softmax(torch.randn(b, N_NODES, N_NODES, generator=g), -1) for _ in range(2)]
        return bt, sup

    def one_step(bt, sup, t_len=None):
        b = bt["x"].shape[0]
        xs = bt["x"].transpose(0, 1)
        if t_len:
            xs = xs[:t_len]
        h0 = torch.zeros(cfg["L"], b, N_NODES * cfg["H"])
        t0 = time.perf_counter()
        ctx, top = O.encoder_forward(xs, h0, sup, layers, cfg["K"], N_NODES, "tanh")
        if cfg["task"] == "ssl":
            out = O.decoder_forward(bt["y"].transpose(0, 1), ctx, sup, dec, pw, pb, cfg["K"], N_NODES, "tanh")
            loss = O.masked_mae(out.reshape(cfg["To"], b, N_NODES, F_IN).transpose(0, 1), bt["y"])
        else:
            lens = torch.full((b,), xs.shape[0], dtype=torch.long)
            loss = loss_of(cfg, O.classification_head(top, lens, fc_w, fc_b, N_NODES), bt["y"])
        loss.backward()
        for v in params:
            v.grad = None
        return time.perf_counter() - t0

    bt, sup = batch_of(sample_b)
    # "all the host threads it can use": ATen's small matmuls slow down badly when oversubscribed
    # (128 threads were 19x slower than 8 on the first B200 host), so take the best thread count
    ncpu = os.cpu_count() or 1
    best_n, best_t = ncpu, None
    for n in sorted({min(ncpu, v) for v in (4, 8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(n)
        one_step(bt, sup, 2)
        dt = one_step(bt, sup, 4)
        if best_t is None or dt < best_t:
            best_n, best_t = n, dt
    torch.set_num_threads(best_n)
    times = [one_step(bt, sup) for _ in range(warmup + steps)][warmup:]
    full = None
    if full_batch_once and cfg["B"] > sample_b:
        btf, supf = batch_of(cfg["B"])
        full = one_step(btf, supf)
    return sample_b / float(np.mean(times)), float(np.mean(times)), full


def run_reference_arm(args, cfg, rank):
    if rank != 0:
        return
    sample_b = min(cfg["B"], 64)
    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    cps, sec, _ = cpu_reference_leg(cfg, sample_b, steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": cps, "unit": "clips/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": cfg["name"], "sample_batch": sample_b},
            "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"B={sample_b} of {cfg['B']}, same T/K/H/L, {steps} timed steps, fwd+loss+bwd "
                                       "(no optimiser step)"},
            "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
class Workload:
    """model + optimiser + synthetic batch of one BASELINE config on one rank"""

    def __init__(self, cfg, dev, rank, world):
        from eeg_gnn_ssl_b200 import ops
        from eeg_gnn_ssl_b200.dist import FlatGradSync, broadcast_parameters
        from eeg_gnn_ssl_b200.model.model import DCRNNModel_classification, DCRNNModel_nextTimePred
        from eeg_gnn_ssl_b200.optim import FusedClipAdam
        self.cfg, self.dev, self.world, self.ops = cfg, dev, world, ops
        torch.manual_seed(123)
        if cfg["task"] == "ssl":
            self.model = DCRNNModel_nextTimePred(model_args(cfg)).to(dev)
        else:
            self.model = DCRNNModel_classification(model_args(cfg), cfg["classes"]).to(dev)
        broadcast_parameters(self.model)
        self.model.train()
        # optimiser tail of the step (train.py:273-275): fused clip + Adam over the flat buffers (optim.cu; parity
        # with torch in tests/test_gpu_optim.py), or torch's own clip_grad_norm_ + Adam with DCGRU_FUSED_OPT=0
        self.fused_opt = os.environ.get("DCGRU_FUSED_OPT", "1") == "1"
        # data-parallel exchange: bucketed all-reduce on a side stream as backward produces the gradients; the 1/world
        # scale is folded into the fused optimiser pass (DCGRU_DP_OVERLAP=0: one collective after backward)
        self.sync = FlatGradSync(self.model.parameters(), world_size=world, align=4 if self.fused_opt else 1,
                                 overlap=os.environ.get("DCGRU_DP_OVERLAP", "1") == "1",
                                 scale_in_optimizer=self.fused_opt,
                                 bucket_counts=[4] * (cfg["L"] * (2 if cfg["task"] == "ssl" else 1)))
        if self.fused_opt:
            self.opt = FusedClipAdam(self.model.parameters(), lr=3e-4, weight_decay=5e-4, max_grad_norm=5.0,
                                     grad_sync=self.sync)
        else:
            self.opt = torch.optim.Adam(self.model.parameters(), lr=3e-4, weight_decay=5e-4, capturable=True)
        self.corr = cfg["filter_type"] != "laplacian"
        hb = make_batch(cfg, 123 + rank)
        self.host = {k: v.pin_memory() for k, v in hb.items()}
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.host.values())
        self.resident = {k: v.to(dev) for k, v in hb.items()}

    def supports_for(self, b):
        if not self.corr:
            return [b["sup"]]
        # per-clip correlation graph built on the device from the raw clip = x*std + mean
        return self.ops.corr_supports(b["x"], top_k=3, scale=1.560, shift=3.924)

    def step(self, b):
        """b: dict of device tensors"""
        self.sync.zero()
        sup = self.supports_for(b)
        if self.cfg["task"] == "ssl":
            out = self.model(b["x"], b["y"], sup)
        else:
            out = self.model(b["x"], b["sl"], sup)
        loss = loss_of(self.cfg, out, b["y"])
        loss.backward()
        self.sync.sync()
        if not self.fused_opt:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), 5.0)
        self.opt.step()
        return loss

    def empty_like_batch(self):
        return {k: torch.empty_like(v) for k, v in self.resident.items()}

    def copy_from_host(self, dst):
        for k, v in self.host.items():
            dst[k].copy_(v, non_blocking=True)


def measure(w, steps, warmup, world, kernel_steps=3, want_e2e=True, want_eager=True):
    """-> dict(ms_step, ms_eager, ms_e2e, kern{name: (count, ms)} over kernel_steps eager steps, notes)"""
    import torch.distributed as dist
    from eeg_gnn_ssl_b200 import _lib
    dev = w.dev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    res = {}
    for _ in range(warmup):
        w.step(w.resident)
    # ---- whole training step captured in a CUDA graph (removes the launch gaps; same work) ------------------
    graph, static_loss, note = None, None, "eager"
    if os.environ.get("DCGRU_BENCH_GRAPH", "1") == "1":
        try:
            barrier()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = w.step(w.resident)
            graph.replay()
            torch.cuda.synchronize()
            note = "cuda_graph(whole step)"
        except Exception as exc:                                  # fall back to eager launches
            graph, static_loss = None, None
            note = f"eager (graph capture failed: {type(exc).__name__}: {str(exc)[:160]})"
            torch.cuda.synchronize()
    res["launch"] = note

    def fast():
        if graph is not None:
            graph.replay()
        else:
            w.step(w.resident)

    for _ in range(2):
        fast()
    res["ms_step"] = timed(fast, steps) / steps
    if want_eager:
        # what the unmodified reference loop does: eager launches and loss.item() every step (train.py:269)
        # (host-bound: best of two runs, the Python launch path of a shared box is noisy)
        res["ms_eager"] = min(timed(lambda: w.step(w.resident).item(), steps) for _ in range(2)) / steps

    # ---- e2e: double-buffered input pipeline ----------------------------------------------------------------
    # Every step's batch is copied from pinned host memory (K copies for K steps, all inside the timed region) and
    # every step's loss is read back; the copy of batch k+1 runs on a copy stream while step k computes, which is
    # what a training loop with a prefetching loader does (the reference's DataLoader + .to(device), train.py:246).
    if want_e2e:
        e2e_note = "serial copy then step"
        bufs = None
        if graph is not None:
            try:
                copy_stream = torch.cuda.Stream()
                b1 = w.empty_like_batch()
                for k, v in w.resident.items():
                    b1[k].copy_(v)
                torch.cuda.synchronize()
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    l1 = w.step(b1)
                bufs, graphs2, losses2 = [w.resident, b1], [graph, g1], [static_loss, l1]
                copied = [torch.cuda.Event(), torch.cuda.Event()]
                consumed = [torch.cuda.Event(), torch.cuda.Event()]
                e2e_note = "double-buffered: H2D copy of batch k+1 on a copy stream overlaps step k"
            except Exception as exc:
                bufs = None
                e2e_note = f"serial copy then step (double buffering failed: {type(exc).__name__})"
                torch.cuda.synchronize()

        def enqueue_copy(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i])           # the step that last read this buffer has finished
                w.copy_from_host(bufs[i])
                copied[i].record(copy_stream)

        def e2e_run(n):
            """n steps; every step = (its own H2D copy) + step + loss.item()"""
            if bufs is None:
                for _ in range(n):
                    w.copy_from_host(w.resident)
                    if graph is None:
                        w.step(w.resident).item()
                    else:
                        graph.replay()
                        static_loss.item()
                return
            main = torch.cuda.current_stream()
            for i in range(2):
                consumed[i].record(main)
            enqueue_copy(0)
            for k in range(n):
                i = k & 1
                main.wait_event(copied[i])
                graphs2[i].replay()
                consumed[i].record(main)
                if k + 1 < n:
                    enqueue_copy(i ^ 1)
                losses2[i].item()                       # D2H read of this step's loss, as train.py:269 does every step

        e2e_run(2)
        res["ms_e2e"] = timed(lambda: e2e_run(steps), 1) / steps
        # the same loop over 4x the steps: the first batch's copy cannot overlap anything (pipeline fill), which costs the
        # K-step figure above one exposed 234 MB copy / K; this one shows where the loop settles
        res["ms_e2e_long"] = timed(lambda: e2e_run(4 * steps), 1) / (4 * steps)
        res["e2e_note"] = e2e_note

    # ---- per-kernel device times: a few eager steps with the library's event hooks on (same kernels as the graph)
    L = _lib.lib()
    L.dcgru_timing_enable(1)
    timed(lambda: w.step(w.resident), kernel_steps)
    buf = ctypes.create_string_buffer(1 << 16)
    _lib.check(L.dcgru_timing_collect(buf, len(buf)), "timing_collect")
    L.dcgru_timing_enable(0)
    kern = {}
    for ln in buf.value.decode().strip().splitlines():
        nm, cnt, ms = ln.split()
        kern[nm] = (int(cnt), float(ms))
    res["kern"], res["kernel_steps"] = kern, kernel_steps
    # graphs that captured NCCL work have to die before the communicator does
    del graph
    if want_e2e and bufs is not None:
        del graphs2, g1
    gc.collect()
    torch.cuda.synchronize()
    return res


def input_pipeline_leg(cfg, dev, hbm_gbs, iters=10):
    """SURVEY 8(f) N2 next to the step: raw resampled EEG (B, 19, T*200) resident in HBM -> fft_features (x and the raw
    features for the correlation graph).  HBM-bound byte work: algorithmic bytes = 800 B read + 400 B written per window
    (+400 B for the raw copy), reported against the measured copy bandwidth."""
    from eeg_gnn_ssl_b200 import ops
    b, t, n = cfg["B"], cfg["T"], 19
    g = torch.Generator(device=dev).manual_seed(5)
    sig = torch.randn((b, n, t * 200), generator=g, device=dev) * 30.0
    mean, std = torch.tensor([3.924], device=dev), torch.tensor([1.560], device=dev)      # scalar scaler, already on the device
    ls = torch.zeros(b, device=dev)
    dest = torch.arange(n, dtype=torch.int32, device=dev).repeat(b, 1)
    want_raw = nsup(cfg) == 2                    # correlation-graph configs also need the un-augmented features
    for _ in range(3):
        ops.fft_features(sig, mean, std, dest, ls, return_raw=want_raw)
    from eeg_gnn_ssl_b200 import _lib
    L = _lib.lib()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    L.dcgru_timing_enable(1)                     # the library's own events around its launches: device time of the kernel alone
    e0.record()
    for _ in range(iters):
        ops.fft_features(sig, mean, std, dest, ls, return_raw=want_raw)
    e1.record()
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 14)
    _lib.check(L.dcgru_timing_collect(buf, len(buf)), "timing_collect")
    L.dcgru_timing_enable(0)
    kern_ms = None
    for ln in buf.value.decode().strip().splitlines():
        nm, cnt, tot = ln.split()
        if nm == "fft_features":
            kern_ms = float(tot) / int(cnt)
    ms_call = e0.elapsed_time(e1) / iters         # back-to-back calls through ops.fft_features (allocations + launch included)
    ms = kern_ms if kern_ms else ms_call
    nbytes = b * n * t * (800 + 400 + (400 if want_raw else 0))
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"kernel": "fft_features", "what": "raw EEG -> per-second log-amplitude FFT + reflect/scale augmentation + "
            "standardisation (DataLoader work of data/dataloader_detection.py:58-72,233-256,382-393 on the device)",
            "ms_per_batch": ms, "ms_per_call_through_ops": ms_call, "windows": b * n * t, "algorithmic_bytes": nbytes,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs / hbm_gbs},
            "l2": "signal + outputs = %.2f GB per launch (larger than L2)" % (nbytes / 1e9)}


def kernel_families(cfg):
    """algorithmic FLOPs per launch of each kernel family (as-written count, SURVEY 8(d) / DESIGN.md section 5):
       forward layer l            : T*B*F_cell(C_l)
       BPTT layer l (dH/dA side)  : the recurrent (h) columns of F_cell;  dX kernel: the input (x) columns
       weight gradient layer l    : T*B*F_cell(C_l)"""
    s_, m_ = nsup(cfg), nsup(cfg) * cfg["K"] + 1
    tb = cfg["T"] * cfg["B"]

    def f_cols(cols):  # diffusion of `cols` columns twice-as-written + projection of cols*M rows onto 3H outputs
        return 2 * (s_ * cfg["K"] * 2 * N_NODES * N_NODES * cols) + 2 * N_NODES * (cols * m_) * 3 * cfg["H"]
    fam = {}
    for l in range(cfg["L"]):
        fin_l = F_IN if l == 0 else cfg["H"]
        full, hpart, xpart = tb * f_cols(fin_l + cfg["H"]), tb * f_cols(cfg["H"]), tb * f_cols(fin_l)
        for nm in ("seq_fwd", "seq_fwd_tc", "dw", "dw_tc", "dw_mm", "dw_mm16"):
            fam.setdefault(nm, []).append(full)
        fam.setdefault("seq_bwd", []).append(full if l > 0 else hpart)
        for nm in ("seq_bwd_tc", "rnn_fwd", "rnn_bwd"):              # recurrent (h) columns only
            fam.setdefault(nm, []).append(hpart)
        fam.setdefault("xproj", []).append(xpart)                    # hoisted x-part of the forward
        if l > 0:
            for nm in ("dx_tc", "dx", "dx16"):
                fam.setdefault(nm, []).append(xpart)
    return fam


def load_json(path, default=None):
    try:
        return json.load(open(path))
    except Exception:
        return default


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other BASELINE configs")
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the CUDA path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {})
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    prof = load_json(os.path.join(ROOT, "profiles", "traffic_r02b.json")) or \
        load_json(os.path.join(ROOT, "profiles", "traffic_r02.json"), {})

    with ClockSampler(local) as clk:
        w = Workload(cfg, dev, rank, world)
        r = measure(w, args.steps, args.warmup, world)
        grad_bytes, h2d, fused_opt = w.sync.nbytes, w.h2d_bytes, w.fused_opt
        del w
        gc.collect()
        torch.cuda.empty_cache()
    clips = cfg["B"] * world
    value = clips / (r["ms_step"] * 1e-3)

    # ---- per-kernel device time -> roofline of the dominant kernel ----------------------------------
    kern, ksteps = r["kern"], r["kernel_steps"]
    launches = sum(c for c, _ in kern.values()) // ksteps
    fam = kernel_families(cfg)
    big = [k for k in fam if k in kern]
    dom = max(big, key=lambda k: kern[k][1])
    dcount, dms = kern[dom]
    flops_per_launch = sum(fam[dom]) / len(fam[dom])
    achieved = flops_per_launch / (dms / dcount * 1e-3) / 1e12
    traffic = (prof.get(dom) if isinstance(prof, dict) else None)
    tensor_pct = (prof.get("tensor_pipe_active_pct", {}) or {}).get(dom) if isinstance(prof, dict) else None
    total_alg = sum(sum(fam[k]) for k in big)
    total_kms = sum(kern[k][1] for k in big) / ksteps
    step_tflops = flops_per_clip(cfg) * cfg["B"] / (r["ms_step"] * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback",
                "frac_of_split_attainable": achieved / (peak / 3.0),
                "tensor_pipe_active_pct": tensor_pct,
                "profile_source": "profiles/ncu_*_r02b.txt (ncu --set full --clock-control none on `bench.py --config 2 --steps 2 "
                                  "--no-extra --no-cpu-baseline`; see profiles/README.md for commit and command)",
                "note": "kernels compute in fp32-equivalent 2xFP16 on tcgen05 (3 kind::f16 MMAs per product at the bf16 rate: "
                        "the attainable peak of this arithmetic is 1/3 of the bf16 peak); FLOPs are the as-written count of "
                        "SURVEY 8(d); H=128 cells run on the fp32 FMA kernels",
                "all_kernels_tflops": total_alg / (total_kms * 1e-3) / 1e12,
                "whole_step_tflops": step_tflops, "whole_step_frac": step_tflops / peak}

    line = {"metric": METRIC, "value": value, "unit": "clips/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": cfg["name"], "global_batch": clips, "seq_len": cfg["T"],
                       "parallelism": f"dp{world}", "l2": "inputs larger than L2 (x = 233 MB/rank, saved "
                       "activations and operand images, several GB/rank, are rewritten every step)",
                       "step": "zero_grad+fwd+loss+bwd+allreduce+clip+adam", "grad_allreduce_bytes": grad_bytes,
                       "launch": r["launch"], "e2e_pipeline": r.get("e2e_note"),
                       "optimizer": "fused clip+Adam (optim.cu)" if fused_opt else "torch clip_grad_norm_ + Adam"},
            "value_eager": clips / (r["ms_eager"] * 1e-3),
            "value_eager_note": "same step, eager launches + loss.item() every step (no CUDA graph): what an "
                                "unmodified train.py loop gets",
            "e2e": {"value": clips / (r["ms_e2e"] * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": r["ms_e2e"],
                    "value_over_4x_steps": clips / (r["ms_e2e_long"] * 1e-3),
                    "note": "K steps incl. the un-overlappable copy of the first batch (pipeline fill); value_over_4x_steps = the "
                            "same loop, same copies and read-backs per step, over 4K steps"},
            "gpu_launches": launches, "roofline": roofline,
            "kernel_ms_per_step": {k: v[1] / ksteps for k, v in kern.items()},
            "clocks": clk.summary()}

    # ---- the other BASELINE configs: short runs, same step definition --------------------------------------------
    if not args.no_extra:
        extra = {}
        for ci in (3, 4, 5):
            if ci == args.config:
                continue
            c2 = CONFIGS[ci]
            try:
                w2 = Workload(c2, dev, rank, world)
                r2 = measure(w2, max(3, min(args.steps, 5)), 3, world, kernel_steps=2, want_e2e=False)
                del w2
                gc.collect()
                torch.cuda.empty_cache()
                k2 = r2["kern"]
                tot = sum(v[1] for v in k2.values()) or 1.0
                d2 = max(k2, key=lambda k: k2[k][1])
                tf = flops_per_clip(c2) * c2["B"] / (r2["ms_step"] * 1e-3) / 1e12
                extra[f"cfg{ci}"] = {
                    "workload": c2["name"], "value": c2["B"] * world / (r2["ms_step"] * 1e-3), "unit": "clips/s",
                    "ms_per_step": r2["ms_step"], "value_eager": c2["B"] * world / (r2["ms_eager"] * 1e-3),
                    "launch": r2["launch"], "gflop_per_clip": flops_per_clip(c2) / 1e9,
                    "whole_step_tflops": tf, "whole_step_frac": tf / peak,
                    "dominant_kernel": d2, "dominant_kernel_share": k2[d2][1] / tot,
                    "kernel_ms_per_step": {k: v[1] / r2["kernel_steps"] for k, v in k2.items()}}
            except Exception as exc:
                extra[f"cfg{ci}"] = {"workload": c2["name"], "error": f"{type(exc).__name__}: {str(exc)[:200]}"}
                torch.cuda.synchronize()
        line["configs"] = extra

    if not args.no_extra:
        try:
            line["input_pipeline"] = input_pipeline_leg(cfg, dev, peaks.get("hbm_gbs", 6500.0))
        except Exception as exc:
            line["input_pipeline"] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
            torch.cuda.synchronize()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cps, sec, full = cpu_reference_leg(cfg, min(cfg["B"], 64), 3, 1, full_batch_once=True)
        line["cpu_baseline"] = {"value": cps, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"B={min(cfg['B'], 64)} of {cfg['B']}, same T/K/H/L, 3 timed steps "
                                          f"({sec:.2f} s/step), fwd+loss+bwd without optimiser step",
                                "full_batch_one_step": None if full is None else
                                {"B": cfg["B"], "s_per_step": full, "clips_per_s": cfg["B"] / full}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # ordered teardown: every captured graph is already destroyed (measure() drops them before returning)
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        guard = threading.Timer(30.0, lambda: os._exit(0))     # a communicator that refuses to die must not hang the job
        guard.daemon = True
        guard.start()
        dist.destroy_process_group()
        guard.cancel()


if __name__ == "__main__":
    main()
