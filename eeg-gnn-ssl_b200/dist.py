"""Batch-sharded data parallelism for the DCGRU models: one process per GPU, weights replicated, the parameter
gradient exchanged through ONE flat fp32 buffer per step.

The reference is single-process; this is the B200 addition.  Clips are independent, so the only
exchange is the parameter gradient: every ``p.grad`` is a view into one flat buffer and the backward kernels'
results are accumulated into it by autograd.

Overlap: the buffer is cut into buckets in the order backward produces them (top layer / head first, layer 0
last).  As soon as every gradient of a bucket has been accumulated (``register_post_accumulate_grad_hook``) its
``all_reduce(SUM)`` is issued on a side stream that waits on an event of the compute stream, so the exchange of
the upper layers runs over NVLink while the lower layers' BPTT and weight-gradient kernels still compute; only the
last bucket (layer 0) is exposed.  ``sync()`` makes the compute stream wait for the side stream.  The 1/world
scale is not a separate pass: ``FusedClipAdam`` applies it inside its clip+Adam kernel (``pending_scale()``);
with another optimiser ``sync()`` applies it.  Gradient clipping must run after ``sync()`` (on the averaged
gradient) to match single-process semantics (train.py:273-274).
"""
import torch
import torch.distributed as dist


class FlatGradSync:
    def __init__(self, params, world_size=None, process_group=None, align=1, overlap=True, scale_in_optimizer=False,
                 bucket_counts=None):
        """``align``: every parameter's slice starts at a multiple of ``align`` elements (4 = 16 bytes: what
        ``optim.FusedClipAdam`` needs to lay the parameters out the same way; the gaps stay zero).
        ``scale_in_optimizer``: leave the 1/world scale to ``FusedClipAdam`` (it reads ``pending_scale()``).
        ``bucket_counts``: how many consecutive parameters form each all-reduce bucket (e.g. 4 per DCGRU cell:
        a layer's four gradients are produced by one backward call); default: buckets of >= 64 KB."""
        self.params = [p for p in params if p.requires_grad]
        # tied parameters appear once in .parameters(); keep it that way
        self.group = process_group
        self.world = world_size if world_size is not None else (
            dist.get_world_size(process_group) if dist.is_initialized() else 1)
        self.offsets, n = [], 0
        for p in self.params:
            n = (n + align - 1) // align * align
            self.offsets.append(n)
            n += p.numel()
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off: off + p.numel()].view_as(p)
        self.scale_in_optimizer = bool(scale_in_optimizer)
        self._scaled = True
        self._launched = 0
        # ---- buckets: maximal runs of parameters that belong to the same top-level child / layer -----------------
        self.overlap = bool(overlap) and self.world > 1
        self._cuda = dev.type == "cuda"
        self._hooks, self._works = [], []
        if self.overlap:
            if self._cuda:
                self.side = torch.cuda.Stream(device=dev)
                self._ready = torch.cuda.Event()
                self._done = torch.cuda.Event()
            self._bucket_of, self._buckets = {}, []          # id(p) -> bucket index; bucket = [lo, hi, pending, total]
            self._make_buckets(bucket_counts)
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _make_buckets(self, counts, min_elems=16384):
        """consecutive parameters are grouped by ``counts`` (the last bucket takes the rest) or, without it, until a
        bucket holds at least ``min_elems`` (64 KB)"""
        lo, cnt = 0, 0
        members = []
        counts = list(counts) if counts else None
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            members.append(p)
            cnt += p.numel()
            last = i == len(self.params) - 1
            if counts is not None:
                full = bool(counts) and len(members) == counts[0]
                if full:
                    counts.pop(0)
            else:
                full = cnt >= min_elems
            if full or last:
                hi = self.flat.numel() if last else self.offsets[i + 1]
                b = len(self._buckets)
                self._buckets.append([lo, hi, len(members), len(members)])
                for q in members:
                    self._bucket_of[id(q)] = b
                lo, cnt, members = hi, 0, []

    def _on_grad(self, p):
        b = self._buckets[self._bucket_of[id(p)]]
        b[2] -= 1
        if b[2] == 0:
            b[2] = b[3]
            if self._cuda:
                cur = torch.cuda.current_stream()
                self._ready.record(cur)
                self.side.wait_event(self._ready)
                with torch.cuda.stream(self.side):
                    dist.all_reduce(self.flat[b[0]: b[1]], op=dist.ReduceOp.SUM, group=self.group)
            else:                                             # host tensors (gloo): asynchronous work handles
                self._works.append(dist.all_reduce(self.flat[b[0]: b[1]], op=dist.ReduceOp.SUM, group=self.group,
                                                   async_op=True))
            self._launched += 1

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def zero(self):
        """replaces optimizer.zero_grad(): keeps the .grad views alive"""
        self.flat.zero_()
        self._launched = 0
        if self.overlap:
            for b in self._buckets:
                b[2] = b[3]

    def pending_scale(self):
        """the factor the optimiser still has to apply to the flat gradient (1/world after an unscaled sync)"""
        s = 1.0 if self._scaled else 1.0 / self.world
        self._scaled = True
        return s

    def sync(self):
        """sum (and, unless left to the optimiser, average) the gradient over all ranks"""
        if self.world <= 1:
            return
        launched = self._launched if self.overlap else 0
        if self.overlap and launched == len(self._buckets):
            if self._cuda:                                     # every bucket is in flight on the side stream: join it
                self._done.record(self.side)
                torch.cuda.current_stream().wait_event(self._done)
            for w in self._works:
                w.wait()
            self._works = []
        elif launched == 0:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        else:
            raise RuntimeError("FlatGradSync: some gradient buckets never became ready (a parameter received no "
                               "gradient this step); construct with overlap=False for such models")
        if self.scale_in_optimizer:
            self._scaled = False
        else:
            self.flat.mul_(1.0 / self.world)


def broadcast_parameters(module, src=0, process_group=None):
    """make every rank start from rank ``src``'s weights (one flat broadcast)"""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    ps = list(module.parameters())
    flat = torch.cat([p.detach().reshape(-1) for p in ps])
    dist.broadcast(flat, src=src, group=process_group)
    off = 0
    with torch.no_grad():
        for p in ps:
            p.copy_(flat[off: off + p.numel()].view_as(p))
            off += p.numel()


def shard_batch(t, rank, world, dim=0):
    """contiguous batch split B/world per rank"""
    n = t.shape[dim]
    if n % world:
        raise ValueError(f"batch {n} is not divisible by world size {world}")
    k = n // world
    return t.narrow(dim, rank * k, k)
