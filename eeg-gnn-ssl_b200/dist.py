"""Batch-sharded data parallelism for the DCGRU models: one process per GPU, weights replicated,
ONE all-reduce of a single flat fp32 gradient buffer per step.

The reference is single-process; this is the B200 addition.  Clips are independent, so
the only exchange is the parameter gradient: every ``p.grad`` is a view into one flat buffer, the
backward kernels' results are accumulated into it by autograd, and ``sync()`` issues one
``all_reduce(SUM)`` over NCCL/NVLink followed by the 1/world scale.  Gradient clipping must run
after ``sync()`` (on the averaged gradient) to match single-process semantics (train.py:273-274).
"""
import torch
import torch.distributed as dist


class FlatGradSync:
    def __init__(self, params, world_size=None, process_group=None, align=1):
        """``align``: every parameter's slice starts at a multiple of ``align`` elements (4 = 16 bytes: what
        ``optim.FusedClipAdam`` needs to lay the parameters out the same way; the gaps stay zero)."""
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

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def zero(self):
        """replaces optimizer.zero_grad(): keeps the .grad views alive"""
        self.flat.zero_()

    def sync(self):
        """average the gradient over all ranks with one collective"""
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
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
