"""Multi-GPU host logic: one process per GPU, rays sharded with NO data-path collective
for rendering; one flat-gradient all-reduce per step for data-parallel training (replaces the
DistributedDataParallel wrap of src/train/trainers/trainer.py:14-22).  Backend-agnostic on purpose:
NCCL on the GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of `n` units for `rank`: ceil(n/world) per rank, last ranks may be short."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def views_for_rank(n_views, rank, world):
    """Round-robin view assignment for the 200-view test set (BASELINE.json configs[3])."""
    return list(range(rank, n_views, world))


def max_over_ranks(values, device):
    """Element-wise max of a list of python floats over all ranks (timing: max over ranks)."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


class FlatGradAllReduce:
    """Averages the gradients of `params` over all ranks with ONE all-reduce of a flat fp32 buffer
    (1 191 688 elements = 4.77 MB for the two NeRF MLPs).  With `flat` given, the parameters' .grad tensors are
    expected to alias that buffer already (training.TrainStep) and nothing is copied."""

    def __init__(self, params, flat=None):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.aliased = flat is not None
        self.flat = flat if flat is not None else torch.zeros(n, dtype=torch.float32, device=p0.device)
        if self.flat.numel() != n:
            raise ValueError("flat gradient buffer has %d elements, parameters have %d" % (self.flat.numel(), n))

    def __call__(self):
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world == 1:
            return
        if self.aliased:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / world)
            return
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / world)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(self.flat[off:off + n].view_as(p))
            off += n
