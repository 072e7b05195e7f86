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

    def start(self, lo, hi):
        """Begin the all-reduce of flat[lo:hi] (aliased buffer only) asynchronously: with NCCL the collective runs on
        the backend's own stream behind the work enqueued so far, so kernels launched after this call -- the coarse
        network's backward while the fine network's gradients travel -- overlap with it.  Returns a handle for finish()."""
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world == 1 or hi <= lo:
            return None
        if not self.aliased:
            raise RuntimeError("FlatGradAllReduce.start needs the aliased flat buffer")
        return dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True)

    def finish(self, handles, scale=True):
        """Wait for the started pieces (the current stream waits, not the host) and turn the sums into means
        (scale=False: the caller folds the 1 / world_size into its optimizer kernel).  Returns 1 / world_size."""
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world == 1:
            return 1.0
        for h in handles:
            if h is not None:
                h.wait()
        if scale:
            self.flat.mul_(1.0 / world)
        return 1.0 / world

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


MAP_KEYS = (("rgb_map_0", 3), ("disp_map_0", 1), ("acc_map_0", 1), ("depth_map_0", 1),
            ("rgb_map", 3), ("disp_map", 1), ("acc_map", 1), ("depth_map", 1))
FLOATS_PER_RAY = sum(c for _, c in MAP_KEYS)     # 12


def block_views(blk, per):
    """The eight maps of one ray block as views of ONE flat buffer of 12*per floats (map after map, each map
    contiguous: rgb [per,3], the others [per]) -- the block travels to rank 0 in a single collective."""
    out, off = {}, 0
    for k, c in MAP_KEYS:
        v = blk[off * per:(off + c) * per]
        out[k] = v.view(per, 3) if c == 3 else v
        off += c
    return out


def assemble_blocks(gathered, n, per, world):
    """gathered: [world, 12*per] (rank r's flat block in row r) -> dict of contiguous full-frame maps [n,3] / [n]."""
    out, off = {}, 0
    for k, c in MAP_KEYS:
        full = gathered[:, off * per:(off + c) * per].reshape(world * per, c)[:n]
        out[k] = full.contiguous() if c == 3 else full.reshape(-1).contiguous()
        off += c
    return out


class ShardedFrame:
    """One frame split into contiguous blocks of ceil(n/world) rays (north_star "ray batches are sharded
    across the 8 GPUs of one box with no inter-GPU traffic for rendering"): every rank renders its block with
    `render_block(lo, hi, views)` -- which must fill views[k][:hi-lo] for the eight maps -- and the blocks are then
    gathered on rank `dst` (the only collective, after the rendering: 48 bytes per ray).  Backend-agnostic: NCCL
    gather between GPUs, gloo in the CPU test."""

    def __init__(self, n, device, dst=0, group=None):
        self.n, self.dst, self.group = int(n), dst, group
        init = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if init else 1
        self.rank = dist.get_rank(group) if init else 0
        self.per = (self.n + self.world - 1) // self.world
        self.lo, self.hi = shard_range(self.n, self.rank, self.world)
        self.blk = torch.zeros(FLOATS_PER_RAY * self.per, dtype=torch.float32, device=device)
        self.views = block_views(self.blk, self.per)
        self.gathered = (torch.empty((self.world, FLOATS_PER_RAY * self.per), dtype=torch.float32, device=device)
                         if self.rank == dst else None)

    def __call__(self, render_block):
        """-> dict of full-frame maps on rank dst (device tensors), None elsewhere."""
        if self.hi > self.lo:
            render_block(self.lo, self.hi, self.views)
        if self.world == 1:
            return assemble_blocks(self.blk.view(1, -1), self.n, self.per, 1)
        if self.rank == self.dst:
            dist.gather(self.blk, list(self.gathered.unbind(0)), dst=self.dst, group=self.group)
            return assemble_blocks(self.gathered, self.n, self.per, self.world)
        dist.gather(self.blk, None, dst=self.dst, group=self.group)
        return None


class ShardedFrameRenderer:
    """Renderer.render(batch) with the rays of ONE frame split over the ranks (strong scaling / single-frame
    latency).  Rank 0 gets the reference's dict of [H,W,..] maps -- in pinned host memory with to_host=True -- the
    other ranks get None."""

    def __init__(self, renderer, to_host=True):
        self.r, self.to_host = renderer, to_host
        self._frame, self._host = None, None

    @torch.no_grad()
    def render(self, batch):
        from . import ops
        r = self.r
        H, W = int(batch["H"]), int(batch["W"])
        n = H * W
        if self._frame is None or self._frame.n != n:
            self._frame = ShardedFrame(n, r.device)
            self._host = None
        pose = batch["pose"].reshape(4, 4).to(r.device, torch.float32)
        K = batch["intrinsics"].reshape(3, 3).to(r.device, torch.float32)
        rays_o, rays_d = ops.raygen(pose, K, H, W)       # 50 us for the whole frame: every rank generates all rays

        def block(lo, hi, views):
            r.render_rays(rays_o[lo:hi], rays_d[lo:hi], out={k: v[:hi - lo] for k, v in views.items()})
        full = self._frame(block)
        if full is None:
            return None
        if self.to_host:
            if self._host is None:
                self._host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in full.items()}
            for k, v in full.items():
                self._host[k].copy_(v, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            full = self._host
        return {k: (v.view(H, W, 3) if k.startswith("rgb") else v.view(H, W)) for k, v in full.items()}
