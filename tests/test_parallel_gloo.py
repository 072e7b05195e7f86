"""CPU, world_size 2 over gloo: ray/view sharding partitions the work exactly, timing reduction is a
max over ranks, and the flat-gradient all-reduce averages like DistributedDataParallel would."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_rep_for_test_b200 import parallel as P


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 640000 + 7
        lo, hi = P.shard_range(n, rank, world)
        cover = torch.zeros(n, dtype=torch.int32)
        cover[lo:hi] += 1
        dist.all_reduce(cover)
        assert int(cover.min()) == 1 and int(cover.max()) == 1          # exact partition, no overlap
        views = torch.zeros(200, dtype=torch.int32)
        views[P.views_for_rank(200, rank, world)] += 1
        dist.all_reduce(views)
        assert bool((views == 1).all())
        t = P.max_over_ranks([10.0 + rank, 5.0 - rank], torch.device("cpu"))
        assert t == [10.0 + world - 1, 5.0]
        # gradient averaging
        torch.manual_seed(0)
        lin = torch.nn.Linear(7, 3)
        x = torch.full((4, 7), float(rank + 1))
        lin(x).sum().backward()
        local = [p.grad.clone() for p in lin.parameters()]
        P.FlatGradAllReduce(lin.parameters())()
        mean_scale = sum(range(1, world + 1)) / world / (rank + 1)
        for g, l in zip([p.grad for p in lin.parameters()], local):
            expect = l * mean_scale if l.dim() == 2 else l        # bias grad does not depend on x
            assert torch.allclose(g, expect, atol=1e-6), (g, expect)
        # aliased variant (training.TrainStep): .grad tensors are views of one flat buffer, reduced in place
        lin2 = torch.nn.Linear(7, 3)
        ps = list(lin2.parameters())
        flat = torch.zeros(sum(p.numel() for p in ps))
        off = 0
        for p in ps:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        flat.fill_(float(rank + 1))
        P.FlatGradAllReduce(ps, flat=flat)()
        assert torch.allclose(flat, torch.full_like(flat, sum(range(1, world + 1)) / world))
        assert all(p.grad.data_ptr() >= flat.data_ptr() for p in ps)
        # the two-piece overlapped variant TrainStep uses: start(fine slice) ... start(coarse slice), finish
        flat.fill_(float(rank + 1))
        ar = P.FlatGradAllReduce(ps, flat=flat)
        cut = ps[0].numel()
        hs = [ar.start(cut, flat.numel()), ar.start(0, cut)]
        ar.finish(hs)
        assert torch.allclose(flat, torch.full_like(flat, sum(range(1, world + 1)) / world))
        # one frame split into contiguous ray blocks, gathered on rank 0 (ShardedFrame: the N>1 single-frame path)
        for n in (1003, 64, 1):
            frame = P.ShardedFrame(n, torch.device("cpu"))
            assert (frame.lo, frame.hi) == P.shard_range(n, rank, world)

            def block(lo, hi, views):
                idx = torch.arange(lo, hi, dtype=torch.float32)
                for j, (k, c) in enumerate(P.MAP_KEYS):
                    views[k][:hi - lo] = (idx[:, None] * 10 + j + torch.arange(c) * 0.25) if c == 3 else idx * 10 + j
            full = frame(block)
            if rank == 0:
                idx = torch.arange(n, dtype=torch.float32)
                for j, (k, c) in enumerate(P.MAP_KEYS):
                    expect = (idx[:, None] * 10 + j + torch.arange(c) * 0.25) if c == 3 else idx * 10 + j
                    assert full[k].is_contiguous() and torch.equal(full[k], expect), (n, k)
            else:
                assert full is None
        out.put((rank, "ok"))
    except Exception as e:      # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharding_and_grad_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    world, port = 2, 29533
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_range_edges():
    assert P.shard_range(0, 0, 4) == (0, 0)
    assert P.shard_range(3, 3, 4) == (3, 3)
    assert [P.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
