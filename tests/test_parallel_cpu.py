"""world_size-2 gloo test (CPU) of the batch-sharding logic: shard bounds, global-index noise,
all_gather reassembly.  The enhance() itself is replaced by a cheap per-clip function with the
same independence property (the GPU twin is tests/test_backbone_gpu.py::test_batch_independence)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flowdec_b200 import parallel


def _fake_enhance(y, noise):
    # per-clip, batch-composition independent (like the real path)
    return y * 0.5 + noise.real.mean(dim=(-2, -1), keepdim=False).reshape(-1, 1, 1)


def _worker(rank, world, port, B, L, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    y = torch.randn(B, 1, L, generator=g)

    def gather(local, bounds):
        sizes = [hi - lo for lo, hi in bounds]
        bufs = [torch.empty(s, 1, L) for s in sizes]
        # gloo all_gather needs equal sizes: pad to max
        mx = max(sizes)
        pad = torch.zeros(mx, 1, L)
        pad[:local.shape[0]] = local
        outs = [torch.empty(mx, 1, L) for _ in range(world)]
        dist.all_gather(outs, pad)
        return torch.cat([o[:s] for o, s in zip(outs, sizes)], 0)

    full = parallel.enhance_sharded(_fake_enhance, y, 3, "midpoint", rank, world, gather=gather)
    if rank == 0:
        ret.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_single():
    B, L = 5, 1536
    g = torch.Generator().manual_seed(0)
    y = torch.randn(B, 1, L, generator=g)
    single, (lo, hi) = parallel.enhance_sharded(_fake_enhance, y, 3, "midpoint", 0, 1)
    assert (lo, hi) == (0, B)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, B, L, ret)) for r in range(2)]
    for p in procs:
        p.start()
    full = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(full, single)


def test_shard_bounds_cover():
    for B in (1, 5, 32, 256):
        for w in (1, 2, 4, 8):
            b = [parallel.shard_bounds(B, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == B
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    a = parallel.clip_noise(7, 64)
    assert torch.equal(a, parallel.clip_noise(7, 64)) and not torch.equal(a, parallel.clip_noise(8, 64))


class _FakeModel:
    """stands in for FlowModel on CPU: per-clip, batch-composition independent, consumes the injected noise"""
    device = torch.device("cpu")

    def enhance(self, y, N, solver, noise):
        return y * 0.25 + noise.real.mean(dim=(-2, -1)).reshape(-1, 1, 1) * N


def _worker_verify(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = parallel.verify_sharding(_FakeModel(), dist, 1536, 3, "midpoint", clips_per_rank=3)
    if rank == 0:
        ret.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_enhance_gather_world2():
    """the collective form bench.py runs on hardware at N > 1 (scatter from rank 0 -> per-rank enhance with
    global-index noise -> all_gather -> bitwise comparison with the unsharded run), here over gloo"""
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker_verify, args=(r, 2, 29613, ret)) for r in range(2)]
    for p in procs:
        p.start()
    ok = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok is True
