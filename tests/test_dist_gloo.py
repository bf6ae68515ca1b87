"""world_size-2 tests of the multi-GPU plumbing on CPU with the gloo backend (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_renderer_v2_pytorch_b200 import parallel


def test_shard_range_covers_everything():
    for n in (1, 7, 8, 64, 65):
        for ws in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    t = torch.arange(10)
    assert parallel.shard_views(t, 1, 3).tolist() == [4, 5, 6]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        torch.manual_seed(0)
        nv, views = 5, 6
        shared = torch.nn.Parameter(torch.randn(1, nv, 3))                 # one mesh for every view
        weights = torch.randn(views, nv, 3)                                 # stands in for the renderer
        b, e = parallel.shard_range(views)
        local = parallel.share_across_views(shared, e - b)
        assert local.shape == (e - b, nv, 3)
        (local * weights[b:e]).sum().backward()
        want = weights.sum(0, keepdim=True)                                 # single-process result
        ok1 = torch.allclose(shared.grad, want, atol=1e-6)
        # replicas path
        p = torch.nn.Parameter(torch.ones(3))
        p.grad = torch.full((3,), float(rank + 1))
        parallel.allreduce_shared_grads([p])
        ok2 = torch.allclose(p.grad, torch.full((3,), float(sum(range(1, ws + 1)))))
        imgs = parallel.gather_images(torch.full((2, 1, 4, 4), float(rank)))
        ok3 = imgs.shape == (2 * ws, 1, 4, 4) and imgs[2 * rank].eq(rank).all().item()
        out.put((rank, bool(ok1), bool(ok2), bool(ok3), parallel.world()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_shared_mesh_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, ok1, ok2, ok3, w in res:
        assert ok1 and ok2 and ok3, (rank, ok1, ok2, ok3)
        assert w == (rank, 2)


def test_single_process_is_a_no_op():
    assert parallel.world() == (0, 1)
    p = torch.nn.Parameter(torch.randn(1, 4, 3))
    y = parallel.share_across_views(p, 3)
    y.sum().backward()
    assert torch.allclose(p.grad, torch.full_like(p, 3.0))


def test_fused_exchange_default(monkeypatch):
    """parallel.fused_allowed: the NVLink exchange inside the camera backward is the default for every world size
    its peer table holds (profiles/r2_scaling_cfg3.jsonl); NR_FUSED_ALLREDUCE=0 keeps ncclAllReduce."""
    from neural_renderer_v2_pytorch_b200 import parallel
    monkeypatch.setattr(parallel, "FUSED_ALLREDUCE", True)
    assert [parallel.fused_allowed(n) for n in (2, 4, 8, 16, 32)] == [True, True, True, True, False]
    monkeypatch.setattr(parallel, "FUSED_ALLREDUCE", False)
    assert not parallel.fused_allowed(2)


def test_bench_shards_config2_strongly():
    """bench.shard_of: config 2's batch of 64 is split over the ranks (contiguous, balanced, complete); the other
    workloads keep their batch per GPU."""
    import bench
    w = bench.WORKLOADS["cfg2"]
    for world in (1, 2, 3, 4, 8):
        got = [bench.shard_of(w, r, world, "strong") for r in range(world)]
        assert all(g[0] == 64 for g in got)
        assert got[0][1] == 0 and got[-1][2] == 64
        assert all(got[i][2] == got[i + 1][1] for i in range(world - 1))
        sizes = [hi - lo for _, lo, hi in got]
        assert max(sizes) - min(sizes) <= 1
    V, lo, hi = bench.shard_of(bench.WORKLOADS["cfg3"], 3, 8, "weak")
    assert (V, lo, hi) == (64, 24, 32)
