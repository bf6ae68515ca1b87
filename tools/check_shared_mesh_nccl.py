#!/usr/bin/env python
"""Run under torchrun on N GPUs:  multi-view optimisation of ONE shared mesh.

Every rank renders its shard of the views of a shared [1,nv,3] parameter
(parallel.share_across_views), the backward all-reduces the gradient over NCCL.  Rank 0 then renders
ALL views alone and compares: the N-rank gradient must equal the single-process gradient on the
concatenated batch (sum order is the only difference).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_shared_mesh_nccl.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_renderer_v2_pytorch_b200 as nr  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "teapot.npz"))
    faces = torch.from_numpy(d["faces"]).to(dev)
    views, S = 4 * world, 128
    g = torch.Generator().manual_seed(0)
    eye = nr.get_points_from_angles(torch.full((views,), 2.732), torch.rand(views, generator=g) * 80 - 20,
                                    torch.rand(views, generator=g) * 360).to(dev)
    G = torch.randn((views, S, S), generator=g).to(dev)

    def loss_of(param, lo, hi, shared_fn):
        vs = nr.perspective(nr.look_at(shared_fn(param, hi - lo), eye[lo:hi]))
        hp = nr.RasterizeHyperparam(image_size=S, anti_aliasing=False)
        hp.deterministic = True
        img = nr.rasterize_silhouettes(vs, faces, nr.RasterizeParam(), hp)
        return (img * G[lo:hi]).sum()

    param = torch.from_numpy(d["vertices"])[None].to(dev).requires_grad_(True)
    lo, hi = nr.parallel.shard_range(views)
    loss_of(param, lo, hi, nr.parallel.share_across_views).backward()
    sharded = param.grad.clone()
    ok = True
    if rank == 0:
        ref = torch.from_numpy(d["vertices"])[None].to(dev).requires_grad_(True)
        loss_of(ref, 0, views, lambda p, n: p.expand(n, -1, -1)).backward()
        err = (sharded - ref.grad).abs().max().item() / ref.grad.abs().max().item()
        ok = err < 1e-5
        print("shared-mesh gradient, %d ranks vs 1 process: max rel-to-scale error %.3g -> %s" % (world, err, "OK" if ok else "MISMATCH"))
    # every rank must hold the same all-reduced gradient
    gathered = [torch.empty_like(sharded) for _ in range(world)]
    dist.all_gather(gathered, sharded)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    if rank == 0:
        print("all ranks hold identical gradients:", same)

    # ---- the same through Renderer on the [1,nv,3] mesh itself: the camera backward exchanges the gradient slices
    # with the peers over NVLink inside the kernel (parallel._Exchange), no NCCL call in the step
    rend = nr.Renderer()
    rend.image_size, rend.anti_aliasing, rend.viewpoints, rend.deterministic = S, False, eye[lo:hi], True

    def fused_step(p):
        img = rend.render_silhouettes(nr.parallel.share_across_ranks(p), faces)
        (img * G[lo:hi]).sum().backward()

    # reference for it: the same Renderer step with the exchange switched off (NCCL all-reduce of the [1,nv,3] sum)
    nr.parallel.FUSED_ALLREDUCE = False
    p1 = torch.from_numpy(d["vertices"])[None].to(dev).requires_grad_(True)
    fused_step(p1)
    nr.parallel.FUSED_ALLREDUCE = True
    p2 = torch.from_numpy(d["vertices"])[None].to(dev).requires_grad_(True)
    fused_step(p2)
    ex = nr.parallel._Exchange.get(p2.shape[1], None, dev)
    used = ex is not None
    scale = p1.grad.abs().max().item()
    err2 = (p2.grad - p1.grad).abs().max().item() / scale         # summation order over the ranks is the only difference
    # many steps in a row (epochs, buffer parity), eagerly and replayed from a CUDA graph; every step must give the
    # same bits (deterministic rasterizer backward + rank-ordered exchange)
    first = p2.grad.clone()
    stable = True
    for _ in range(40):
        p2.grad = None
        fused_step(p2)
        stable &= torch.equal(p2.grad, first)
    replay = nr.capture_step(lambda: fused_step(p2), params=[p2], warmup=2)
    for _ in range(40):
        replay()
        stable &= torch.equal(p2.grad, first)
    torch.cuda.synchronize()
    timed_out = int(ex.epoch[2]) if used else 0
    g2 = [torch.empty_like(first) for _ in range(world)]
    dist.all_gather(g2, p2.grad)
    same2 = all(torch.equal(g2[0], t) for t in g2)
    ok2 = err2 < 2e-6 and stable and same2 and not timed_out
    if rank == 0:
        print("fused exchange in use: %s; vs NCCL path: max rel-to-scale error %.3g; 80 steps bit-stable: %s; identical on "
              "all ranks: %s; peer time-outs: %d -> %s" % (used, err2, stable, same2, timed_out, "OK" if ok2 else "MISMATCH"))
    ok = ok and ok2
    flag = torch.tensor([1 if (ok and same) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if int(flag.item()) else 1)      # (no NCCL teardown: a captured graph holds the communicator)


if __name__ == "__main__":
    main()
