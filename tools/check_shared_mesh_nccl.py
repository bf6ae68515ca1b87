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
    dist.destroy_process_group()
    sys.exit(0 if (ok and same) else 1)


if __name__ == "__main__":
    main()
