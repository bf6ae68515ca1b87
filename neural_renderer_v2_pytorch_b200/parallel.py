"""Multi-GPU plumbing for the renderer: one process per GPU (``torchrun``), the batch of views
sharded across ranks, NCCL all-reduce ONLY for gradients of tensors shared by every view.

The reference has nothing here ("multiple GPUs" in its README means ``--gpu N``,
``examples_pytorch/example1.py:46-47``).  The path shards trivially: every view is independent in
the forward pass and in the per-view backward pass, so there is no data-path collective.  The one
real exchange is multi-view optimisation of ONE mesh / texture (``examples_pytorch/example2.py``):
the parameter is expanded to the local views, its gradient is summed over them on the device and
then summed over ranks (12*nv bytes for vertices, 12*T for textures).

Works with any ``torch.distributed`` backend: ``nccl`` on the B200 box, ``gloo`` in the CPU tests.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(num_views, rank=None, world_size=None):
    """Contiguous, balanced [begin, end) slice of ``num_views`` owned by ``rank``.
    The first ``num_views % world_size`` ranks get one extra view."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, extra = divmod(num_views, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_views(tensor, rank=None, world_size=None):
    """Slice dim 0 (views) of ``tensor`` for this rank."""
    b, e = shard_range(tensor.shape[0], rank, world_size)
    return tensor[b:e]


class _ShareAcrossViews(torch.autograd.Function):
    """[1, ...] parameter -> [views, ...] expanded view; backward sums the gradient over the local
    views and all-reduces it over ranks so every rank holds the full multi-view gradient."""

    @staticmethod
    def forward(ctx, param, views, group):
        ctx.group = group
        return param.expand(views, *param.shape[1:])

    @staticmethod
    def backward(ctx, grad):
        g = grad.sum(0, keepdim=True).contiguous()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None, None


def share_across_views(param, local_views, group=None):
    """Use one mesh / texture ``param`` of shape [1, ...] for ``local_views`` views on this rank.
    After ``backward()`` ``param.grad`` is identical on every rank: the sum over ALL views of the job."""
    assert param.shape[0] == 1, "shared parameter must have a leading dimension of 1"
    return _ShareAcrossViews.apply(param, local_views, group)


class _ShareAcrossRanks(torch.autograd.Function):
    """Identity; the backward sums the gradient over the ranks (in place, on the stream of the backward, so a
    captured step replays the collective too)."""

    @staticmethod
    def forward(ctx, param, group):
        ctx.group = group
        return param.view_as(param)

    @staticmethod
    def backward(ctx, grad):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            grad = grad.contiguous()
            dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=ctx.group)
        return grad, None


def share_across_ranks(param, group=None):
    """Mark ``param`` (e.g. ONE mesh [1,nv,3] handed to ``Renderer.render*`` with per-view ``viewpoints`` [B,3]:
    the fused camera transform projects it into the local views and sums its gradient over them) as shared by
    every rank: after ``backward()`` ``param.grad`` is the sum over ALL views of the job, identical on every
    rank.  The all-reduce moves 12*nv bytes (vertices) or 12*T (textures) per step."""
    return _ShareAcrossRanks.apply(param, group)


def allreduce_shared_grads(params, group=None):
    """Sum ``p.grad`` over ranks for parameters that every rank holds a replica of (when they were
    not routed through :func:`share_across_views`)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for p in params:
        if p.grad is not None:
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)


def gather_images(local_images, group=None):
    """All-gather per-rank image batches (equal local batch on every rank) into the global batch
    in rank order.  Not on the training path; for inspection / evaluation."""
    rank, ws = world()
    if ws == 1:
        return local_images
    out = [torch.empty_like(local_images) for _ in range(ws)]
    dist.all_gather(out, local_images.contiguous(), group=group)
    return torch.cat(out, 0)
