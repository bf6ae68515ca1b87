"""CUDA-graph capture of a whole render step (forward + backward).

A render step at small batch is a dozen short kernels; launched eagerly from Python the host, not
the GPU, sets the pace.  ``capture_step`` records the step once (torch.cuda.graphs whole-step
capture: the rasterizer's kernels, its side-stream background fill and the autograd backward all
land in one graph) and returns a callable that replays it with ~10 us of host work.

    v = vertices.clone().requires_grad_(True)          # static input buffers
    def step():
        images = nr.rasterize_rgba(v, faces, params, hp)
        images.backward(upstream)                      # or a loss
        return images
    replay = nr.capture_step(step, params=[v], warmup=3)
    v.data.copy_(new_vertices); images = replay(); v.grad  # -> gradients of the replayed step

The reference has nothing comparable (every call synchronises with the host several times per view,
utils.py:111-112)."""
import torch
import torch.distributed


def capture_step(step, params=(), warmup=3):
    """Capture ``step()`` (a closure over STATIC input tensors) into a CUDA graph.

    ``params`` are the leaf tensors whose ``.grad`` the step produces; their ``.grad`` is reset before
    capture so that the captured backward writes (not accumulates) it.  Returns ``replay()``, which
    re-runs the step on whatever the static inputs currently hold and returns the step's outputs
    (static tensors, overwritten by the next replay)."""
    params = list(params)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(warmup, 2)):       # eager runs: size workspaces, validate indices
            for p in params:
                p.grad = None
            step()
            # each run sees the bin statistics of the one before it (pair capacity, binning mode), so
            # the captured step is the settled one
            side.synchronize()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for p in params:
        p.grad = None
    graph = torch.cuda.CUDAGraph()
    # capture on the stream the warm-up ran on: the rasterizer keeps one workspace per stream.
    # With a process group alive its watchdog thread polls CUDA events; "thread_local" keeps those calls
    # legal while this thread captures (a step may contain the all-reduce of parallel.share_across_views).
    mode = "global"
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        mode = "thread_local"
    from . import rasterize as _rasterize
    keep = _rasterize._capture_keepalive = []
    try:
        with torch.cuda.graph(graph, stream=side, capture_error_mode=mode):
            out = step()
    finally:
        _rasterize._capture_keepalive = None

    def replay():
        graph.replay()
        return out

    replay.graph = graph
    replay.workspaces = keep      # the rasterizer workspaces the graph replays on stay alive with it
    return replay
