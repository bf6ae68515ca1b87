"""Fused camera transform: world-space vertices + per-view camera -> screen space in ONE kernel,
with ONE backward kernel.

Replaces, inside ``Renderer.transform_vertices`` (reference ``renderer.py:24-35``), the chain
``look_at`` / ``look`` (``look_at.py:28-42``, ``look.py:27-40``) -> ``perspective``
(``perspective.py:9-17``): ~10 elementwise / bmm launches forward and ~15 backward over [B,nv,3]
tensors.  The 3x3 camera rotation is still built with torch ops on [B,3] tensors (so gradients to
``viewpoints`` keep flowing through normalize / cross exactly as in the reference); only the O(nv)
part is fused.  The standalone ``look_at`` / ``look`` / ``perspective`` functions remain plain torch.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib
from .look_at import _as_batch, _camera_rotation
from .perspective import _PI_REF


class _CameraTransform(torch.autograd.Function):
    """vertices [B,nv,3], or ONE mesh [1,nv,3] seen by all B cameras (``rotation`` [B,3,3], ``eye`` [B,3]):
    the kernels broadcast it, and the backward sums its gradient over the views in registers, view by view
    (deterministic, no [B,nv,3] intermediate and no separate reduction)."""

    @staticmethod
    def forward(ctx, vertices, rotation, eye, perspective, width, exchange=None):
        ctx.exchange = exchange
        v = vertices.detach().to(torch.float32).contiguous()
        r = rotation.detach().to(torch.float32).contiguous()
        e = eye.detach().to(torch.float32).contiguous()
        B, nv = r.shape[0], v.shape[1]
        shared = int(v.shape[0] == 1 and B > 1)
        out = torch.empty((B, nv, 3), dtype=torch.float32, device=v.device)
        with torch.cuda.device(v.device):
            stream = torch.cuda.current_stream(v.device).cuda_stream
            rc = _lib.lib().nr_camera_forward(ctypes.c_void_p(v.data_ptr()), ctypes.c_void_p(r.data_ptr()),
                                              ctypes.c_void_p(e.data_ptr()), ctypes.c_void_p(out.data_ptr()), B, nv,
                                              int(perspective), float(width), shared, ctypes.c_void_p(stream))
        _lib.check(rc, "nr_camera_forward")
        ctx.save_for_backward(v, r, e)
        ctx.perspective, ctx.width, ctx.shared = int(perspective), float(width), shared
        ctx.in_dtype = vertices.dtype
        return out

    @staticmethod
    def backward(ctx, grad_out):
        v, r, e = ctx.saved_tensors
        g = grad_out.detach().to(torch.float32).contiguous()
        B, nv = r.shape[0], v.shape[1]
        L = _lib.lib()
        gv = torch.empty_like(v)
        need_cam = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        partial = None
        if need_cam:
            partial = torch.empty((B, L.nr_camera_partial_blocks(nv), 12), dtype=torch.float32, device=v.device)
        ex = ctx.exchange
        with torch.cuda.device(v.device):
            stream = torch.cuda.current_stream(v.device).cuda_stream
            pp = ctypes.c_void_p(partial.data_ptr()) if partial is not None else None
            if ex is not None:
                # shared mesh on several GPUs: the sum over the local views AND over the ranks, in this one kernel
                rc = L.nr_camera_backward_shared_allreduce(
                    ctypes.c_void_p(v.data_ptr()), ctypes.c_void_p(r.data_ptr()), ctypes.c_void_p(e.data_ptr()),
                    ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(gv.data_ptr()), pp, B, nv, ctx.perspective, ctx.width,
                    ex.rank, ex.world, ex.pointers, ctypes.c_void_p(ex.epoch.data_ptr()), ctypes.c_void_p(stream))
            else:
                rc = L.nr_camera_backward(ctypes.c_void_p(v.data_ptr()), ctypes.c_void_p(r.data_ptr()),
                                          ctypes.c_void_p(e.data_ptr()), ctypes.c_void_p(g.data_ptr()),
                                          ctypes.c_void_p(gv.data_ptr()), pp, B, nv,
                                          ctx.perspective, ctx.width, ctx.shared, ctypes.c_void_p(stream))
        _lib.check(rc, "nr_camera_backward")
        if gv.dtype != ctx.in_dtype:
            gv = gv.to(ctx.in_dtype)
        if not need_cam:
            return gv, None, None, None, None, None
        red = partial.sum(1)                               # fixed-order reduction of the per-block sums
        return gv, red[:, :9].reshape(B, 3, 3), red[:, 9:], None, None, None


def transform_vertices(vertices, viewpoints, camera_mode="look_at", camera_direction=None, perspective=True,
                       viewing_angle=30., at=None, up=None):
    """Screen-space vertices [B,nv,3] of ``vertices`` [B,nv,3] (CUDA) seen from ``viewpoints``.

    ``vertices`` may also be ONE mesh [1,nv,3] with ``viewpoints`` [B,3] (multi-view optimisation of a shared mesh,
    examples_pytorch/example2.py): it is projected into every view, and its gradient is the sum over the views."""
    assert vertices.ndim == 3
    dev, B = vertices.device, vertices.shape[0]
    if B == 1:
        for t in (viewpoints, camera_direction if camera_mode == "look" else None, at, up):
            if torch.is_tensor(t) and t.ndim == 2 and t.shape[0] > 1:
                B = t.shape[0]
                break
    eye = _as_batch(viewpoints, None, B, dev)
    up = _as_batch(up, [0., 1., 0.], B, dev)
    if camera_mode == "look_at":
        z_axis = F.normalize(_as_batch(at, [0., 0., 0.], B, dev) - eye, dim=-1)
    elif camera_mode == "look":
        z_axis = F.normalize(_as_batch(camera_direction, [0., 0., 1.], B, dev), dim=-1)
    else:                                                  # no viewpoint transformation (renderer.py:26-29)
        z_axis = None
    if z_axis is None:
        rot = torch.eye(3, device=dev).expand(B, 3, 3)
        eye = torch.zeros((B, 3), device=dev)
    else:
        rot = _camera_rotation(z_axis, up)
    if torch.is_tensor(viewing_angle):
        raise TypeError("the fused transform takes a scalar viewing angle; use look_at + perspective for a tensor")
    width = float(torch.tan(torch.tensor(float(viewing_angle), dtype=torch.float32) / 180. * _PI_REF))
    exchange = None
    shared = getattr(vertices, "_nr_shared", None)
    if shared is not None and vertices.shape[0] == 1 and B > 1 and vertices.shape[1] <= 256 * 128 * 4:      # (its grid must be resident)
        # a mesh shared by every rank (parallel.share_across_ranks): exchange its gradient inside the camera backward
        from . import parallel
        exchange = parallel._Exchange.get(vertices.shape[1], shared[1], vertices.device)
        if exchange is not None:
            vertices = shared[0]
    return _CameraTransform.apply(vertices, rot, eye.expand(B, 3), bool(perspective), width, exchange)
