"""``differentiation(images, coordinates)``: identity in the forward pass, v2 approximate gradient
in the backward pass (reference: ``neural_renderer_torch/differentiation.py:6-40``).

The backward is one CUDA kernel (``nr_differentiation_backward``) instead of ~40 elementwise /
cat / masked-assign launches.  Inside ``rasterize_*`` the same stencil runs fused with the vertex
scatter; this standalone op exists because the reference exports it (``__init__.py:12``)."""
import ctypes

import torch

from . import _lib


class Differentiation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, coordinates):
        ctx.save_for_backward(images)
        return images.view_as(images)

    @staticmethod
    def backward(ctx, gradients):
        images, = ctx.saved_tensors
        if not images.is_cuda:
            raise RuntimeError("images must be a CUDA tensor")
        assert images.ndim == 4 and images.shape[1] == images.shape[2], "images must be [B, S, S, C]"
        img = images.detach().to(torch.float32).contiguous()
        g = gradients.detach().to(torch.float32).contiguous()
        B, R, _, C = img.shape
        grad_xy = torch.empty((B, R, R, 2), dtype=torch.float32, device=img.device)
        with torch.cuda.device(img.device):
            stream = torch.cuda.current_stream(img.device).cuda_stream
            rc = _lib.lib().nr_differentiation_backward(
                ctypes.c_void_p(img.data_ptr()), ctypes.c_void_p(g.data_ptr()),
                ctypes.c_void_p(grad_xy.data_ptr()), B, R, C, ctypes.c_void_p(stream))
        _lib.check(rc, "nr_differentiation_backward")
        return gradients, grad_xy


def differentiation(images, coordinates):
    return Differentiation.apply(images, coordinates)
